"""TEST INFRASTRUCTURE — deterministic, torch-RNG-independent weights and inputs shared by the golden-vector
generator (run against the real reference), the oracle and the GPU parity tests.

Parameter names/shapes follow the reference's state_dict exactly (SURVEY.md Appendix A;
codes/models/interaction_transformer.py:397-509) so the same dict loads with strict=True into the reference
model, the oracle and the B200 module.  Values come from numpy's legacy RandomState (bit-stable across numpy
versions), drawn in sorted-key order.  The reference zero-initialises out/out2, every StylizationBlock
out_layers.2 and FFN.linear2 (zero_module, :62-68) which makes the default-init network the identity map and
any parity check vacuous, so those get N(0, 0.02^2) here ("de-zeroing").
"""
import numpy as np
import torch


def param_shapes(num_layers=8, latent_dim=512, ff_size=1024, input_feats=263, num_frames=196, text_latent_dim=256,
                 cap_id=True):
    D, F, E, Dt = latent_dim, ff_size, 4 * latent_dim, text_latent_dim
    sh = {}
    if cap_id:
        sh["cap_embedding"] = (43, Dt)
    sh["text_proj.0.weight"] = (E, Dt)
    sh["text_proj.0.bias"] = (E,)
    sh["sequence_embedding"] = (num_frames, D)
    sh["joint_embed.weight"] = (D, input_feats)
    sh["joint_embed.bias"] = (D,)
    sh["joint_embed2.weight"] = (D, 4)
    sh["joint_embed2.bias"] = (D,)
    sh["time_embed.0.weight"] = (E, D)
    sh["time_embed.0.bias"] = (E,)
    sh["time_embed.2.weight"] = (E, E)
    sh["time_embed.2.bias"] = (E,)

    def styl(p):
        sh[p + "proj_out.emb_layers.1.weight"] = (2 * D, E)
        sh[p + "proj_out.emb_layers.1.bias"] = (2 * D,)
        sh[p + "proj_out.norm.weight"] = (D,)
        sh[p + "proj_out.norm.bias"] = (D,)
        sh[p + "proj_out.out_layers.2.weight"] = (D, D)
        sh[p + "proj_out.out_layers.2.bias"] = (D,)

    for i in range(num_layers):
        b = f"temporal_decoder_blocks.{i}."
        for blk, kdim in (("sa_block.", D), ("ca_block.", Dt), ("int_ca_block.", D)):
            p = b + blk
            sh[p + "norm.weight"] = (D,)
            sh[p + "norm.bias"] = (D,)
            if blk == "ca_block.":
                sh[p + "text_norm.weight"] = (Dt,)
                sh[p + "text_norm.bias"] = (Dt,)
            sh[p + "query.weight"] = (D, D)
            sh[p + "query.bias"] = (D,)
            sh[p + "key.weight"] = (D, kdim)
            sh[p + "key.bias"] = (D,)
            sh[p + "value.weight"] = (D, kdim)
            sh[p + "value.bias"] = (D,)
            styl(p)
        p = b + "ffn."
        sh[p + "linear1.weight"] = (F, D)
        sh[p + "linear1.bias"] = (F,)
        sh[p + "linear2.weight"] = (D, F)
        sh[p + "linear2.bias"] = (D,)
        styl(p)
    sh["out.weight"] = (input_feats, D)
    sh["out.bias"] = (input_feats,)
    sh["out2.weight"] = (input_feats, D)
    sh["out2.bias"] = (input_feats,)
    return sh


_ZERO_INIT_SUFFIXES = ("out_layers.2.weight", "out_layers.2.bias", "linear2.weight", "linear2.bias")


def make_state_dict(seed=0, **cfg):
    """name -> fp32 torch tensor (CPU)."""
    rs = np.random.RandomState(seed)
    sd = {}
    shapes = param_shapes(**cfg)
    for name in sorted(shapes):
        shape = shapes[name]
        zero_init = name.endswith(_ZERO_INIT_SUFFIXES) or name.startswith(("out.", "out2."))
        if name in ("sequence_embedding", "cap_embedding"):
            v = rs.standard_normal(shape)
        elif "norm.weight" in name:
            v = 1.0 + 0.1 * rs.standard_normal(shape)
        elif "norm.bias" in name:
            v = 0.05 * rs.standard_normal(shape)
        elif zero_init:
            v = 0.02 * rs.standard_normal(shape)
        elif name.endswith(".bias"):
            v = 0.02 * rs.standard_normal(shape)
        else:  # dense weight [out, in]: default nn.Linear variance 1/(3 fan_in)
            v = rs.standard_normal(shape) / np.sqrt(3.0 * shape[1])
        sd[name] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return sd


def make_inputs(seed, S, T, C=263, n_text=1, text_latent_dim=256, time_embed_dim=2048, n_steps=1000, lengths=None,
                timesteps=None):
    """Seeded denoiser inputs: x [S,T,C], t [S], length [S], xf_proj [S,E], xf_out [S,n_text,Dt], cap ids."""
    rs = np.random.RandomState(seed)
    x = torch.from_numpy(rs.standard_normal((S, T, C)).astype(np.float32))
    if timesteps is None:
        tb = rs.randint(0, n_steps, size=(S // 2,))
        timesteps = np.concatenate([tb, tb])
    t = torch.from_numpy(np.asarray(timesteps, dtype=np.int64))
    if lengths is None:
        lb = rs.randint(max(1, T // 3), T + 1, size=(S // 2,))
        lengths = np.concatenate([lb, lb])
    length = torch.from_numpy(np.asarray(lengths, dtype=np.int64))
    xf_proj = torch.from_numpy((0.5 * rs.standard_normal((S, time_embed_dim))).astype(np.float32))
    xf_out = torch.from_numpy(rs.standard_normal((S, n_text, text_latent_dim)).astype(np.float32))
    cap = rs.randint(0, 43, size=(2, S // 2))
    return {"x": x, "t": t, "length": length, "xf_proj": xf_proj, "xf_out": xf_out,
            "cap1": torch.from_numpy(cap[0].astype(np.int64)), "cap2": torch.from_numpy(cap[1].astype(np.int64))}


def make_noise(seed, steps, S, T, C=263):
    """[steps+1, S, T, C]: index 0 is x_T, index 1+k the noise drawn at the k-th reverse step."""
    rs = np.random.RandomState(seed)
    return torch.from_numpy(rs.standard_normal((steps + 1, S, T, C)).astype(np.float32))


def make_joint_inputs(seed, S, T, C=263):
    """Seeded stand-ins for a sampled batch and the dataset statistics of tools/visualization.py:103-111 (mean/std of
    the 263 motion features, init_mean/init_std of the 4 init-state features; the real files are not available
    offline).  Returns (x [S,T,C] torch fp32, mean [C], std [C], init_mean [4], init_std [4]) — numpy float32."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(S, T, C, generator=g)
    mean = (torch.randn(C, generator=g) * 0.3).numpy()
    std = (torch.rand(C, generator=g) * 0.5 + 0.5).numpy()
    # yaw velocity is small in the data (radians per frame); keep the integrated yaw within a few turns
    mean[0], std[0] = 0.01, 0.05
    init_mean = (torch.randn(4, generator=g) * 0.5).numpy()
    init_std = (torch.rand(4, generator=g) * 0.5 + 0.5).numpy()
    return x, mean, std, init_mean, init_std
