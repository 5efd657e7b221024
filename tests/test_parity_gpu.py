"""Model-level parity on the B200, all through the public drop-in API (which calls the C ABI):
  * against the committed goldens produced by the REAL reference (tests/golden, oracle/make_golden.py),
  * against the oracle (CPU, fp32) on seeded inputs at full depth / full length,
  * size-independent reference properties at the BASELINE sizes (padding invariance, role-swap equivariance,
    batch-composition independence, identity at zero-init, t == 0 adds no noise).
Tolerances (BASELINE.json north_star): per-step rel-L2 <= 1e-2 in bf16, <= 1e-5 in fp32 mode."""
import ast
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

pytestmark = pytest.mark.gpu

TOL = {"bf16": 1e-2, "fp32": 1e-5}
# MPJPE (north_star: "final sampled joints within a stated MPJPE tolerance") of the joints recovered from a sampled
# x_{t-1} against the joints recovered from the reference's, in the de-normalised units of joint_stats() below (local
# joint coordinates ~ N(0.3, 0.75), root trajectory integrated over up to 63 frames): one reverse step from identical
# x_t.  Measured on B200: bf16 2.2e-3 .. 3.4e-3 at mean |joint| = 31..34 (1e-4 relative), fp32-mode 50-step chain 1.6e-6.
MPJPE_TOL = {"bf16": 2e-2, "fp32": 1e-4}
GOLDEN = os.path.join(ROOT, "tests", "golden")


def joint_stats():
    import weights
    _, mean, std, im, isd = weights.make_joint_inputs(16, 2, 2)
    return mean, std, im, isd


def joints_mpjpe(x_cuda, x_ref, length):
    """MPJPE over the valid frames between CUDA joints of x_cuda (hig_recover_joints) and oracle joints of x_ref."""
    import joints_oracle as JO
    from hig_b200 import motion_process as mp
    mean, std, im, isd = joint_stats()
    j = mp.joints_from_samples(x_cuda.float().contiguous(), mean, std, im, isd, length=length)
    jr = torch.from_numpy(JO.joints_from_samples(x_ref.float().cpu().numpy(), mean, std, im, isd))
    return mp.mpjpe(j.cpu(), jr, length=torch.as_tensor(length).cpu()).item(), jr.norm(dim=-1).mean().item()


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def valid_rel(a, b, length):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    T = a.shape[1]
    m = (torch.arange(T)[None] < torch.as_tensor(length).reshape(-1, 1).cpu())
    return ((a - b)[m].norm() / b[m].norm()).item()


def build(layers, precision, cuda, cap_id=True, seed=0, zero_init=False):
    import weights
    from hig_b200.interaction_transformer import MotionInteractionTransformer
    m = MotionInteractionTransformer(263, num_frames=196, num_layers=layers, latent_dim=512, cap_id=cap_id,
                                     precision=precision)
    sd = weights.make_state_dict(seed=seed, num_layers=layers)
    if zero_init:
        for k in sd:
            if k.endswith(("out_layers.2.weight", "out_layers.2.bias", "linear2.weight", "linear2.bias")) or \
                    k.startswith(("out.", "out2.")):
                sd[k] = torch.zeros_like(sd[k])
    m.load_state_dict(sd, strict=True)
    return m.to(cuda).eval(), sd


def load_case(name):
    import weights
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = ast.literal_eval(str(d["cfg"]))
    inp = weights.make_inputs(cfg["seed"], cfg["S"], cfg["T"], n_text=cfg.get("n_text", 1), lengths=cfg["lengths"],
                              timesteps=cfg.get("timesteps"))
    return d, cfg, inp


def run(m, inp, mode, cuda):
    g = lambda k: inp[k].to(cuda)
    with torch.no_grad():
        if mode == "cap":
            m.cap_id = True
            return m(g("x"), g("t"), length=g("length"), text=[g("cap1"), g("cap2")])
        m.cap_id = False
        return m(g("x"), g("t"), length=g("length"), xf_proj=g("xf_proj"), xf_out=g("xf_out"))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["fwd_cap", "fwd_text", "fwd_text_full"])
def test_forward_matches_reference_golden(cuda, name, precision):
    d, cfg, inp = load_case(name)
    m, _ = build(cfg["layers"], precision, cuda)
    out = run(m, inp, cfg["mode"], cuda)
    assert out.shape == d["eps"].shape and out.dtype == torch.float32
    err = valid_rel(out, d["eps"], cfg["lengths"])
    print(f"{name} {precision}: valid-region rel-L2 {err:.3e}, full {rel(out, d['eps']):.3e}")
    assert err < TOL[precision], err
    assert rel(out, d["eps"]) < 2 * TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_full_depth_against_oracle(cuda, precision):
    """8 layers, T=196, mixed per-person lengths (BASELINE config 5 semantics incl. the query-side mask quirk)."""
    import denoiser_oracle as DO
    import weights
    m, sd = build(8, precision, cuda)
    inp = weights.make_inputs(31, 6, 196, n_text=77, lengths=[196, 150, 60, 120, 196, 33])
    with torch.no_grad():
        ref = DO.denoiser_forward(sd, inp["x"], inp["t"], inp["length"], inp["xf_proj"], inp["xf_out"])
    out = run(m, inp, "text", cuda)
    err = valid_rel(out, ref, inp["length"])
    print(f"full depth {precision}: rel-L2 {err:.3e}")
    assert err < TOL[precision], err


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_c2_shape_against_oracle(cuda, precision):
    """The bench shape itself (BASELINE configs[1]: 64 pairs = 128 sequences x 196 frames, 8 layers, 77 text tokens) with
    mixed per-person lengths, one denoiser forward against the CPU oracle (one oracle forward at this size is ~2 s of CPU)."""
    import denoiser_oracle as DO
    import weights
    S, T = 128, 196
    m, sd = build(8, precision, cuda)
    rs = np.random.RandomState(5)
    lens = [int(v) for v in rs.randint(20, 197, S)]
    lens[0], lens[64], lens[1], lens[65] = 196, 196, 196, 21      # a full pair and a very uneven one
    inp = weights.make_inputs(77, S, T, n_text=77, lengths=lens)
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        ref = DO.denoiser_forward(sd, inp["x"], inp["t"], inp["length"], inp["xf_proj"], inp["xf_out"])
    out = run(m, inp, "text", cuda)
    err = valid_rel(out, ref, inp["length"])
    worst = max(valid_rel(out[s:s + 1], ref[s:s + 1], inp["length"][s:s + 1]) for s in range(S))
    print(f"C2 shape {precision}: rel-L2 {err:.3e}, worst sequence {worst:.3e}")
    assert err < TOL[precision], err
    assert worst < 3 * TOL[precision], worst


def test_heavy_tailed_residual_stream_bf16(cuda):
    """What trained checkpoints do to the residual stream and random-init goldens do not: outlier channels of 1e2-1e3 (a few
    output channels of every block's out-projection carry a large bias, accumulating over the 32 residual adds) and rows
    whose mean is >= 50 standard deviations away from zero (a common offset on a few positional rows).  The product path
    stores the stream in fp16 (saturating) and takes LayerNorm statistics as E[x^2] - mu^2 from fp32 partials: the output
    must stay within the bf16 tolerance of the fp32 oracle and NO store may saturate (hig_debug_saturation)."""
    import denoiser_oracle as DO
    import weights
    from hig_b200 import ops
    L, S, T = 8, 8, 64
    m, sd = build(L, "bf16", cuda)
    sd = {k: v.clone() for k, v in sd.items()}
    g = torch.Generator().manual_seed(3)
    hot = torch.randperm(512, generator=g)[:6]
    for k in sd:
        if k.endswith("proj_out.out_layers.2.bias"):
            sd[k][hot] += 25.0 * torch.sign(torch.randn(6, generator=g))      # ~ +-800 after 32 blocks on 6 channels
    sd["sequence_embedding"][5:9] += 60.0                                       # rows with |mu| / sigma ~ 60 at the first LayerNorms
    m.load_state_dict(sd, strict=True)
    inp = weights.make_inputs(9, S, T, n_text=77, lengths=[64, 40, 64, 17, 64, 40, 64, 17])
    with torch.no_grad():
        ref = DO.denoiser_forward(sd, inp["x"], inp["t"], inp["length"], inp["xf_proj"], inp["xf_out"])
    counter = torch.zeros(1, device=cuda, dtype=torch.int64)
    ops.debug_saturation(counter)
    try:
        out = run(m, inp, "text", cuda)
        torch.cuda.synchronize()
    finally:
        ops.debug_saturation(None)
    eng = m.engine()
    stream = eng.workspace(S, T)["xres"].float()
    mu, sg = stream.mean(1), stream.std(1)
    print(f"heavy tail: stream |max| {stream.abs().max().item():.0f}, max |mu|/sigma {(mu.abs() / sg).max().item():.1f}, "
          f"saturating stores {int(counter.item())}")
    assert stream.abs().max().item() > 300            # the outlier channels really are there
    assert int(counter.item()) == 0
    err = valid_rel(out, ref, inp["length"])
    print(f"heavy tail bf16: rel-L2 {err:.3e}")
    assert err < 1e-2, err
    # (B) rows whose mean sits >= 50 standard deviations from zero, without outlier channels to inflate sigma: a common
    # offset of 60 on four positional rows of an otherwise O(1) stream.  LayerNorm is shift invariant, the fp16 stream is
    # not: its ulp at 60 is 0.03, i.e. ~3 % of those rows' sigma per rounding — the error is confined to those rows.
    sdb = {k: v.clone() for k, v in weights.make_state_dict(seed=0, num_layers=L).items()}
    sdb["sequence_embedding"][5:9] += 60.0
    m.load_state_dict(sdb, strict=True)
    with torch.no_grad():
        ref_b = DO.denoiser_forward(sdb, inp["x"], inp["t"], inp["length"], inp["xf_proj"], inp["xf_out"])
    out_b = run(m, inp, "text", cuda)
    torch.cuda.synchronize()
    stream_b = eng.workspace(S, T)["xres"].float()
    ratio = (stream_b.mean(1).abs() / stream_b.std(1)).view(S, T)
    err_b = valid_rel(out_b, ref_b, inp["length"])
    rows_b = valid_rel(out_b[:, 6:10], ref_b[:, 6:10], inp["length"].clamp(max=4))
    print(f"heavy tail (B): max |mu|/sigma of the final stream {ratio.max().item():.1f} (rows 6..9: {ratio[:, 6:10].mean().item():.1f}), "
          f"rel-L2 {err_b:.3e}, on the offset rows {rows_b:.3e}")
    assert ratio.max().item() >= 20
    assert err_b < 1e-2, err_b
    # and the counter does count: push the stream over the fp16 range
    sd2 = {k: v.clone() for k, v in sd.items()}
    for k in sd2:
        if k.endswith("proj_out.out_layers.2.bias"):
            sd2[k][hot] *= 200.0
    m.load_state_dict(sd2, strict=True)
    counter.zero_()
    ops.debug_saturation(counter)
    try:
        run(m, inp, "text", cuda)
        torch.cuda.synchronize()
    finally:
        ops.debug_saturation(None)
    assert int(counter.item()) > 0


def test_gelu_erf_option_matches_exact_gelu(cuda, monkeypatch):
    """HIG_GELU=erf: the FFN epilogue computes the erf-form GELU (torch.nn.GELU's default, :257) instead of the tanh form."""
    import math
    from hig_b200 import ops
    M, N, K = 512, 1024, 512
    g = torch.Generator(device=cuda).manual_seed(1)
    x = torch.randn(M, K, device=cuda, generator=g).half()
    w = (torch.randn(N, K, device=cuda, generator=g) * 3 / math.sqrt(K)).half()
    b = torch.randn(N, device=cuda, generator=g)
    pre = x.float() @ w.float().t() + b
    exact, tanh = torch.nn.functional.gelu(pre), torch.nn.functional.gelu(pre, approximate="tanh")
    out = {}
    for mode in ("tanh", "erf"):
        monkeypatch.setenv("HIG_GELU", mode)
        o = torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
        ops.gemm_stream(ops.GS_BF16_GELU, x, w, b, o)
        out[mode] = o.float()
    assert (out["erf"] - exact.bfloat16().float()).abs().max() <= (out["tanh"] - exact.bfloat16().float()).abs().max()
    assert rel(out["erf"], exact) < 3e-3 and rel(out["tanh"], tanh) < 3e-3
    # on the region where the two formulas differ most (|v| ~ 2) the erf epilogue tracks the exact GELU, not the tanh form
    sel = (pre.abs() > 1.5) & (pre.abs() < 2.5)
    assert (out["erf"] - exact)[sel].abs().mean() < (out["erf"] - tanh)[sel].abs().mean()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sampling_loop_golden_and_graph(cuda, precision):
    """50-step chain (BASELINE config 1 schedule) with injected noise: fp32 mode against the real reference's final
    sample; graph replay must be bit-identical to eager stepping in both precisions."""
    import weights
    from hig_b200.gaussian_diffusion import (GaussianDiffusion, LossType, ModelMeanType, ModelVarType,
                                             get_named_beta_schedule)
    d, cfg, inp = load_case("loop")
    m, _ = build(cfg["layers"], precision, cuda)
    m.cap_id = False
    diff = GaussianDiffusion(betas=get_named_beta_schedule("linear", cfg["steps"]),
                             model_mean_type=ModelMeanType.EPSILON, model_var_type=ModelVarType.FIXED_SMALL,
                             loss_type=LossType.MSE)
    noise = weights.make_noise(cfg["seed"] + 100, cfg["steps"], cfg["S"], cfg["T"]).to(cuda)
    kw = {"xf_proj": inp["xf_proj"].to(cuda), "xf_out": inp["xf_out"].to(cuda), "length": inp["length"].to(cuda)}
    shape = (cfg["S"], cfg["T"], 263)
    a = diff.p_sample_loop(m, shape, noise=noise[0], clip_denoised=False, model_kwargs=kw, noise_seq=noise[1:],
                           use_graph=True)
    b = diff.p_sample_loop(m, shape, noise=noise[0], clip_denoised=False, model_kwargs=kw, noise_seq=noise[1:],
                           use_graph=False)
    c = diff.p_sample_loop(m, shape, noise=noise[0], clip_denoised=False, model_kwargs=kw, noise_seq=noise[1:],
                           use_graph=True)  # cached graph
    assert torch.equal(a, b) and torch.equal(a, c)
    err = rel(a, d["final"])
    print(f"loop {precision}: final-sample rel-L2 vs reference {err:.3e}")
    # 50 steps through an untrained eps-network are expansive (|x| reaches 1e4).  Measured on B200: fp32
    # mode 3.3e-7, bf16 2.9e-3 .. 4.0e-3 relative; gates 1e-4 and 5e-2 (per-step parity is the primary bf16 gate)
    assert err < (1e-4 if precision == "fp32" else 5e-2), err
    if precision == "fp32":
        # final sampled joints: the chain's |x| reaches 1e4, so the comparison is made on the sample rescaled to unit RMS
        sc = 1.0 / float(np.sqrt((d["final"].astype(np.float64) ** 2).mean()))
        mp_err, scale = joints_mpjpe(a * sc, torch.from_numpy(d["final"]) * sc, inp["length"])
        print(f"loop fp32: final joints MPJPE {mp_err:.3e} (mean |joint| {scale:.2f})")
        assert mp_err < MPJPE_TOL["fp32"], mp_err
    # generic per-step API (public model call + fused posterior kernel) agrees with the fast path
    img = noise[0].clone()
    for k, i in enumerate(range(cfg["steps"] - 1, -1, -1)):
        t = torch.full((cfg["S"],), i, device=cuda, dtype=torch.long)
        img = diff.p_sample(m, img, t, clip_denoised=False, model_kwargs=kw, noise=noise[1 + k])["sample"]
    assert rel(img, a) < (1e-6 if precision == "fp32" else 1e-6)


def test_branched_sampling_matches_single_branch(cuda, monkeypatch):
    """Production sampling splits the pairs over concurrent graph branches (HIG_BRANCHES, default 2).  With the
    posterior noise switched off (sigma table zeroed) the chain is deterministic, and because no kernel mixes rows of
    different sequences except through the pair, every branch count must give the same samples — also with ragged
    lengths, per-sequence text and a second call that reuses the cached graph with new inputs."""
    import weights
    from hig_b200.gaussian_diffusion import (GaussianDiffusion, LossType, ModelMeanType, ModelVarType,
                                             get_named_beta_schedule)
    m, _ = build(2, "bf16", cuda)
    m.cap_id = False
    S, T, steps = 32, 40, 50
    lens = [40, 13, 27, 31, 40, 5, 40, 22] * 4
    res = {}
    for nb in ("1", "2"):
        monkeypatch.setenv("HIG_BRANCHES", nb)
        diff = GaussianDiffusion(betas=get_named_beta_schedule("linear", steps), model_mean_type=ModelMeanType.EPSILON,
                                 model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
        diff._tables(cuda)["coef"][4].zero_()
        outs = []
        for seed in (21, 22):
            inp = weights.make_inputs(seed, S, T, n_text=7, lengths=lens)
            kw = {"xf_proj": inp["xf_proj"].to(cuda), "xf_out": inp["xf_out"].to(cuda), "length": inp["length"].to(cuda)}
            outs.append(diff.p_sample_loop(m, (S, T, 263), noise=inp["x"].to(cuda), clip_denoised=False, model_kwargs=kw))
        res[nb] = outs
        assert diff.last_launches > 0
    for a, b in zip(res["1"], res["2"]):
        assert torch.isfinite(a).all() and not torch.equal(res["1"][0], res["1"][1])
        print(f"branches 1 vs 2: bit-identical={torch.equal(a, b)} rel={rel(a, b):.2e}")
        assert rel(a, b) < 1e-6


def test_teacher_forced_steps_bf16(cuda):
    """Primary bf16 gate: feed the ORACLE's x_t to the CUDA path at several t and compare eps and x_{t-1}."""
    import denoiser_oracle as DO
    import diffusion_oracle as DF
    import weights
    from hig_b200 import ops
    from hig_b200.gaussian_diffusion import (GaussianDiffusion, LossType, ModelMeanType, ModelVarType,
                                             get_named_beta_schedule)
    m, sd = build(4, "bf16", cuda)
    m.cap_id = False
    S, T = 4, 64
    inp = weights.make_inputs(77, S, T, n_text=77, lengths=[64, 50, 64, 50])
    sch = DF.Schedule(1000)
    diff = GaussianDiffusion(betas=get_named_beta_schedule("linear", 1000), model_mean_type=ModelMeanType.EPSILON,
                             model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
    z = weights.make_noise(5, 0, S, T)[0]
    for tv in (999, 500, 1, 0):
        t = torch.full((S,), tv, dtype=torch.long)
        with torch.no_grad():
            eps_ref = DO.denoiser_forward(sd, inp["x"], t, inp["length"], inp["xf_proj"], inp["xf_out"])
            x_ref = DF.p_sample_step(sch, inp["x"], eps_ref, t, z)
            out = diff.p_sample(m, inp["x"].to(cuda), t.to(cuda), clip_denoised=False,
                                model_kwargs={"xf_proj": inp["xf_proj"].to(cuda), "xf_out": inp["xf_out"].to(cuda),
                                              "length": inp["length"].to(cuda)}, noise=z.to(cuda))
            eps = m(inp["x"].to(cuda), t.to(cuda), length=inp["length"].to(cuda), xf_proj=inp["xf_proj"].to(cuda),
                    xf_out=inp["xf_out"].to(cuda))
        e1, e2 = valid_rel(eps, eps_ref, inp["length"]), valid_rel(out["sample"], x_ref, inp["length"])
        mp_err, scale = joints_mpjpe(out["sample"], x_ref, inp["length"])
        print(f"t={tv}: eps rel {e1:.3e}  x_prev rel {e2:.3e}  joints MPJPE {mp_err:.3e} (mean |joint| {scale:.2f})")
        assert e1 < 1e-2 and e2 < 1e-2
        assert mp_err < MPJPE_TOL["bf16"]


# ------------------------------------------------------------------------------------------ properties at BASELINE sizes
@pytest.fixture(scope="module")
def big(cuda):
    import weights
    m, sd = build(8, "bf16", cuda)
    m.cap_id = False
    S, T = 128, 196   # BASELINE config 2 shape: 64 pairs, 196 frames
    rs = np.random.RandomState(3)
    pair_len = rs.randint(20, 197, size=S // 2)
    pair_len[:4] = 196
    inp = weights.make_inputs(41, S, T, n_text=77, lengths=np.concatenate([pair_len, pair_len]))
    inp = {k: v.to(cuda) for k, v in inp.items()}
    with torch.no_grad():
        base = m(inp["x"], inp["t"], length=inp["length"], xf_proj=inp["xf_proj"], xf_out=inp["xf_out"])
    return m, inp, base


def _mask(inp):
    T = inp["x"].shape[1]
    return torch.arange(T, device=inp["x"].device)[None] < inp["length"][:, None]


def test_big_finite_and_nontrivial(big):
    m, inp, base = big
    assert torch.isfinite(base).all()
    assert base.abs().mean().item() > 0.05


def test_big_pad_garbage_invariance(big):
    m, inp, base = big
    x2 = inp["x"].clone()
    x2[~_mask(inp)] = 1e3
    with torch.no_grad():
        out = m(x2, inp["t"], length=inp["length"], xf_proj=inp["xf_proj"], xf_out=inp["xf_out"])
    assert torch.equal(out[_mask(inp)], base[_mask(inp)])


def test_big_role_swap_equivariance(big):
    m, inp, base = big
    B = inp["x"].shape[0] // 2
    sw = lambda a: torch.cat([a[B:], a[:B]])
    with torch.no_grad():
        out = m(sw(inp["x"]), sw(inp["t"]), length=sw(inp["length"]), xf_proj=sw(inp["xf_proj"]),
                xf_out=sw(inp["xf_out"]))
    assert torch.equal(sw(out)[_mask(inp)], base[_mask(inp)])


def test_big_batch_composition_independence(big):
    m, inp, base = big
    B = inp["x"].shape[0] // 2
    idx = torch.tensor([3, 17, 40, B + 3, B + 17, B + 40], device=inp["x"].device)
    with torch.no_grad():
        out = m(inp["x"][idx], inp["t"][idx], length=inp["length"][idx], xf_proj=inp["xf_proj"][idx],
                xf_out=inp["xf_out"][idx].contiguous())
    mk = _mask(inp)[idx]
    assert rel(out[mk], base[idx][mk]) < 1e-6   # same kernels, same per-row arithmetic: only tile placement differs


def test_padding_vs_trimmed(cuda):
    """The reference's __main__ smoke (interaction_transformer.py:849-854): T=32,len=31 padded vs T=31 trimmed."""
    import weights
    m, _ = build(2, "bf16", cuda)
    m.cap_id = False
    inp = weights.make_inputs(8, 2, 32, n_text=77, lengths=[31, 31])
    g = lambda k: inp[k].to(cuda)
    with torch.no_grad():
        a = m(g("x"), g("t"), length=g("length"), xf_proj=g("xf_proj"), xf_out=g("xf_out"))
        b = m(g("x")[:, :31].contiguous(), g("t"), length=g("length"), xf_proj=g("xf_proj"), xf_out=g("xf_out"))
    assert rel(a[:, :31], b) < 1e-6


def test_zero_init_identity(cuda):
    import weights
    m, _ = build(2, "bf16", cuda, zero_init=True)
    m.cap_id = False
    inp = weights.make_inputs(9, 4, 40, n_text=77)
    g = lambda k: inp[k].to(cuda)
    with torch.no_grad():
        out = m(g("x"), g("t"), length=g("length"), xf_proj=g("xf_proj"), xf_out=g("xf_out"))
    assert out.abs().max().item() == 0.0


def test_length_forms_and_errors(cuda):
    import weights
    m, _ = build(1, "bf16", cuda)
    m.cap_id = False
    inp = weights.make_inputs(10, 4, 24, n_text=3, lengths=[24, 10, 24, 10])
    g = lambda k: inp[k].to(cuda)
    with torch.no_grad():
        a = m(g("x"), g("t"), length=g("length"), xf_proj=g("xf_proj"), xf_out=g("xf_out"))
        b = m(g("x"), g("t"), length=g("length").view(-1, 1), xf_proj=g("xf_proj"), xf_out=g("xf_out"))  # [2B,1]
        c = m(g("x"), g("t"), length=[24, 10, 24, 10], xf_proj=g("xf_proj"), xf_out=g("xf_out"))         # list
    assert torch.equal(a, b) and torch.equal(a, c)
    with pytest.raises(RuntimeError):
        m(inp["x"], inp["t"], length=inp["length"], xf_proj=inp["xf_proj"], xf_out=inp["xf_out"])  # CPU input
    with pytest.raises(ValueError):
        with torch.no_grad():
            m(g("x")[:3], g("t")[:3], length=g("length")[:3], xf_proj=g("xf_proj")[:3], xf_out=g("xf_out")[:3])


def test_bucketed_generation_and_joints_end_to_end(cuda):
    """caption ids -> length-bucketed sampling (ddp.generate_bucketed, single process) -> per-pair motions trimmed to
    their own lengths, in the caller's order; then generate_joints: caption -> joints without leaving the GPU."""
    import argparse
    from hig_b200.ddp import generate_bucketed
    from hig_b200.mul_ddpm_trainer import DDPMMulTrainer
    m, _ = build(1, "bf16", cuda, cap_id=True)
    opt = argparse.Namespace(device=cuda, multi=True, label_path=None, cap_id=True, diffusion_steps=50, is_train=False)
    tr = DDPMMulTrainer(opt, m)
    lens = torch.tensor([[40], [12], [33], [40], [7]])
    c1, c2 = [3, 5, 7, 9, 11], [4, 6, 8, 10, 12]
    out = generate_bucketed(tr, c1, c2, lens, 263, batch_size=2)
    assert [a.shape for a, _ in out] == [(n, 263) for n in (40, 12, 33, 40, 7)]
    assert all(a.is_cuda and torch.isfinite(a).all() and torch.isfinite(b).all() for a, b in out)
    mean, std, im, isd = joint_stats()
    js = tr.generate_joints(c1, c2, lens.view(-1), 263, mean=mean, std=std, init_mean=im, init_std=isd)
    assert [j1.shape for j1, _ in js] == [(n - 1, 22, 3) for n in (40, 12, 33, 40, 7)]
    assert all(torch.isfinite(j1).all() and torch.isfinite(j2).all() for j1, j2 in js)


def test_bucketed_generation_is_bit_identical_to_plain_generate(cuda):
    """ddp.generate_bucketed regroups the pairs by length (other batch compositions, other padded T) — with the posterior
    noise switched off (sigma table zeroed) and x_T given per pair, every pair's sample must equal plain
    DDPMMulTrainer.generate's on its valid frames BIT FOR BIT: no kernel mixes rows of different pairs, and no reduction
    order depends on the padded length or the batch size."""
    import argparse
    from hig_b200.ddp import generate_bucketed
    from hig_b200.mul_ddpm_trainer import DDPMMulTrainer
    m, _ = build(2, "bf16", cuda, cap_id=True)
    opt = argparse.Namespace(device=cuda, multi=True, label_path=None, cap_id=True, diffusion_steps=50, is_train=False)
    tr = DDPMMulTrainer(opt, m)
    tr.diffusion._tables(cuda)["coef"][4].zero_()

    class RowWise(torch.nn.Module):
        """text_proj evaluated one row at a time: torch's batched Linear (cuBLAS) may pick another kernel / summation order
        for another batch size, which would make xf_proj — an INPUT of the path under test — depend on the batching."""
        def __init__(self, lin):
            super().__init__()
            self.w, self.b = lin.weight.detach().double(), lin.bias.detach().double()

        def forward(self, e):
            return torch.stack([self.w @ r.double() + self.b for r in e]).float()

    m.text_proj = RowWise(m.text_proj[0])
    lens = torch.tensor([40, 12, 33, 40, 7, 25, 33, 18, 2])
    n = len(lens)
    c1, c2 = list(range(1, n + 1)), list(range(n + 1, 2 * n + 1))
    g = torch.Generator().manual_seed(5)
    x_T = torch.randn(n, 2, 40, 263, generator=g)
    plain = tr.generate(c1, c2, lens, 263, batch_size=512, pair_noise=x_T)
    plain_small = tr.generate(c1, c2, lens, 263, batch_size=4, pair_noise=x_T)
    buck = generate_bucketed(tr, c1, c2, lens, 263, batch_size=3, pair_noise=x_T)
    for i in range(n):
        L = int(lens[i])
        for j in (0, 1):
            assert buck[i][j].shape == (L, 263)
            assert torch.equal(buck[i][j], plain[i][j][:L]), (i, j, rel(buck[i][j], plain[i][j][:L]))
            assert torch.equal(plain_small[i][j][:L], plain[i][j][:L]), (i, j)
    assert not torch.equal(plain[0][0], plain[3][0])          # same length, different captions / x_T: different samples


def test_text_state_cache_not_fooled_by_recycled_memory(cuda):
    """The step-invariant text state (A_text) is cached per xf_out tensor.  A second batch whose xf_out lands on the
    recycled address of the first (the caching allocator hands the freed block straight back) must not hit that cache."""
    import weights
    m, _ = build(1, "bf16", cuda)
    m.cap_id = False
    inp = weights.make_inputs(31, 4, 24, n_text=77, lengths=[24, 10, 24, 10])
    g = lambda k: inp[k].to(cuda)
    x, t, ln, xp = g("x"), g("t"), g("length"), g("xf_proj")
    with torch.no_grad():
        a = g("xf_out")
        ptr_a = a.data_ptr()
        ya = m(x, t, length=ln, xf_proj=xp, xf_out=a)
        del a
        b = (inp["xf_out"] * -1.5 + 0.3).to(cuda)          # different text, same shape: usually the same address
        same_address = b.data_ptr() == ptr_a
        yb = m(x, t, length=ln, xf_proj=xp, xf_out=b)
        m.engine()._text_cache = None
        yb_fresh = m(x, t, length=ln, xf_proj=xp, xf_out=b)
    print(f"xf_out address recycled: {same_address}")
    assert torch.equal(yb, yb_fresh) and not torch.equal(ya, yb)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_c1_full_depth_50_step_sample_against_reference_golden(cuda, precision):
    """BASELINE config 1 on the CUDA path: 8 layers, one pair, 196 frames, 50-step schedule, injected noise, against the
    real reference's final sample (tests/golden/c1.npz), |x| ~ 2e4 after 50 steps through the untrained network.
    Measured on B200: fp32 mode 3.8e-7, bf16 3.1e-3 relative; gates 1e-4 and 5e-2."""
    import weights
    from hig_b200.gaussian_diffusion import (GaussianDiffusion, LossType, ModelMeanType, ModelVarType,
                                             get_named_beta_schedule)
    d, cfg, inp = load_case("c1")
    m, _ = build(cfg["layers"], precision, cuda)
    m.cap_id = False
    diff = GaussianDiffusion(betas=get_named_beta_schedule("linear", cfg["steps"]),
                             model_mean_type=ModelMeanType.EPSILON, model_var_type=ModelVarType.FIXED_SMALL,
                             loss_type=LossType.MSE)
    noise = weights.make_noise(cfg["seed"] + 100, cfg["steps"], cfg["S"], cfg["T"]).to(cuda)
    kw = {"xf_proj": inp["xf_proj"].to(cuda), "xf_out": inp["xf_out"].to(cuda), "length": inp["length"].to(cuda)}
    out = diff.p_sample_loop(m, (cfg["S"], cfg["T"], 263), noise=noise[0], clip_denoised=False, model_kwargs=kw,
                             noise_seq=noise[1:])
    err = rel(out, d["final"])
    print(f"C1 {precision}: final-sample rel-L2 vs reference {err:.3e}")
    assert torch.isfinite(out).all()
    assert err < (1e-4 if precision == "fp32" else 5e-2), err
