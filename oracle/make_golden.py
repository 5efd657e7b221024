"""TEST INFRASTRUCTURE — generate tests/golden/*.npz by running the REAL reference (imported read-only from
/root/reference/codes behind oracle/ref_shims.py) on seeded weights/inputs from oracle/weights.py.

Run in the build container only:   python oracle/make_golden.py [case ...]   (no argument = every case)
The fixtures are small (inputs are re-derived from seeds at test time; only outputs are stored).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402
import weights  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = {
    # name: model cfg, input cfg
    "fwd_cap": dict(layers=2, S=4, T=24, seed=11, lengths=[24, 9, 24, 9], timesteps=[999, 3, 999, 3], mode="cap"),
    "fwd_text": dict(layers=2, S=6, T=40, seed=12, lengths=[40, 13, 27, 31, 40, 5], timesteps=None, mode="text",
                     n_text=77),
    "fwd_text_full": dict(layers=1, S=2, T=196, seed=13, lengths=[196, 120], timesteps=[500, 500], mode="text",
                          n_text=77),
    "loop": dict(layers=2, S=2, T=16, seed=14, lengths=[16, 11], mode="text", n_text=5, steps=50),
    "train": dict(layers=2, S=4, T=12, seed=15, lengths=[12, 7, 12, 7], mode="text", n_text=3),
    # BASELINE config 1 exactly: full-depth denoiser, one pair, 196 frames, the 50-step schedule (CPU plumbing case)
    "c1": dict(layers=8, S=2, T=196, seed=17, lengths=[196, 196], mode="text", n_text=77, steps=50),
    # sample -> joints post-processing (tools/visualization.py:149-155 + utils/motion_process.recover_from_ric2)
    "joints": dict(S=6, T=24, seed=16),
}


def joints_golden(c):
    """Runs the reference's recover_from_ric2 on seeded 'samples' de-normalised exactly as tools/visualization.py
    does (:149-155 is inline script code, restated here statement by statement around the imported function)."""
    import importlib
    ref_shims.import_reference()
    mp = importlib.import_module("utils.motion_process")
    x, mean, std, init_mean, init_std = weights.make_joint_inputs(c["seed"], c["S"], c["T"])
    B = c["S"] // 2
    j1, j2 = [], []
    for i in range(B):
        motion1, motion2 = x[i].numpy().copy(), x[i + B].numpy().copy()
        motion1[1:] = motion1[1:] * std + mean
        motion2[1:] = motion2[1:] * std + mean
        motion1[0, :4] = motion1[0, :4] * init_std + init_mean
        motion2[0, :4] = motion2[0, :4] * init_std + init_mean
        motion1 = np.concatenate([motion1[1:], motion1[0][None, :]], axis=0)
        motion2 = np.concatenate([motion2[1:], motion2[0][None, :]], axis=0)
        a, b = mp.recover_from_ric2(torch.from_numpy(motion1).unsqueeze(0).float(),
                                    torch.from_numpy(motion2).unsqueeze(0).float(), 22)
        j1.append(a.squeeze(0).numpy())
        j2.append(b.squeeze(0).numpy())
    return {"joints": np.stack(j1 + j2)}      # [S, T-1, 22, 3], persons stacked like the sampler's batch


def build_reference_model(it, layers, wseed=0):
    m = it.MotionInteractionTransformer(263, num_frames=196, num_layers=layers, latent_dim=512, cap_id=True)
    sd = weights.make_state_dict(seed=wseed, num_layers=layers)
    m.load_state_dict(sd, strict=True)
    m.eval()
    return m, sd


def ref_forward(m, inp, mode):
    if mode == "cap":
        m.cap_id = True
        return m(inp["x"], inp["t"], length=inp["length"], text=[inp["cap1"], inp["cap2"]])
    m.cap_id = False  # take the sampling branch that receives precomputed xf_proj / xf_out (:585-589)
    return m(inp["x"], inp["t"], length=inp["length"], xf_proj=inp["xf_proj"], xf_out=inp["xf_out"])


def main():
    it, gd = ref_shims.import_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(8)
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    for name, c in CASES.items():
        if only and name not in only:
            continue
        if name == "joints":
            out = {"cfg": np.array(repr(c)), **joints_golden(c)}
            np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
            print(name, {k: getattr(v, "shape", None) for k, v in out.items()})
            continue
        m, sd = build_reference_model(it, c["layers"])
        inp = weights.make_inputs(c["seed"], c["S"], c["T"], n_text=c.get("n_text", 1), lengths=c["lengths"],
                                  timesteps=c.get("timesteps"))
        out = {"cfg": np.array(repr(c))}
        if name.startswith("fwd"):
            with torch.no_grad():
                out["eps"] = ref_forward(m, inp, c["mode"]).numpy()
        elif name in ("loop", "c1"):
            steps = c["steps"]
            noise = weights.make_noise(c["seed"] + 100, steps, c["S"], c["T"])
            diff = gd.GaussianDiffusion(betas=gd.get_named_beta_schedule("linear", steps),
                                        model_mean_type=gd.ModelMeanType.EPSILON,
                                        model_var_type=gd.ModelVarType.FIXED_SMALL, loss_type=gd.LossType.MSE)
            # the reference draws th.randn_like inside p_sample (:657); feed it our pre-generated noise
            seq = iter([noise[1 + k] for k in range(steps)])
            orig = gd.th.randn_like
            gd.th.randn_like = lambda x: next(seq)
            try:
                m.cap_id = False
                final = diff.p_sample_loop(m, (c["S"], c["T"], 263), noise=noise[0].clone(), clip_denoised=False,
                                           model_kwargs={"xf_proj": inp["xf_proj"], "xf_out": inp["xf_out"],
                                                         "length": inp["length"]})
            finally:
                gd.th.randn_like = orig
            out["final"] = final.numpy()
        elif name == "train":
            tr = ref_shims.import_reference_trainer()
            steps = 1000
            diff = gd.GaussianDiffusion(betas=gd.get_named_beta_schedule("linear", steps),
                                        model_mean_type=gd.ModelMeanType.EPSILON,
                                        model_var_type=gd.ModelVarType.FIXED_SMALL, loss_type=gd.LossType.MSE)
            m.cap_id = False
            m.train()
            noise = weights.make_noise(c["seed"] + 100, 0, c["S"], c["T"])[0]
            terms = diff.training_losses(m, inp["x"], inp["t"], noise=noise,
                                         model_kwargs={"xf_proj": inp["xf_proj"], "xf_out": inp["xf_out"],
                                                       "length": inp["length"]})
            out["x_t"] = diff.q_sample(inp["x"], inp["t"], noise=noise).numpy()
            out["pred"] = terms["pred"].detach().numpy()
            trainer = tr.DDPMMulTrainer.__new__(tr.DDPMMulTrainer)
            trainer.with_label = True
            trainer.encoder = types.SimpleNamespace(module=types.SimpleNamespace(two_embed=True))
            trainer.mse_criterion = torch.nn.MSELoss(reduction="none")
            trainer.fake_noise, trainer.real_noise = terms["pred"], terms["target"]
            trainer.src_mask = m.generate_src_mask(c["T"], inp["length"])
            trainer.backward_G()
            out["loss_label"] = trainer.loss_mot_rec.detach().numpy()
            trainer.loss_mot_rec.backward()
            out["grad_norms"] = np.array([float(p.grad.norm()) if p.grad is not None else -1.0
                                          for _, p in sorted(m.named_parameters())], dtype=np.float64)
            out["grad_names"] = np.array([n for n, _ in sorted(m.named_parameters())])
            named = dict(m.named_parameters())
            for g in ("out.bias", "joint_embed2.weight", "temporal_decoder_blocks.0.sa_block.norm.weight",
                      "temporal_decoder_blocks.1.ffn.linear1.bias"):
                out["grad:" + g] = named[g].grad.numpy().copy()
            # PIT branch (no labels): 4B stacked sequences, min over the two caption assignments
            trainer.with_label = False
            pit_pred = torch.cat([terms["pred"].detach(), terms["pred"].detach().flip(0)])
            pit_tgt = torch.cat([terms["target"], terms["target"]])
            trainer.fake_noise, trainer.real_noise = pit_pred, pit_tgt
            trainer.src_mask = torch.cat([trainer.src_mask, trainer.src_mask])
            trainer.backward_G()
            out["loss_pit"] = trainer.loss_mot_rec.detach().numpy()
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
        print(name, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
