"""ms per denoiser+posterior step at the C2 shape via CUDA-graph replay (the production path), 200 steps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda:0")
model = bench.build_model(dev)
bench.CFG["diffusion_steps"] = n
trainer = bench.make_trainer(model, dev, n)
S, T, C = 128, 196, 263
g = torch.Generator(device=dev).manual_seed(0)
kw = {"xf_proj": torch.randn(S, 2048, device=dev, generator=g) * 0.5,
      "xf_out": torch.randn(S, 77, 256, device=dev, generator=g),
      "length": torch.full((S,), T, device=dev, dtype=torch.long)}
x_T = torch.randn(S, T, C, device=dev, generator=g)
for i in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = trainer.diffusion.p_sample_loop(model, (S, T, C), noise=x_T, clip_denoised=False, model_kwargs=kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"run {i}: {ms / n * 1e3:.1f} us/step  ({ms:.1f} ms for {n} steps)  finite={bool(torch.isfinite(out).all())}")
f = bench.flops_per_denoiser_step(S, T)
print(f"roofline frac (sustained 1377.6 TF): {f / (ms / n * 1e-3) / 1e12 / 1377.6:.3f}")
