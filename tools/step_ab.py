"""Interleaved A/B of env-selected variants on the production sampling path (C2 shape, CUDA-graph replay).
usage: step_ab.py <steps> <reps> VAR=a,b [VAR2=c,d ...]   — every combination is run <reps> times, round-robin, so that
clock / power drift hits all variants alike; prints us/step per run with the SM clock read right after it, then
min / median per variant."""
import itertools
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

n = int(sys.argv[1])
reps = int(sys.argv[2])
axes = [(a.split("=")[0], a.split("=")[1].split(",")) for a in sys.argv[3:]]
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


def clock():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                              capture_output=True, text=True).stdout.strip()
    except Exception:
        return "?"


def run():
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = trainer.diffusion.p_sample_loop(model, (S, T, C), noise=x_T, clip_denoised=False, model_kwargs=kw)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, bool(torch.isfinite(out).all())


combos = list(itertools.product(*[v for _, v in axes]))
res = {c: [] for c in combos}
run(); run()   # lazy initialisation, clocks up
for r in range(reps):
    for c in combos:
        for (k, _), v in zip(axes, c):
            os.environ[k] = v
        us, ok = run()
        res[c].append(us)
        print(f"rep {r} {dict(zip([k for k, _ in axes], c))}: {us:8.1f} us/step finite={ok} clk/power={clock()}", flush=True)
for c in combos:
    v = sorted(res[c])
    print(f"{dict(zip([k for k, _ in axes], c))}: min {v[0]:.1f}  median {v[len(v) // 2]:.1f}  max {v[-1]:.1f} us/step")
