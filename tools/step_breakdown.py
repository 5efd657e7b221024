"""In-graph cost of each kernel class at the C2 shape: capture the step graph with one class of C-ABI calls disabled
(the buffers keep whatever they held; timing only) and subtract from the full step.  Development aid."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from hig_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device("cuda:0")
model = bench.build_model(dev)
eng = model.engine()
S, T, C = 128, 196, 263
g = torch.Generator(device=dev).manual_seed(0)
xf_proj = torch.randn(S, 2048, device=dev, generator=g) * 0.5
xf_out = torch.randn(S, 77, 256, device=dev, generator=g)
x = torch.randn(S, T, C, device=dev, generator=g)
ws = eng.workspace(S, T)
eng.set_lengths(ws, None, S, T)
a_text = eng.text_state(xf_out)
t = torch.full((S,), 500, device=dev, dtype=torch.long)
ops.pack_motion(x, ws["xa"])


def time_graph(label, disabled=()):
    saved = {}
    for name in disabled:
        saved[name] = getattr(ops, name)
        setattr(ops, name, lambda *a, **k: None)
    try:
        def step():
            eng.run_packed(ws, t, xf_proj, a_text, S, T)
        step()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            step()
        for _ in range(5):
            gr.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                gr.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / n * 1e3)
    finally:
        for name, fn in saved.items():
            setattr(ops, name, fn)
    print(f"{label:34s} {best:8.1f} us/step")
    return best


full = time_graph("full denoiser step")
for label, dis in [("without GEMMs", ("gemm",)), ("without eff_attn (K/V half)", ("eff_attn",)),
                   ("without attn_apply_stylize", ("attn_apply_stylize",)), ("without ln_film_silu", ("ln_film_silu",)),
                   ("only GEMMs", ("eff_attn", "attn_apply_stylize", "ln_film_silu", "timestep_embed"))]:
    tt = time_graph(label, dis)
    print(f"    -> class cost {full - tt:8.1f} us ({100 * (full - tt) / full:4.1f}%)")
