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
S, T, C = int(os.environ.get("BD_S", "128")), 196, 263   # BD_S: sequences (C2 = 128; 1024 = the C3 batch on one GPU)
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


def time_fn(label, fn):
    fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
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
    print(f"{label:34s} {best:8.1f} us")


from hig_b200.gaussian_diffusion import GaussianDiffusion, LossType, ModelMeanType, ModelVarType, get_named_beta_schedule  # noqa: E402
diff = GaussianDiffusion(betas=get_named_beta_schedule("linear", 1000), model_mean_type=ModelMeanType.EPSILON,
                         model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
coef = diff._tables(dev)["coef"]
W = eng.packed()
time_fn("embed (time MLP + 32 FiLM linears)", lambda: eng.embed(ws, t, xf_proj, S))
time_fn("embed_motion", lambda: eng.embed_motion(ws, T))
time_fn("heads (out / out2)", lambda: eng.heads(ws, S, T))
time_fn("ddpm_step + pack + t--", lambda: ops.ddpm_step(x, ws["eps16"] if eng.heads16 else ws["eps"], t, coef, noise=None, seed=1, packed=ws["xa"], t_next=None))
D = 512
qkv = ws["qkv"]
time_fn("1 x qkv GEMM", lambda: eng._gemm(ws["n"], W["l0.sa.qkv.w"], W["l0.sa.qkv.b"], out=qkv))
time_fn("1 x q GEMM", lambda: eng._gemm(ws["n"], W["l0.ca.q.w"], W["l0.ca.q.b"], out=qkv.view(-1)[:S * T * D].view(S * T, D)))
time_fn("1 x ffn1 GEMM (+GELU)", lambda: eng._gemm(ws["xb"], W["l0.ffn.w1"], W["l0.ffn.b1"], out=ws["g"], act=ops.ACT_GELU))
time_fn("1 x ffn2 GEMM", lambda: eng._gemm(ws["g"], W["l0.ffn.w2"], W["l0.ffn.b2"], out=ws["y"]))
time_fn("1 x out-proj GEMM (res fp16)", lambda: eng._project(ws, W, "l0.sa", False))
time_fn("1 x out-proj GEMM (res fp16 + xb)", lambda: eng._project(ws, W, "l0.ffn", True))
time_fn("1 x pre-LN (fp16 -> bf16)", lambda: ops.ln_film_silu(ws["xres"], W["l0.sa.ln.w"], W["l0.sa.ln.b"], ws["n"]))
time_fn("1 x attn K/V half", lambda: ops.eff_attn(ops.ATTN_KV_ONLY, S, T, 8, k=qkv[:, D:2 * D], v=qkv[:, 2 * D:], a_out=ws["a_blk"], length=ws["len"]))
time_fn("1 x attn apply + stylize", lambda: eng._attend(ws, W, "l0.ca", S, T, qkv[:, :D], a_in=ws["a_blk"]))
