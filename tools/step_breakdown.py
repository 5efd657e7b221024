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
# classes of the PRODUCT schedule (DenoiserEngine.layers_stream): every call of a class is replaced by a no-op at capture
for label, dis in [("without the token-sized projections (hig_gemm_stream)", ("gemm_stream",)),
                   ("without the general GEMM (FiLM linears, M = S)", ("gemm",)),
                   ("without the K/V half (attn_kv / eff_attn)", ("attn_kv", "eff_attn")),
                   ("without attention-apply + stylize", ("attn_apply_stylize", "attn_apply_stylize_tc")),
                   ("without ln_film_silu (FFN branch)", ("ln_film_silu",)),
                   ("without time table / tile_rows", ("time_table_silu", "tile_rows", "timestep_embed")),
                   ("only the projections", ("gemm", "attn_kv", "eff_attn", "attn_apply_stylize", "attn_apply_stylize_tc",
                                             "ln_film_silu", "time_table_silu", "tile_rows", "timestep_embed"))]:
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
    print(f"{label:46s} {best:8.1f} us")


from hig_b200.gaussian_diffusion import GaussianDiffusion, LossType, ModelMeanType, ModelVarType, get_named_beta_schedule  # noqa: E402
diff = GaussianDiffusion(betas=get_named_beta_schedule("linear", 1000), model_mean_type=ModelMeanType.EPSILON,
                         model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
coef = diff._tables(dev)["coef"]
W = eng.packed()
D = 512
tok = S * T
qkv, xres, sact, stats = ws["qkv"], ws["xres"], ws["sact"], ws["stats"]
q_ca = qkv.view(-1)[:tok * D].view(tok, D)
gs = ops.gemm_stream
ln_kind = ops.GS_LN_QSM if eng.qsm else ops.GS_LN_BF16
# single kernels of the product path, replayed back to back (operands stay L2-warm: compare with the in-graph classes above)
time_fn("embed (time MLP + 32 FiLM linears)", lambda: eng.embed(ws, t, xf_proj, S))
time_fn("embed_motion (tile_rows + in-place projection)", lambda: eng.embed_motion(ws, T))
time_fn("heads (out / out2, fp16 eps)", lambda: eng.heads(ws, S, T))
time_fn("ddpm_step + pack + t--", lambda: ops.ddpm_step(x, ws["eps16"] if eng.heads16 else ws["eps"], t, coef, noise=None, seed=1,
                                                        packed=ws["xa"], t_next=None))
time_fn("Q|K|V projection (LN folded, query softmax)", lambda: gs(ln_kind, xres, W["l0.sa.qkv.wg"], W["l0.sa.qkv.bg"], qkv,
                                                                  wsum=W["l0.sa.qkv.wsum"], stats_in=stats, ln_width=D))
time_fn("Q projection (text CA)", lambda: gs(ln_kind, xres, W["l0.ca.q.wg"], W["l0.ca.q.bg"], q_ca, wsum=W["l0.ca.q.wsum"],
                                             stats_in=stats, ln_width=D))
time_fn("out-projection + fp16 residual + row stats", lambda: gs(ops.GS_RES_H, sact, W["l0.sa.po.w"], W["l0.sa.po.b"], xres,
                                                                 stats_out=stats))
time_fn("FFN linear1 + GELU", lambda: gs(ops.GS_BF16_GELU, xres, W["l0.ffn.w1h"], W["l0.ffn.b1"], ws["g"]))
time_fn("FFN linear2 (streamed, K = 1024)", lambda: gs(ops.GS_BF16, ws["g"], W["l0.ffn.w2"], W["l0.ffn.b2"], ws["y"]))
time_fn("FFN LayerNorm + FiLM + SiLU", lambda: ops.ln_film_silu(ws["y"], W["l0.ffn.po.ln.w"], W["l0.ffn.po.ln.b"], sact, rows_per_seq=T,
                                                                scale_shift=eng._ss(ws, W, "l0.ffn"), silu=True))
time_fn("K/V half (A or A^T)", lambda: ops.attn_kv(qkv[:, D:2 * D], qkv[:, 2 * D:], ws["a_blk"], S, T, 8, length=ws["len"],
                                                   transposed=eng.apply_tc))
time_fn("attention apply + stylize (product kernel)", lambda: eng._attend(ws, W, "l0.ca", S, T, q_ca, a_in=ws["a_blk"],
                                                                           q_softmaxed=eng.qsm))
