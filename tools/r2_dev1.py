"""Round-2 dev check #1 (GPU): MN-major GEMM correctness + timing, tensor-core attention backward vs the CUDA-core one."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hig_b200  # noqa
from hig_b200 import ops

dev = torch.device("cuda:0")
def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

tok = 23296
for (M, N) in [(512, 512), (1536, 512), (1024, 512), (512, 1024)]:
    dy = torch.randn(tok, M, device=dev).bfloat16(); x = torch.randn(tok, N, device=dev).bfloat16()
    out = torch.zeros(M, N, device=dev)
    try:
        ops.gemm_t(dy, x, trans_a=True, trans_b=True, out_f32=out, split_k=-1)
        torch.cuda.synchronize()
        e = rel(out, dy.double().t() @ x.double())
        us = timeit(lambda: ops.gemm_t(dy, x, trans_a=True, trans_b=True, out_f32=out, split_k=-1))
        # old path: transposes + splitk
        dyT = torch.empty(M, tok, device=dev, dtype=torch.bfloat16); xT = torch.empty(N, tok, device=dev, dtype=torch.bfloat16)
        def old():
            ops.transpose(dy, out_t=dyT); ops.transpose(x, out_t=xT); ops.gemm_splitk(dyT, xT, out)
        us_old = timeit(old)
        print(f"wgrad {M}x{N} K={tok}: rel {e:.2e}  new {us:.1f} us  old(transposes+splitk) {us_old:.1f} us  "
              f"{2*M*N*tok/us/1e6:.0f} TFLOP/s", flush=True)
    except Exception as ex:
        print("wgrad FAIL", M, N, ex, flush=True)
for (N, K) in [(512, 512), (512, 1536), (512, 1024), (1024, 512)]:
    dy = torch.randn(tok, K, device=dev).bfloat16(); w = (torch.randn(K, N, device=dev) / K ** .5).bfloat16()
    o = torch.empty(tok, N, device=dev, dtype=torch.bfloat16); zb = torch.zeros(N, device=dev)
    try:
        ops.gemm_t(dy, w, trans_b=True, bias=zb, out_bf16=o)
        torch.cuda.synchronize()
        e = rel(o, dy.double() @ w.double())
        us = timeit(lambda: ops.gemm_t(dy, w, trans_b=True, bias=zb, out_bf16=o))
        wT = w.t().contiguous()
        us_old = timeit(lambda: ops.gemm(dy, wT, bias=zb, out_bf16=o))
        print(f"dgrad tok x{N} K={K}: rel {e:.2e}  new {us:.1f} us  old(W^T copy) {us_old:.1f} us  {2*tok*N*K/us/1e6:.0f} TFLOP/s", flush=True)
    except Exception as ex:
        print("dgrad FAIL", N, K, ex, flush=True)

# attention backward old vs new
S, T, H = 256, 91, 8
D = 512
qkv = torch.randn(S * T, 3 * D, device=dev).bfloat16(); dyy = torch.randn(S * T, D, device=dev).bfloat16()
lens = torch.randint(20, T + 1, (S,), device=dev).int()
d = torch.empty_like(qkv)
f = lambda: ops.eff_attn_bwd(ops.ATTN_SELF, S, T, H, q=qkv[:, :D], k=qkv[:, D:2*D], v=qkv[:, 2*D:], dy=dyy, dq=d[:, :D], dk=d[:, D:2*D], dv=d[:, 2*D:], length=lens)
print(f"eff_attn_bwd SELF S={S} T={T}: {timeit(f):.1f} us (HIG_ATTN_BWD_TC={os.environ.get('HIG_ATTN_BWD_TC','1')})", flush=True)
f2 = lambda: ops.eff_attn_bwd(ops.ATTN_INTER, S, T, H, q=qkv[:, :D], k=qkv[:, D:2*D], v=qkv[:, 2*D:], dy=dyy, dq=d[:, :D], dk=d[:, D:2*D], dv=d[:, 2*D:], length=lens, pair_shift=S//2)
print(f"eff_attn_bwd INTER: {timeit(f2):.1f} us", flush=True)
