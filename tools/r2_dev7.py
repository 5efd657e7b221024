"""Round-2 dev check: cost of MN-major operands in the tcgen05 GEMM (wgrad form, K = 23296 tokens)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hig_b200  # noqa
from hig_b200 import ops
dev = torch.device("cuda:0")
def hot(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
tok = 23296
for (M, N) in [(512, 512), (1536, 512), (1024, 512)]:
    dy = torch.randn(tok, M, device=dev).bfloat16(); x = torch.randn(tok, N, device=dev).bfloat16()
    dyT, xT = dy.t().contiguous(), x.t().contiguous()
    out = torch.zeros(M, N, device=dev)
    res = {}
    res["A mn, B mn"] = hot(lambda: ops.gemm_t(dy, x, trans_a=True, trans_b=True, out_f32=out, split_k=-1))
    res["A mn, B k "] = hot(lambda: ops.gemm_t(dy, xT, trans_a=True, trans_b=False, out_f32=out, split_k=-1))
    res["A k , B mn"] = hot(lambda: ops.gemm_t(dyT, x, trans_a=False, trans_b=True, out_f32=out, split_k=-1))
    res["A k , B k "] = hot(lambda: ops.gemm_splitk(dyT, xT, out))
    print(f"dW {M}x{N} over {tok} tokens: " + "  ".join(f"[{k}] {v:.1f} us" for k, v in res.items()) +
          f"   (floor at 1400 TFLOP/s: {2*M*N*tok/1.4e15*1e6:.1f} us)", flush=True)
