"""Experiments on the projection kernels: operand residency (R buffer sets), K length, epilogue knobs (HIG_GS_DBG)."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hig_b200  # noqa
from hig_b200 import ops
dev = torch.device("cuda:0")
M = 128 * 196
for (N, K, R) in [(1536, 512, 6), (1536, 512, 1), (512, 512, 6), (512, 512, 1), (1536, 2048, 3), (1024, 512, 6)]:
    A = [torch.randn(M, K, device=dev).bfloat16() for _ in range(R)]
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
    b = torch.randn(N, device=dev)
    O = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(R)]
    fns = {"stream": lambda i: ops.gemm_stream(ops.GS_BF16, A[i], w, b, O[i]),
           "old": lambda i: ops.gemm(A[i], w, bias=b, out_bf16=O[i]),
           "cublas": lambda i: torch.matmul(A[i], w.t(), out=O[i])}
    line = f"N={N} K={K} sets={R}:"
    for nm, fn in fns.items():
        for i in range(max(R, 3)):
            fn(i % R)
        torch.cuda.synchronize()
        n = 60
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i % R)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        line += f"  {nm} {us:6.1f} us ({2.0 * M * N * K / us / 1e6:6.0f} TF)"
    print(line)
