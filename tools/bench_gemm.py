"""GEMM micro-benchmark with rotating operand sets (no dirty-L2 flush artefacts): back-to-back launches cycling
over enough distinct A/out buffers to exceed the 126 MB L2, CUDA-event timed."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hig_b200  # noqa
from hig_b200 import ops

dev = torch.device("cuda:0")
S, T = 128, 196
tok = S * T
SHAPES = {"qkv": (tok, 1536, 512, "bf16"), "q": (tok, 512, 512, "bf16"), "ffn1": (tok, 1024, 512, "gelu"),
          "ffn2": (tok, 512, 1024, "bf16"), "outproj": (tok, 512, 512, "res"), "outproj_xb": (tok, 512, 512, "res2")}
only = sys.argv[1:] or list(SHAPES)
for name in only:
    M, N, K, kind = SHAPES[name]
    per_set = M * K * 2 + M * N * (2 if kind in ("bf16", "gelu") else 10)
    R = max(2, int(400e6 // per_set) + 1)
    A = [torch.randn(M, K, device=dev).bfloat16() for _ in range(R)]
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
    b = torch.randn(N, device=dev)
    if kind in ("bf16", "gelu"):
        O = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(R)]
        fn = lambda i: ops.gemm(A[i], w, bias=b, out_bf16=O[i], act=1 if kind == "gelu" else 0)
    else:
        X = [torch.randn(M, N, device=dev) for _ in range(R)]
        O2 = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(R)]
        fn = lambda i: ops.gemm(A[i], w, bias=b, residual=X[i], out_f32=X[i], out_bf16=O2[i] if kind == "res2" else None)
    for i in range(R):
        fn(i)
    torch.cuda.synchronize()
    n = 5 * R
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i % R)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / n * 1e-3
    # cuBLAS reference for the plain product
    e0.record()
    for i in range(n):
        torch.matmul(A[i % R], w.t())
    e1.record()
    torch.cuda.synchronize()
    tc = e0.elapsed_time(e1) / n * 1e-3
    print(f"{name:12s} M={M} N={N} K={K} {kind:5s}: {t * 1e6:7.1f} us  {2.0 * M * N * K / t / 1e12:7.1f} TF/s   "
          f"(cuBLAS plain {tc * 1e6:6.1f} us)  sets={R}")
    del A
