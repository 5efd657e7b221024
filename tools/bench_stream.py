"""Timing of the TMA-epilogue projections (gemm_stream.cu) against the register-transposing ones at the C2 shapes,
rotating operand sets larger than L2, CUDA events."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hig_b200  # noqa
from hig_b200 import ops

dev = torch.device("cuda:0")
tok = 128 * 196
CASES = [("qkv bf16", ops.GS_BF16, 1536, 512), ("qkv LN-folded", ops.GS_LN_BF16, 1536, 512), ("q LN-folded", ops.GS_LN_BF16, 512, 512),
         ("ffn1 gelu", ops.GS_BF16_GELU, 1024, 512), ("ffn2", ops.GS_BF16, 512, 1024), ("outproj res_h", ops.GS_RES_H, 512, 512)]
for name, kind, N, K in CASES:
    M = tok
    R = 6
    op_dt = torch.float16 if kind == ops.GS_LN_BF16 else torch.bfloat16
    A = [torch.randn(M, K, device=dev).to(op_dt) for _ in range(R)]
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(op_dt)
    b = torch.randn(N, device=dev)
    wsum = w.float().sum(1).contiguous()
    stats = torch.empty(M, 8, device=dev)
    if kind == ops.GS_RES_H:
        O = [torch.randn(M, N, device=dev).half() for _ in range(R)]
    else:
        O = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(R)]
    if kind == ops.GS_LN_BF16:
        ops.row_stats(A[0], stats)

    def new(i):
        if kind == ops.GS_LN_BF16:
            ops.gemm_stream(kind, A[i], w, b, O[i], wsum=wsum, stats_in=stats, ln_width=K)
        elif kind == ops.GS_RES_H:
            ops.gemm_stream(kind, A[i], w, b * 0, O[i], stats_out=stats)
        else:
            ops.gemm_stream(kind, A[i], w, b, O[i])

    def old(i):
        if kind == ops.GS_RES_H:
            ops.gemm(A[i], w, bias=b * 0, residual=O[i], out_f32=O[i])
        elif kind == ops.GS_LN_BF16:
            return
        else:
            ops.gemm(A[i], w, bias=b, out_bf16=O[i], act=1 if kind == ops.GS_BF16_GELU else 0)

    res = []
    for fn in (new, old):
        for i in range(R):
            fn(i)
        torch.cuda.synchronize()
        n = 10 * R
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i % R)
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / n * 1e3)
    fl = 2.0 * M * N * K
    print(f"{name:16s} N={N:5d} K={K:5d}: stream {res[0]:6.1f} us ({fl / res[0] / 1e6:7.1f} TF/s)   old {res[1]:6.1f} us")
