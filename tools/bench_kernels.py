"""Micro-benchmarks of the individual kernels at the C2 shapes (S=128, T=196): CUDA-event timing, L2 flushed
between iterations.  Development aid; the contract benchmark is /bench.py."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hig_b200  # noqa
from hig_b200 import ops

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def main():
    S, T, D = 128, 196, 512
    tok = S * T
    res = {}
    for name, (M, N, K, kind) in {
        "qkv": (tok, 1536, 512, "bf16out"), "q": (tok, 512, 512, "bf16out"), "ffn1": (tok, 1024, 512, "gelu"),
        "ffn2": (tok, 512, 1024, "bf16out"), "outproj": (tok, 512, 512, "res"), "emb": (S, 32768, 2048, "f32out"),
        "embed_in": (tok, 512, 272, "f32out"), "out": (tok, 263, 512, "f32out263"),
    }.items():
        a = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
        bias = torch.randn(N, device=dev)
        if kind == "bf16out":
            o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            fn = lambda: ops.gemm(a, w, bias=bias, out_bf16=o)
            byt = (M * K + N * K + M * N) * 2
        elif kind == "gelu":
            o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            fn = lambda: ops.gemm(a, w, bias=bias, out_bf16=o, act=1)
            byt = (M * K + N * K + M * N) * 2
        elif kind == "res":
            r = torch.randn(M, N, device=dev)
            o = torch.empty(M, N, device=dev)
            o2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            fn = lambda: ops.gemm(a, w, bias=bias, residual=r, out_f32=o, out_bf16=o2)
            byt = (M * K + N * K) * 2 + M * N * 10
        else:
            o = torch.empty(M, N, device=dev)
            fn = lambda: ops.gemm(a, w, bias=bias, out_f32=o)
            byt = (M * K + N * K) * 2 + M * N * 4
        t = timeit(fn)
        fl = 2.0 * M * N * K
        res["gemm_" + name] = {"us": t * 1e6, "tflops": fl / t / 1e12, "gbs": byt / t / 1e9}
        tt = timeit(lambda: torch.matmul(a, w.t()))
        res["gemm_" + name]["cublas_us"] = tt * 1e6
    # LN
    x = torch.randn(tok, D, device=dev)
    g, b = torch.randn(D, device=dev), torch.randn(D, device=dev)
    ob = torch.empty(tok, D, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: ops.ln_film_silu(x, g, b, ob, rows_per_seq=T))
    res["ln_f32_in"] = {"us": t * 1e6, "gbs": tok * D * 6 / t / 1e9}
    xb = x.bfloat16()
    ss = torch.randn(S, 1024, device=dev)
    t = timeit(lambda: ops.ln_film_silu(xb, g, b, ob, rows_per_seq=T, scale_shift=ss, silu=True))
    res["ln_film_silu_bf16"] = {"us": t * 1e6, "gbs": tok * D * 4 / t / 1e9}
    # attention
    qkv = torch.randn(tok, 3 * D, device=dev).bfloat16()
    lens = torch.full((S,), T, device=dev, dtype=torch.int32)
    y = torch.empty(tok, D, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: ops.eff_attn(ops.ATTN_SELF, S, T, 8, q=qkv[:, :D], k=qkv[:, D:2 * D], v=qkv[:, 2 * D:], y=y,
                                    length=lens))
    res["attn_self"] = {"us": t * 1e6, "gbs": tok * D * 8 / t / 1e9}
    t = timeit(lambda: ops.eff_attn(ops.ATTN_INTER, S, T, 8, q=qkv[:, :D], k=qkv[:, D:2 * D], v=qkv[:, 2 * D:], y=y,
                                    length=lens, pair_shift=S // 2, mask_v=False))
    res["attn_inter"] = {"us": t * 1e6, "gbs": tok * D * 8 / t / 1e9}
    a_t = torch.randn(S, 8, 64, 64, device=dev).bfloat16()
    qb = qkv[:, :D].contiguous()
    t = timeit(lambda: ops.eff_attn(ops.ATTN_Q_ONLY, S, T, 8, q=qb, a_in=a_t, y=y))
    res["attn_text_apply"] = {"us": t * 1e6, "gbs": (tok * D * 4 + a_t.numel() * 2) / t / 1e9}
    a_blk = torch.empty(S, 8, 64, 64, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: ops.eff_attn(ops.ATTN_KV_ONLY, S, T, 8, k=qkv[:, D:2 * D], v=qkv[:, 2 * D:], a_out=a_blk,
                                    length=lens))
    res["attn_kv_only"] = {"us": t * 1e6, "gbs": (tok * D * 4 + a_blk.numel() * 2) / t / 1e9}
    t = timeit(lambda: ops.attn_apply_stylize(qkv[:, :D], a_blk, g, b, ob, S, T, 8, scale_shift=ss, silu=True))
    res["attn_apply_stylize"] = {"us": t * 1e6, "gbs": (tok * D * 4 + a_blk.numel() * 2) / t / 1e9}
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/bench_kernels.json", "w"), indent=1)


if __name__ == "__main__":
    main()
