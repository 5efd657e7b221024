"""Tile-by-tile timeline of the resident-W projection kernel (hig_debug_trace): for each C2 projection shape, the mean
over CTA pairs of the clock64 stamps relative to kernel entry, plus the globaltimer span of the launch.
Back-to-back launches (PDL) with L2-warm operands, the last one traced."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hig_b200  # noqa
from hig_b200 import _lib, ops

dev = torch.device("cuda:0")
tok = 128 * 196
lib = _lib.load()
CASES = [("qkv LN-folded", ops.GS_LN_BF16, 1536, 512), ("q LN-folded", ops.GS_LN_BF16, 512, 512),
         ("ffn1 gelu", ops.GS_BF16_GELU, 1024, 512), ("outproj res_h", ops.GS_RES_H, 512, 512)]
for name, kind, N, K in CASES:
    M = tok
    op_dt = torch.float16 if kind == ops.GS_LN_BF16 else torch.bfloat16
    A = torch.randn(M, K, device=dev).to(op_dt)
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(op_dt)
    b = torch.randn(N, device=dev)
    wsum = w.float().sum(1).contiguous()
    stats = torch.empty(M, 8, device=dev)
    O = torch.randn(M, N, device=dev).half() if kind == ops.GS_RES_H else torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    if kind == ops.GS_LN_BF16:
        ops.row_stats(A, stats)

    def run():
        if kind == ops.GS_LN_BF16:
            ops.gemm_stream(kind, A, w, b, O, wsum=wsum, stats_in=stats, ln_width=K)
        elif kind == ops.GS_RES_H:
            ops.gemm_stream(kind, A, w, b * 0, O, stats_out=stats)
        else:
            ops.gemm_stream(kind, A, w, b, O)

    NL = 6
    buf = torch.zeros(NL * 74 * 32, device=dev, dtype=torch.long)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    lib.hig_debug_trace(buf.data_ptr(), NL)
    for _ in range(NL):
        run()
    torch.cuda.synchronize()
    lib.hig_debug_trace(None, 0)
    allt = buf.view(NL, 74, 32).cpu().double()
    # kernel boundary on the global timer: last CTA exit of launch n -> dependents of launch n+1 released (griddepcontrol.wait
    # returns), and -> first / mean CTA entry of launch n+1
    for n in range(2, NL - 1):
        end_n = allt[n, :, 30].max().item()
        nxt = allt[n + 1]
        print(f"   boundary {n}->{n + 1}: last exit -> gate open {(nxt[:, 31].min().item() - end_n) / 1e3:6.2f} us | "
              f"mean exit -> mean entry {(nxt[:, 29].mean().item() - allt[n, :, 30].mean().item()) / 1e3:6.2f} us | "
              f"entry spread {(nxt[:, 29].max().item() - nxt[:, 29].min().item()) / 1e3:5.2f} us | "
              f"exit spread {(allt[n, :, 30].max().item() - allt[n, :, 30].min().item()) / 1e3:5.2f} us")
    t = allt[NL - 1]
    if os.environ.get("TRACE_PER_PAIR"):
        # gate -> exit per pair (us) for two consecutive launches: the spread is the tile quantisation (2 vs 3 tiles at N = 512)
        for n in (NL - 2, NL - 1):
            a = allt[n]
            d = (a[:, 30] - a[:, 31]) / 1e3
            print(f"   launch {n}: gate->exit per pair: " + " ".join(f"{d[p]:.1f}" for p in range(74)))
    rel = t[:, :29] - t[:, :1]
    mean = rel.mean(0)
    ntile = [(t[:, 4 + i] > 0).sum().item() for i in range(8)]
    span = (t[:, 30].max() - t[:, 29].min()).item()
    own = (t[:, 30] - t[:, 29])
    print(f"== {name} N={N} K={K}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us/launch back-to-back; traced launch: globaltimer span "
          f"{span / 1e3:.1f} us, per-pair lifetime mean {own.mean().item() / 1e3:.1f} us (min {own.min().item() / 1e3:.1f}, max {own.max().item() / 1e3:.1f})")
    print(f"   clk from entry (mean over pairs): setup done {mean[1]:.0f} | producer past pdl_wait {mean[2]:.0f} | W resident {mean[3]:.0f} | end {mean[28]:.0f}")
    for i in range(8):
        if ntile[i] == 0:
            continue
        m = t[:, 4 + i] > 0
        f = lambda k: (rel[:, k][m]).mean().item()
        print(f"   tile {i} ({ntile[i]:2d} pairs): MMA issued {f(4 + i):7.0f} | accumulator ready {f(12 + i):7.0f} | epilogue done {f(20 + i):7.0f}")
