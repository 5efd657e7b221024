"""Round-2 dev check #2 (GPU): attn_apply mma.sync vs tcgen05 at the C2 shape, L2 flushed between launches."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hig_b200  # noqa
from hig_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e3
def hot(fn, n=50):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
S, T, H, D = 128, 196, 8, 512
qkv = torch.softmax(torch.randn(S * T, 3, H, 64, device=dev), -1).view(S * T, 3 * D).bfloat16()
a = (torch.randn(S, H, 64, 64, device=dev) * .3).bfloat16(); a_t = a.transpose(-1, -2).contiguous()
gamma = torch.ones(D, device=dev); beta = torch.zeros(D, device=dev); ss = torch.randn(S, 2 * D, device=dev) * .3
o1 = torch.empty(S * T, D, device=dev, dtype=torch.bfloat16); o2 = torch.empty_like(o1)
f_old = lambda: ops.attn_apply_stylize(qkv[:, :D], a, gamma, beta, o1, S, T, H, scale_shift=ss, silu=True, q_softmaxed=True)
f_new = lambda: ops.attn_apply_stylize_tc(qkv[:, :D], a_t, gamma, beta, o2, S, T, H, scale_shift=ss, silu=True)
f_old(); f_new(); torch.cuda.synchronize()
print("rel new vs old", ((o1.float() - o2.float()).norm() / o1.float().norm()).item())
print(f"apply mma.sync : cold {timeit(f_old):.1f} us  hot {hot(f_old):.1f} us")
print(f"apply tcgen05  : cold {timeit(f_new):.1f} us  hot {hot(f_new):.1f} us")
lens = torch.full((S,), T, device=dev, dtype=torch.int32)
ak = torch.empty_like(a)
f_kv = lambda: ops.attn_kv(qkv[:, D:2 * D], qkv[:, 2 * D:], ak, S, T, H, length=lens)
f_kvt = lambda: ops.attn_kv(qkv[:, D:2 * D], qkv[:, 2 * D:], ak, S, T, H, length=lens, transposed=True)
print(f"attn_kv: cold {timeit(f_kv):.1f} us hot {hot(f_kv):.1f};  transposed: cold {timeit(f_kvt):.1f} hot {hot(f_kvt):.1f}")
