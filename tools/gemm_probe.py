import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hig_b200  # noqa
from hig_b200 import ops
dev = torch.device("cuda:0")
M = 25088
for (N, K) in [(1536, 512), (512, 512), (512, 1024)]:
    R = 4
    A = [torch.randn(M, K, device=dev).bfloat16() for _ in range(R)]
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
    b = torch.randn(N, device=dev)
    O = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(R)]
    for act in (0, 100, 101):
        for i in range(R):
            ops.gemm(A[i], w, bias=b, out_bf16=O[i], act=act)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for i in range(n):
            ops.gemm(A[i % R], w, bias=b, out_bf16=O[i % R], act=act)
        e1.record()
        torch.cuda.synchronize()
        print(f"N={N} K={K} act={act}: {e0.elapsed_time(e1) / n * 1e3:.1f} us")
