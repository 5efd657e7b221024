"""A few launches of the QKV-shaped projection through hig_gemm_stream (for ncu)."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hig_b200  # noqa
from hig_b200 import ops
dev = torch.device("cuda:0")
M = 25088
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1536
K = int(sys.argv[2]) if len(sys.argv) > 2 else 512
R = 4
a = [torch.randn(M, K, device=dev).bfloat16() for _ in range(R)]
w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
b = torch.randn(N, device=dev)
o = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(R)]
for i in range(8):
    ops.gemm_stream(ops.GS_BF16, a[i % R], w, b, o[i % R])
torch.cuda.synchronize()
print("done")
