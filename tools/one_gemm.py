import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hig_b200  # noqa
from hig_b200 import ops
dev = torch.device("cuda:0")
M, N, K = 25088, int(sys.argv[1]) if len(sys.argv) > 1 else 1536, int(sys.argv[2]) if len(sys.argv) > 2 else 512
kind = sys.argv[3] if len(sys.argv) > 3 else "bf16"
a = [torch.randn(M, K, device=dev).bfloat16() for _ in range(4)]
w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
b = torch.randn(N, device=dev)
o = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(4)]
x = [torch.randn(M, N, device=dev) for _ in range(4)] if kind == "res" else None
for i in range(8):
    if kind == "res":
        ops.gemm(a[i % 4], w, bias=b, residual=x[i % 4], out_f32=x[i % 4])
    else:
        ops.gemm(a[i % 4], w, bias=b, out_bf16=o[i % 4])
torch.cuda.synchronize()
print("done")
