"""Clocks / power while one projection shape runs back to back for a few seconds (is the GEMM power-capped?)."""
import math, os, subprocess, sys, threading, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hig_b200  # noqa
from hig_b200 import ops
dev = torch.device("cuda:0")
M, N, K, R = 25088, 1536, 512, 4
A = [torch.randn(M, K, device=dev).bfloat16() for _ in range(R)]
w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
b = torch.randn(N, device=dev)
O = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(R)]
samples = []
stop = False
def sampler():
    while not stop:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,temperature.gpu",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip()
        samples.append(out)
        time.sleep(0.05)
for name, fn in (("stream", lambda i: ops.gemm_stream(ops.GS_BF16, A[i], w, b, O[i])),
                 ("cublas", lambda i: torch.matmul(A[i], w.t(), out=O[i])),
                 ("big cublas 8192^3", None)):
    if fn is None:
        X = torch.randn(8192, 8192, device=dev).bfloat16(); Y = torch.randn(8192, 8192, device=dev).bfloat16()
        fn = lambda i: torch.matmul(X, Y)
        flops = 2.0 * 8192 ** 3
    else:
        flops = 2.0 * M * N * K
    samples.clear(); stop = False
    th = threading.Thread(target=sampler); th.start()
    torch.cuda.synchronize()
    for seg in range(3):
        n = 20000 if flops < 1e11 else 1000
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i % R)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        print(f"{name} seg{seg}: {us:.1f} us  {flops / us / 1e6:.0f} TF")
    stop = True; th.join()
    print("   samples (MHz, W, power_cap, C):", " | ".join(samples[::max(1, len(samples) // 12)]))
