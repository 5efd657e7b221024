#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -4
for p in 1 0; do echo "HIG_PDL=$p"; HIG_PDL=$p python tools/step_breakdown.py 200 2>&1 | grep "full denoiser"; HIG_PDL=$p timeout 300 python tools/step_time.py 300 2>&1 | grep "run 2"; done
