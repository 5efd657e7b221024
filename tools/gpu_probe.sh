#!/bin/bash
mkdir -p gpurun_out
python tools/gemm_probe.py 2>&1 | tee gpurun_out/gemm_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
timeout 300 python tools/step_time.py 300 2>&1 | tail -3 | tee gpurun_out/step_time.txt
