#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "stream" 2>&1 | tail -5 | tee gpurun_out/pytest_stream.txt
timeout 300 python tools/bench_stream.py 2>&1 | tee gpurun_out/bench_stream.txt
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_parity.txt
timeout 300 python tools/step_time.py 300 2>&1 | tail -3 | tee gpurun_out/step_time.txt
timeout 300 python tools/step_breakdown.py 200 2>&1 | tail -25 | tee gpurun_out/step_breakdown.txt
