#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
python tools/step_breakdown.py 100 2>&1 | grep -v Warn | tail -25
timeout 300 python tools/step_time.py 300 2>&1 | grep -v Warn | tail -4
timeout 600 python tools/train_step.py --denoiser-only --iters 5 2>/dev/null | tail -1
timeout 600 python tools/train_step.py --iters 5 2>/dev/null | tail -1
