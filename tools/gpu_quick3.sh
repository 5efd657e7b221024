#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -5
python tools/step_breakdown.py 100 2>&1 | grep -v Warn | tail -12
for c in 1 3 4 7; do echo "HIG_APPLY_CHUNKS=$c"; HIG_APPLY_CHUNKS=$c python tools/step_breakdown.py 100 2>&1 | grep -A1 "without attn_apply"; done
