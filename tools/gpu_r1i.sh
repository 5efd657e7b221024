#!/bin/bash
mkdir -p gpurun_out
echo "== branch test"; timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -s -k "branched or loop" 2>&1 | grep -v "Warn\|textTrans" | tail -8
echo "== A/B"; timeout 400 python tools/step_ab.py 200 4 HIG_BRANCHES=1,2,4 2>&1 | grep -v "Warn\|textTrans" | tee gpurun_out/step_ab_branches.txt
