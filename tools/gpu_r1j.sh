#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -4
echo "== trace"; timeout 200 python tools/gemm_trace.py 2>&1 | grep -v Warn | tee gpurun_out/gemm_trace_kbw.txt
echo "== A/B"; timeout 300 python tools/step_ab.py 200 4 HIG_TIME_TABLE=0,1 2>&1 | grep -v "Warn\|textTrans" | tee gpurun_out/step_ab_timetable.txt
