#!/bin/bash
# direct-store epilogue of the resident-W GEMM: correctness, then A/B
mkdir -p gpurun_out
echo "== stream GEMM tests"; timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm_stream" 2>&1 | tail -4
echo "== parity"; timeout 400 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
for w in 0 1; do echo "== bench_stream HIG_WR_DIRECT=$w"; HIG_WR_DIRECT=$w timeout 200 python tools/bench_stream.py 2>&1 | grep -v Warning | tee gpurun_out/bench_stream_direct$w.txt; done
echo "== A/B"; timeout 300 python tools/step_ab.py 200 4 HIG_WR_DIRECT=0,1 2>&1 | grep -v "Warn\|textTrans" | tee gpurun_out/step_ab_direct.txt
echo "== step_breakdown"; timeout 400 python tools/step_breakdown.py 200 2>&1 | grep -v "Warn\|textTrans" | tee gpurun_out/step_breakdown_direct.txt
