#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -4
echo "== trace PDL on"; timeout 200 python tools/gemm_trace.py 2>&1 | grep -v Warn | tee gpurun_out/gemm_trace_boundary.txt
echo "== trace PDL off"; HIG_PDL=0 timeout 200 python tools/gemm_trace.py 2>&1 | grep -v Warn | tee gpurun_out/gemm_trace_boundary_nopdl.txt
