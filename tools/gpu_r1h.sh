#!/bin/bash
mkdir -p gpurun_out
for d in 1 0; do echo "#### HIG_WR_DIRECT=$d"; HIG_WR_DIRECT=$d timeout 200 python tools/gemm_trace.py 2>&1 | grep -v Warn | tee gpurun_out/gemm_trace_direct$d.txt; done
