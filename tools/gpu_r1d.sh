#!/bin/bash
# checkpoint D visit 1: resident-W GEMM + packed-math attention kernels — correctness first, then A/B timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
echo "== stream GEMM tests (resident-W)"; timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm_stream" 2>&1 | tail -5
echo "== attention tests"; timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attn" 2>&1 | tail -5
echo "== all gpu tests"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r1d.txt
for w in 0 1; do echo "== bench_stream HIG_WRES=$w"; HIG_WRES=$w timeout 200 python tools/bench_stream.py 2>&1 | grep -v Warning | tee gpurun_out/bench_stream_wres$w.txt; done
for w in 0 1; do echo "== full step HIG_WRES=$w"; HIG_WRES=$w timeout 200 python tools/step_time.py 200 2>&1 | tail -3 | tee gpurun_out/step_time_wres$w.txt; done
echo "== step_breakdown HIG_WRES=1"; HIG_WRES=1 timeout 400 python tools/step_breakdown.py 200 2>&1 | grep -v "Warn\|textTrans" | tee gpurun_out/step_breakdown_wres1.txt
