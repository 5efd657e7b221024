#!/bin/bash
# checkpoint D visit 3: joints kernel + device seed + cached sampling graph; full suite; bench line
mkdir -p gpurun_out
echo "== new tests"; timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -x -q -s -k "joints or seed or loop or teacher" 2>&1 | grep -v "Warn\|textTrans" | tail -25
echo "== all gpu tests"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r1f.txt
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 3 2>&1 | grep -v "Warn\|textTrans" | tail -2 | tee gpurun_out/bench_r1f.json
