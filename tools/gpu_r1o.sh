#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== breakdown"; timeout 300 python tools/step_breakdown.py 200 2>&1 | grep -v "Warn\|textTrans" | grep "full\|embed\|heads\|ddpm\|out-proj"
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep -v "Warn\|textTrans" | tail -1 | cut -c1-900 | tee gpurun_out/bench_r1o.json
