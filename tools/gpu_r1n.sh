#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== A/B"; timeout 300 python tools/step_ab.py 200 4 HIG_EMBED_STREAM=0,1 2>&1 | grep -v "Warn\|textTrans" | tee gpurun_out/step_ab_embed.txt
echo "== breakdown"; timeout 300 python tools/step_breakdown.py 200 2>&1 | grep -v "Warn\|textTrans" | grep "full\|embed\|heads\|ddpm"
