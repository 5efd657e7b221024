#!/bin/bash
mkdir -p gpurun_out
for p in 74 64 56 37 18; do echo "pairs=$p"; HIG_GS_PAIRS=$p timeout 200 python tools/bench_stream.py 2>&1 | grep -E "qkv bf16|ffn2|outproj"; done | tee gpurun_out/pairs.txt
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv
