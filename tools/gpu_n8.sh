#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep -v "Warn\|textTrans\|OMP_NUM\|^\*\*\*" | tail -1 | tee gpurun_out/bench_n$N.json | cut -c1-700
