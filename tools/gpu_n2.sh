#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
echo "== 2-GPU tests"; timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -x -q -k "nccl" 2>&1 | tail -3
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 2>&1 | grep -v "Warn\|textTrans\|OMP_NUM" | tail -2 | tee gpurun_out/bench_n2.json
echo "== reference arm N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>&1 | grep -v "Warn\|textTrans\|OMP_NUM" | tail -1 | cut -c1-300
