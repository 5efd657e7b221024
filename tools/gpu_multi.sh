#!/bin/bash
# 2-GPU visit: NCCL gradient all-reduce check, sharded sampling bench, DDP training step
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29611 tools/ddp_check.py > gpurun_out/ddp_check_n$N.log 2>&1; tail -3 gpurun_out/ddp_check_n$N.log
timeout 600 $TR --master-port 29612 tools/train_step.py --denoiser-only --iters 5 > gpurun_out/train_n$N.json 2> gpurun_out/train_n$N.err; cat gpurun_out/train_n$N.json; tail -2 gpurun_out/train_n$N.err
timeout 900 $TR --master-port 29613 bench.py --gpus $N --steps 1 --warmup 3 --diffusion-steps 200 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json; tail -2 gpurun_out/bench_n$N.err
