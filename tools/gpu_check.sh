#!/bin/bash
# One GPU-box visit: parity tests, per-kernel micro-bench, graph-replay step time, ncu launch list, bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_kernels.py > gpurun_out/bench_kernels.log 2>&1
timeout 300 python tools/step_time.py 200 > gpurun_out/step_time.log 2>&1; cat gpurun_out/step_time.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 3 > gpurun_out/ncu_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv 0.667 > gpurun_out/launch_summary.txt 2>&1; head -30 gpurun_out/launch_summary.txt
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
