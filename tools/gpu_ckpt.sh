#!/bin/bash
# One GPU-box visit = one evidence checkpoint: parity tests, per-GEMM timing, in-graph class costs, bench line,
# ncu launch list (cold), ncu --set full captures of the dominant kernels.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_stream.py > gpurun_out/bench_stream.txt 2>&1; cat gpurun_out/bench_stream.txt
timeout 300 python tools/step_breakdown.py 200 > gpurun_out/step_breakdown.txt 2>&1; cat gpurun_out/step_breakdown.txt
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
P="python tools/profile_step.py 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $P > gpurun_out/ncu_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv 0.667 > gpurun_out/launch_summary.txt 2>&1; head -16 gpurun_out/launch_summary.txt
P="python tools/profile_step.py 2"
for spec in "gemm_stream:gemm_stream_kernel:12:9" "apply:attn_apply_stylize_kernel:30:1" "kv:attn_kv_kernel:20:1"; do
  IFS=: read name pat skip cnt <<< "$spec"
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:${pat}" -s $skip -c $cnt -f -o gpurun_out/prof_${name} $P > gpurun_out/ncu_${name}.log 2>&1
  echo "$name rc=$?"
done
