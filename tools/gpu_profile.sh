#!/bin/bash
# Evidence pass at the C2 shape (S=128, T=196): bench line, ncu launch lists (cold + warm L2), --set full captures.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err
P="python tools/profile_step.py 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $P > gpurun_out/ncu_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv 0.667 > gpurun_out/launch_summary.txt 2>&1; head -14 gpurun_out/launch_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1200 --csv --log-file gpurun_out/launches_warm.csv $P > gpurun_out/ncu_warm.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_warm.csv 0.667 > gpurun_out/launch_summary_warm.txt 2>&1
P="python tools/profile_step.py 2"
for spec in "gemm2cta:gemm_bf16_2cta_kernel:38:7" "apply:attn_apply_stylize_kernel:30:1" "kv:attn_kv_kernel:20:1" "ln:ln_film_silu_kernel:20:2" "ddpm:ddpm_step_kernel:0:1"; do
  IFS=: read name pat skip cnt <<< "$spec"
  timeout 500 ncu --set full --clock-control none --import-source on -k "regex:${pat}" -s $skip -c $cnt -f -o gpurun_out/prof_${name} $P > gpurun_out/ncu_${name}.log 2>&1
  echo "$name rc=$?"
done
