#!/bin/bash
# ncu captures at the C2 shape (S=128, T=196): --set full of the dominant kernels + a warm-L2 launch list.
mkdir -p gpurun_out
P="python tools/profile_step.py 2"
for spec in "qkv:gemm_bf16_2cta_kernel<1>:24" "res:gemm_bf16_2cta_kernel<4>:10" "attn:eff_attn_bf16_kernel:18" "ln:ln_film_silu_kernel<512, __nv_bfloat16:20"; do
  IFS=: read name pat skip <<< "$spec"
  timeout 500 ncu --set full --clock-control none --import-source on -k "regex:${pat}" -s $skip -c 2 -f -o gpurun_out/prof_${name} $P > gpurun_out/ncu_${name}.log 2>&1
  echo "$name rc=$?"
done
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 800 --csv --log-file gpurun_out/launches_warm.csv $P > gpurun_out/ncu_warm.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_warm.csv 0.5 > gpurun_out/launch_summary_warm.txt 2>&1; head -20 gpurun_out/launch_summary_warm.txt
