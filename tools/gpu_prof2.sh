#!/bin/bash
mkdir -p gpurun_out
P="python tools/profile_step.py 2"
for spec in "apply:attn_apply_stylize_kernel:30:1" "kv:eff_attn_bf16_kernel:20:1"; do
  IFS=: read name pat skip cnt <<< "$spec"
  timeout 500 ncu --set full --clock-control none --import-source on -k "regex:${pat}" -s $skip -c $cnt -f -o gpurun_out/prof_${name} $P > gpurun_out/ncu_${name}.log 2>&1
  echo "$name rc=$?"
done
