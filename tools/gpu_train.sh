#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/train_step.py --denoiser-only --iters 5 > gpurun_out/train_denoiser_only.json 2> gpurun_out/train1.err; cat gpurun_out/train_denoiser_only.json; tail -3 gpurun_out/train1.err
timeout 600 python tools/train_step.py --iters 5 > gpurun_out/train_text.json 2> gpurun_out/train2.err; cat gpurun_out/train_text.json; tail -3 gpurun_out/train2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_train.csv python tools/train_step.py --denoiser-only --iters 1 --warmup 1 > gpurun_out/ncu_train.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_train.csv 0.5 > gpurun_out/launch_summary_train.txt 2>&1; head -40 gpurun_out/launch_summary_train.txt
P="python tools/profile_step.py 2"
for spec in "gemm2cta:gemm_bf16_2cta_kernel:40:6" "ln:ln_film_silu_kernel:20:2"; do
  IFS=: read name pat skip cnt <<< "$spec"
  timeout 500 ncu --set full --clock-control none --import-source on -k "regex:${pat}" -s $skip -c $cnt -f -o gpurun_out/prof_${name} $P > gpurun_out/ncu_${name}.log 2>&1
  echo "$name rc=$?"
done
