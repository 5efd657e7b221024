#!/bin/bash
# checkpoint D evidence: full GPU suite, bench line (both arms), in-graph class costs, ncu launch list, ncu --set full of the
# dominant kernels.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 300 python tools/step_breakdown.py 200 2>&1 | grep -v "Warn\|textTrans" > gpurun_out/step_breakdown.txt; cat gpurun_out/step_breakdown.txt
P="python tools/profile_step.py 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $P > gpurun_out/ncu_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv 0.667 > gpurun_out/launch_summary.txt 2>&1; head -14 gpurun_out/launch_summary.txt
P="python tools/profile_step.py 2"
for spec in "gemm_wres:gemm_wres_kernel:10:8" "gemm_stream:gemm_stream_kernel:1:1" "apply:attn_apply_stylize_kernel:30:1" "kv:attn_kv_kernel:20:1"; do
  IFS=: read name pat skip cnt <<< "$spec"
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:${pat}" -s $skip -c $cnt -f -o gpurun_out/prof_${name} $P > gpurun_out/ncu_${name}.log 2>&1
  echo "$name rc=$?"
done
