#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_kernels.py > gpurun_out/bench_kernels.log 2>&1; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_kernels.json"))
for k,v in d.items(): print(k, {a: round(b,1) for a,b in v.items()})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 800 --csv --log-file gpurun_out/launches_warm.csv python tools/profile_step.py 3 > gpurun_out/ncu_warm.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_warm.csv 0.667 > gpurun_out/launch_summary_warm.txt 2>&1; head -16 gpurun_out/launch_summary_warm.txt
for i in 1 2; do timeout 300 python tools/step_time.py 300 2>&1 | grep -v Warn | tail -3; done
