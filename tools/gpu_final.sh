#!/bin/bash
# final evidence of the round: full GPU suite, smoke, bench (both arms), ncu launch list, in-graph class costs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "Warn\|textTrans" | tail -5 | tee gpurun_out/smoke.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
timeout 300 python tools/step_breakdown.py 200 2>&1 | grep -v "Warn\|textTrans" > gpurun_out/step_breakdown.txt; cat gpurun_out/step_breakdown.txt
P="python tools/profile_step.py 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $P > gpurun_out/ncu_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv 0.667 > gpurun_out/launch_summary.txt 2>&1; head -14 gpurun_out/launch_summary.txt
