#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1; do
HIG_GS_DBG=$dbg timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_ws -s 5 -c 2 -o gpurun_out/prof_ws_dbg$dbg -f python tools/one_stream.py > gpurun_out/ncu_ws_$dbg.log 2>&1
done
HIG_GS_WS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_stream -s 5 -c 2 -o gpurun_out/prof_stream -f python tools/one_stream.py > gpurun_out/ncu_stream.log 2>&1
tail -3 gpurun_out/ncu_ws_0.log
ls -la gpurun_out/*.ncu-rep
