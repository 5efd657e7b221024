#!/bin/bash
# checkpoint D visit 2: interleaved A/B (resident-W x L2 persistence), ncu full of the resident-W kernels
mkdir -p gpurun_out
echo "== A/B"; timeout 300 python tools/step_ab.py 200 4 HIG_WRES=0,1 HIG_L2_PERSIST=1,0 2>&1 | grep -v "Warn\|textTrans" | tee gpurun_out/step_ab.txt
echo "== ncu wres qkv"; HIG_WRES=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_wres -s 5 -c 2 -o gpurun_out/prof_wres_qkv -f python tools/one_stream.py 1536 512 > gpurun_out/ncu_wres_qkv.log 2>&1; tail -2 gpurun_out/ncu_wres_qkv.log
echo "== ncu stream qkv"; HIG_WRES=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_stream -s 5 -c 2 -o gpurun_out/prof_stream_qkv -f python tools/one_stream.py 1536 512 > gpurun_out/ncu_stream_qkv.log 2>&1; tail -2 gpurun_out/ncu_stream_qkv.log
ls -la gpurun_out/*.ncu-rep
