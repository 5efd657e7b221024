#!/bin/bash
mkdir -p gpurun_out
echo "== new tests"; timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "query_softmax or presoftmaxed or layernorm_folded" 2>&1 | tail -4
echo "== all"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== A/B"; timeout 400 python tools/step_ab.py 200 4 HIG_QSM=0,1 2>&1 | grep -v "Warn\|textTrans" | tee gpurun_out/step_ab_qsm.txt
