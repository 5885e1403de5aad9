#!/bin/bash
TAG=${1:-umma}
OUT=gpurun_out; mkdir -p $OUT
for m in 0 1 2; do
  echo "== mode $m" | tee -a $OUT/${TAG}_umma.log
  timeout 300 python -m pytest tests/test_gpu_umma.py -q -m gpu -k "${m}]" 2>&1 | tail -15 | tee -a $OUT/${TAG}_umma.log
done
echo "== parity" | tee -a $OUT/${TAG}_umma.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8 | tee -a $OUT/${TAG}_umma.log
