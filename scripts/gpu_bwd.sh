#!/bin/bash
# backward parity on the GPU box.  usage: gpu_bwd.sh TAG [pytest -k expr]
TAG=${1:-bwd}; K=${2:-}
OUT=gpurun_out; mkdir -p $OUT
if [ -n "$K" ]; then
  timeout 900 python -m pytest tests/test_gpu_backward.py -x -q -k "$K" 2>&1 | tail -40 | tee $OUT/${TAG}.log
else
  timeout 900 python -m pytest tests/test_gpu_backward.py -x -q 2>&1 | tail -40 | tee $OUT/${TAG}.log
fi
