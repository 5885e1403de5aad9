#!/bin/bash
# ncu --set full capture of kernels matching a regex inside the forward bench.  usage: gpu_ncu_k.sh TAG REGEX SKIP COUNT
TAG=$1; KREGEX=$2; SKIP=$3; COUNT=${4:-1}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $SKIP -c $COUNT \
  -o $OUT/${TAG}_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log | cut -c1-300
