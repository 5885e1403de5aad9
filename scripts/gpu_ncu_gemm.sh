#!/bin/bash
# ncu --set full of a window of training-GEMM launches.  usage: gpu_ncu_gemm.sh TAG SKIP COUNT [bench_train args]
TAG=$1; SKIP=$2; COUNT=$3; shift 3
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_group -s $SKIP -c $COUNT \
  -o $OUT/${TAG}_gemm -f python bench_train.py --steps 2 --warmup 3 "$@" > $OUT/${TAG}_ncu_gemm.log 2>&1
tail -2 $OUT/${TAG}_ncu_gemm.log | cut -c1-200
ls -la $OUT/${TAG}_gemm.ncu-rep
