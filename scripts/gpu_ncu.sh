#!/bin/bash
# ncu: launch list + full capture of one kernel.  usage: gpu_ncu.sh TAG KERNEL_REGEX SKIP [bench args...]
TAG=$1; KREGEX=$2; SKIP=$3; shift 3
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > $OUT/${TAG}_ncu_bench.log 2>&1
tail -2 $OUT/${TAG}_ncu_bench.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $SKIP -c 3 \
  -o $OUT/${TAG}_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log | cut -c1-300
ls -la $OUT | tail -5
