#!/bin/bash
# usage: gpu_train_ncu.sh TAG [bench_train args]  -- tests, bench, ncu launch list of ~2 training steps
TAG=${1:-train}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_dist_train.py tests/test_gpu_optim.py -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.log
timeout 900 python bench_train.py "$@" > $OUT/${TAG}_bench_train.json 2> $OUT/${TAG}_bench_train.err
tail -c 1800 $OUT/${TAG}_bench_train.json; tail -5 $OUT/${TAG}_bench_train.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv \
  --log-file $OUT/${TAG}_train_launches.csv python bench_train.py --steps 2 --warmup 3 "$@" > $OUT/${TAG}_ncu_train.log 2>&1
tail -2 $OUT/${TAG}_ncu_train.log | cut -c1-300
