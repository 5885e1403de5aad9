#!/bin/bash
# training path on the GPU box: DP-path test + bench_train.  usage: gpu_train.sh TAG [bench args]
TAG=${1:-train}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_dist_train.py tests/test_gpu_backward.py -x -q 2>&1 | tail -30 | tee $OUT/${TAG}_pytest.log
timeout 900 python bench_train.py "$@" > $OUT/${TAG}_bench_train.json 2> $OUT/${TAG}_bench_train.err
tail -c 2500 $OUT/${TAG}_bench_train.json; tail -5 $OUT/${TAG}_bench_train.err
