#!/bin/bash
# One gpurun call: full GPU parity suite, phase profile of the fused sequence kernel, forward + training bench.
TAG=${1:-chk}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.log
timeout 300 python scripts/seq_multi_profile.py 2>&1 | tee $OUT/${TAG}_phase.log | tail -48
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 2500 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
timeout 600 python bench_train.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_train.json 2> $OUT/${TAG}_bench_train.err
tail -c 1500 $OUT/${TAG}_bench_train.json; tail -3 $OUT/${TAG}_bench_train.err
