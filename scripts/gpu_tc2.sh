#!/bin/bash
# bf16 sequence path: parity tests, phase profile, forward bench (+ the same bench on the v1 kernel with V1=1)
TAG=${1:-tc}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_umma.py -x -q -m gpu 2>&1 | tail -15 | tee $OUT/${TAG}.log
timeout 300 python scripts/seq_multi_profile.py 2>&1 | tee $OUT/${TAG}_phase.log | tail -45
timeout 600 python bench.py --precision bf16 --steps 30 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python -c "
import json; d=json.load(open('$OUT/${TAG}_bench.json')); r=d['roofline']; print('ms/step', d['ms_per_step'], 'value', d['value']); print(r['stage_share'], r['avg_launch_ms']); print(d['e2e'])"
tail -3 $OUT/${TAG}_bench.err
