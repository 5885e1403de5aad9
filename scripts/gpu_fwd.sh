#!/bin/bash
# forward path: all forward parity tests + bench (stage shares) ; usage: gpu_fwd.sh TAG
TAG=${1:-fwd}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_demo_e2e.py -x -q -m gpu 2>&1 | tail -8 | tee $OUT/${TAG}.log
timeout 600 python bench.py --precision bf16 --steps 30 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python -c "
import json; d=json.load(open('$OUT/${TAG}_bench.json')); r=d['roofline']; print('ms/step', d['ms_per_step'], 'value', d['value']); print(r['stage_share'], r['avg_launch_ms']); print(d['e2e']['value'], d['e2e']['ms_per_step'])"
tail -3 $OUT/${TAG}_bench.err
