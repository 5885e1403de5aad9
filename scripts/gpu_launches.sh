#!/bin/bash
# quick forward bench + ncu launch list.  usage: gpu_launches.sh TAG
TAG=${1:-ll}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python bench.py --precision bf16 --steps 30 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python -c "
import json; d=json.load(open('$OUT/${TAG}_bench.json')); r=d['roofline']; print('ms/step', d['ms_per_step'], 'value', d['value']); print(r['stage_share'], r['avg_launch_ms']); print(d['e2e']['value'], d['e2e']['ms_per_step'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
python scripts/launch_table.py $OUT/${TAG}_launches.csv | tee $OUT/${TAG}_launch_table.md
