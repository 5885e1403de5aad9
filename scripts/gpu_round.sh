#!/bin/bash
# One gpurun call: parity tests, smoke, bench, ncu launch list + one full capture of the top kernel.
# Usage (from the repo root, on the GPU box): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu" | tee $OUT/${TAG}_pytest.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 | tee -a $OUT/${TAG}_pytest.log
echo "== smoke" | tee $OUT/${TAG}_smoke.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee -a $OUT/${TAG}_smoke.log
echo "== bench"
timeout 900 python bench.py --steps 30 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 3000 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
tail -c 1500 $OUT/${TAG}_bench_ref.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
tail -3 $OUT/${TAG}_ncu_bench.log
echo "== ncu full capture"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:seq_encode -s 9 -c 3 \
  -o $OUT/${TAG}_seq_encode -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
tail -3 $OUT/${TAG}_ncu_full.log
ls -la $OUT
