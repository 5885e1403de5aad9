#!/bin/bash
TAG=${1:-tc}
OUT=gpurun_out; mkdir -p $OUT
for t in "test_mmoe_bf16" "test_seq_encode_bf16_matches_oracle" "test_seq_encode_bf16_edge" "test_inference_bf16" "test_bf16_rejects"; do
  echo "== $t" | tee -a $OUT/${TAG}.log
  timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$t" 2>&1 | tail -25 | tee -a $OUT/${TAG}.log
done
echo "== bench bf16" | tee -a $OUT/${TAG}.log
timeout 600 python bench.py --precision bf16 --steps 30 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 2500 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
