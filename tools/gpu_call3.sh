#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode_variants.py -m gpu -x -q -k "batched" 2>&1 | tail -5 | tee gpurun_out/call3_tests.log
for mb in 0 40 80 110; do
  for b in 8 32; do
    echo "-- VX_KV_L2_MB=$mb B=$b"
    VX_KV_L2_MB=$mb timeout 120 python tools/bench_configs.py --config 3 --batch $b --steps 300 --reps 1 2>&1 | tail -1 | cut -c1-330
  done
done | tee gpurun_out/call3_cfg3.log
for b in 8 32; do timeout 120 python tools/decode_timeline.py $b 2>&1 | grep -E "us/step|us per phase|rror" | head -9; done | tee gpurun_out/call3_timeline.log
echo "-- full config 3 (1500 steps) B=32"; timeout 200 python tools/bench_configs.py --config 3 --batch 32 --steps 1500 --reps 1 2>&1 | tail -1 | tee -a gpurun_out/call3_cfg3.log
