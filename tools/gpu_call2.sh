#!/bin/bash
# batched decode kernel v2: parity tests, config-3 timing, in-kernel timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode_variants.py tests/test_gpu_kernels.py -m gpu -x -q -k "batched or rejects" 2>&1 | tail -25 > gpurun_out/call2_tests.log
cat gpurun_out/call2_tests.log
for b in 8 32; do
  timeout 120 python tools/bench_configs.py --config 3 --batch $b --steps 300 --reps 1 2>&1 | tail -1 | cut -c1-400
done | tee gpurun_out/call2_cfg3.log
for b in 8 32; do timeout 120 python tools/decode_timeline.py $b 2>&1 | grep -E "us/step|us per phase|rror" | head -9; done | tee gpurun_out/call2_timeline.log
