#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/call5_tests.log
for b in 8 32; do timeout 120 python tools/decode_timeline.py $b 2>&1 | grep -E "us/step|us per phase|rror" | head -9; done | tee gpurun_out/call5_timeline.log
for b in 8 32; do timeout 200 python tools/bench_configs.py --config 3 --batch $b --steps 1500 --reps 1 2>&1 | tail -1 | cut -c1-420; done | tee gpurun_out/call5_cfg3.log
