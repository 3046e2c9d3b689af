#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/call12_tests.log
timeout 400 python tools/bench_configs.py --config 5 --batch 8 --reps 1 2>&1 | tail -1 | cut -c1-700 | tee gpurun_out/call12_cfg5.log
