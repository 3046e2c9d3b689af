#!/bin/bash
mkdir -p gpurun_out
VX_ATTN_T5=4 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_nar8.csv python tools/bench_configs.py --config 4 --batch 8 --reps 0 > gpurun_out/call18.log 2>&1
tail -1 gpurun_out/call18.log | cut -c1-200
