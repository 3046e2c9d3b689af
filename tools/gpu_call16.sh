#!/bin/bash
mkdir -p gpurun_out
VX_ATTN_T5=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_t5_onepass -s 20 -c 1 -f -o gpurun_out/prof_attn_onepass \
  python tools/bench_configs.py --config 4 --batch 1 --reps 0 > gpurun_out/call16_ncu.log 2>&1
tail -2 gpurun_out/call16_ncu.log | cut -c1-300
