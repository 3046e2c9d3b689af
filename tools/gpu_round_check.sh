#!/bin/bash
# One GPU call that refreshes the round's evidence: full GPU test suite, the bench line (with configs + CPU baseline), ncu --set full
# of the batch-1 decode kernel, the ncu launch list of one bench step.  Usage on the GPU box: bash tools/gpu_round_check.sh <tag>
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${tag}_gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/${tag}_bench.err | tail -1 > gpurun_out/${tag}_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ar_decode_fused -c 1 -f -o gpurun_out/prof_decode_fused \
  python bench.py --steps 1 --warmup 1 --min-warmup 0 --t-new 200 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/${tag}_ncu_fused.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --min-warmup 0 --no-configs --no-cpu-baseline --no-e2e > gpurun_out/${tag}_launches.log 2>&1
cut -c1-400 gpurun_out/${tag}_bench.json
