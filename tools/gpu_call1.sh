#!/bin/bash
# Round-2 first GPU call: measure every unmeasured variant, then ncu the kernels the bench actually runs.
mkdir -p gpurun_out
timeout 1500 bash tools/run_experiments.sh full > gpurun_out/exp_stdout.log 2>&1
# ncu --set full of the batch-1 fused decode kernel (bench path, short forced length)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ar_decode_fused -c 1 -f -o gpurun_out/r02_prof_decode_fused \
  python bench.py --steps 1 --warmup 1 --min-warmup 0 --t-new 200 --no-cpu-baseline --no-e2e > gpurun_out/ncu_fused.log 2>&1
# ncu --set full of the batched mma decode kernel (config 3, B=32, 100 steps)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ar_decode_mma -c 1 -f -o gpurun_out/r02_prof_decode_mma32 \
  python tools/bench_configs.py --config 3 --batch 32 --steps 100 --reps 0 > gpurun_out/ncu_mma.log 2>&1
ls -la gpurun_out
