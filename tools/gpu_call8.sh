#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -m gpu -x -q -k "vocos or smoke or rejects or tcgen05_gemm" 2>&1 | tail -15 | tee gpurun_out/call8_tests.log
for b in 8 32; do timeout 300 python tools/bench_configs.py --config 4 --batch $b --reps 1 2>&1 | tail -1 | cut -c1-600; done | tee gpurun_out/call8_cfg4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ar_decode_mma -c 1 -f -o gpurun_out/prof_decode_mma32 \
  python tools/bench_configs.py --config 3 --batch 32 --steps 60 --reps 0 > gpurun_out/call8_ncu_mma.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --min-warmup 0 --no-configs --no-cpu-baseline --no-e2e > gpurun_out/call8_launches.log 2>&1
ls -la gpurun_out | tail -8
