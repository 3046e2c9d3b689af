#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q -k "tcgen05_gemm or bf16_teacher or vocos or smoke" 2>&1 | tail -4 | tee gpurun_out/call14_tests.log
for b in 8 32; do timeout 300 python tools/bench_configs.py --config 4 --batch $b --reps 1 2>&1 | tail -1 | cut -c1-420; done | tee gpurun_out/call14_cfg4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ar_decode_fused -c 1 -f -o gpurun_out/prof_decode_fused \
  python bench.py --steps 1 --warmup 1 --min-warmup 0 --t-new 200 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/call14_ncu_fused.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/call14_bench.err | tail -1 > gpurun_out/call14_bench.json
cut -c1-600 gpurun_out/call14_bench.json
