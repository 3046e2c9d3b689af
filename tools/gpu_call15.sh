#!/bin/bash
mkdir -p gpurun_out
VX_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_decode_variants.py -m gpu -q -k "t5one or (tcgen05_attention and 4)" 2>&1 | tail -12 | tee gpurun_out/call15_tests.log
for v in "" "VX_ATTN_T5=3" "VX_ATTN_T5=4"; do for b in 1 8; do echo "-- ${v:-default} B=$b"; env $v timeout 200 python tools/bench_configs.py --config 4 --batch $b --reps 1 2>&1 | tail -1 | cut -c130-330; done; done | tee gpurun_out/call15_cfg4.log
