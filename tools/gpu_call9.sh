#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_decode_variants.py -m gpu -x -q -k "bf16 or fused or bench_shape or eos or smoke or longctx" 2>&1 | tail -8 | tee gpurun_out/call9_tests.log
VX_TIMELINE_ALL=0 timeout 120 python tools/decode_timeline.py 1 2>&1 | grep -E "us/step|us per phase|rror" | head -8 | tee gpurun_out/call9_timeline.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-configs --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-1500 | tee gpurun_out/call9_bench.json
