#!/bin/bash
mkdir -p gpurun_out
VX_TIMELINE_ALL=1 timeout 120 python tools/decode_timeline.py 1 2>&1 | grep -E "us/step|us per phase|rror|CTA slot" | head -24 | tee gpurun_out/call10_timeline.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bf16_teacher or fused or bench_shape or kv_cache or large_ragged" 2>&1 | tail -5 | tee gpurun_out/call10_tests.log
