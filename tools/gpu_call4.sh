#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode_variants.py tests/test_gpu_parity.py -m gpu -x -q -k "batched or longctx or top_p or large_ragged or rejects or topk or sampling" 2>&1 | tail -15 | tee gpurun_out/call4_tests.log
for b in 32; do timeout 120 python tools/decode_timeline.py $b 2>&1 | grep -E "us/step|us per phase|rror" | head -9; done | tee gpurun_out/call4_timeline.log
timeout 900 python bench.py --steps 3 --warmup 3 2> gpurun_out/call4_bench.err | tail -1 > gpurun_out/call4_bench.json
tail -3 gpurun_out/call4_bench.err; cut -c1-3000 gpurun_out/call4_bench.json
