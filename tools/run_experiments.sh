#!/bin/bash
# One GPU call that measures every decode / attention variant built behind an environment switch (DESIGN.md sections 5, 9).
# Usage on the GPU box (from the repo root):   bash tools/run_experiments.sh [quick|full]   -> gpurun_out/experiments.log
# Every step has its own timeout; a variant that traps or times out only loses its own line.
mode=${1:-quick}
mkdir -p gpurun_out
out=gpurun_out/experiments.log
: > "$out"
say() { echo "== $*" | tee -a "$out"; }

say "parity of the variants that have not run on a GPU yet (VX_TEST_EXPERIMENTAL=1)"
VX_TEST_EXPERIMENTAL=1 timeout 240 python -m pytest tests/test_gpu_decode_variants.py tests/test_gpu_kernels.py -m gpu -q \
  -k "t5max or tcgen05_attention" 2>&1 | tail -15 >> "$out"

say "batch-1 decode, us/step (default = 3-barrier fused kernel)"
for v in "" "VX_DECODE_FUSED=0"; do
  echo "-- ${v:-default}" >> "$out"
  env $v VX_TIMELINE_ALL=0 timeout 60 python tools/decode_timeline.py 1 2>&1 | grep -E "us/step|us per phase|Error|error" | head -8 >> "$out"
done

say "batched decode (config 3): default (balanced warp attention) / CTA-per-head attention units / CUDA-core GEMV kernel"
for v in "" "VX_DECODE_ATTN=0" "VX_DECODE_MMA=0"; do
  for b in 8 32; do
    echo "-- ${v:-default} B=$b" >> "$out"
    env $v timeout 90 python tools/bench_configs.py --config 3 --batch $b --steps 300 --reps 1 2>&1 | tail -1 | cut -c1-260 >> "$out"
  done
done
if [ "$mode" = "full" ]; then
  say "where the batched step goes (in-kernel timeline)"
  for b in 8 32; do timeout 90 python tools/decode_timeline.py $b 2>&1 | grep -E "us/step|us per phase" | head -8 >> "$out"; done
  say "NAR + Vocos (config 4): mma.sync attention vs tcgen05 attention (1: statistics pass with exponentials, 2: max-only first pass, 3: + head images from the QKV GEMM epilogue)"
  for v in "" "VX_ATTN_T5=1" "VX_ATTN_T5=2" "VX_ATTN_T5=3"; do
    echo "-- ${v:-default}" >> "$out"
    env $v timeout 120 python tools/bench_configs.py --config 4 --batch 8 --reps 1 2>&1 | tail -1 | cut -c1-260 >> "$out"
  done
  say "long text through the public API (config 5): batched vs sentence-by-sentence"
  timeout 240 python tools/bench_configs.py --config 5 --batch 8 --reps 1 2>&1 | tail -1 | cut -c1-600 >> "$out"
fi
cat "$out"
