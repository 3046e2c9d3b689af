#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/call19_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/call19_bench.err | tail -1 > gpurun_out/call19_bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/call19_bench.json'))
print(d['value'], d['phase_ms'], d['roofline']['frac'], d['parity'] and d['parity']['ok'])
for k,v in d['configs'].items(): print(k, v.get('us_per_step'), v.get('ms_nar'), v.get('ms_vocos'), v['roofline']['frac'])
PY
