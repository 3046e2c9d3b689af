#!/usr/bin/env python
"""Per-phase timeline of the persistent decode kernel (VX_DECODE_PROF=<first step>): CTAs 0 / grid/2 / grid-1 stamp
(tag, clock64) at every grid barrier arrival (1) / release (2) and inside the GEMV phases: after staging (3), after the
dot products + emit (4), after issuing the next phase's loads (5).  Prints mean segment durations in microseconds."""
import collections
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "vall-e-x_b200")):
    sys.path.insert(0, p)
os.environ.setdefault("VX_DECODE_PROF", "600")
import bench  # noqa: E402
from vallex_b200.engine import Engine  # noqa: E402
from vallex_b200.weights import synthetic_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
x, Lp, prompt = bench.workload_inputs()
eng = Engine(0, B, 2048, "bf16")
eng.load_valle_state_dict(synthetic_state_dict(0))
eng.finalize()
lang = np.zeros((len(x),), np.int32)
eng.set_batch([x] * B, [lang] * B, [prompt] * B)
eng.ar_prefill()
toks, lens, _ = eng.ar_decode(max_new=1275, top_k=1, ignore_eos=True)
t = eng.timing()
print("B", B, "us/step", 1e3 * t["ms_decode"] / (t["decode_steps"] - 1))
raw = eng.debug_read("decode_prof", 3 * 4096 * 2).view(np.int64).reshape(3, 2048, 2)
mhz = 1965.0
fused = os.environ.get("VX_DECODE_FUSED", "") != "0" and B <= (8 if os.environ.get("VX_DECODE_FUSED") == "1" else 2)
per_layer = ("qkv", "attn", "ffn") if fused else ("qkv", "attn", "out", "ff1", "ff2")   # fused: 38 barriers per step, else 62
if B == 1 and os.environ.get("VX_DECODE_CLUSTER", "") == "1":
    per_layer = ("PA", "PB")   # cluster kernel: 26 barriers per step; PA stamps: 3 staged, 4 q|k|v gathered, 5 head output folded, 6 out-proj + issues done
names = ["S0"] + [f"{n}" for l in range(12) for n in per_layer] + ["PF"]
for c in range(3):
    ev = [(int(tag), int(clk)) for tag, clk in raw[c] if tag != 0]
    seg = collections.defaultdict(list)
    phase = 0
    prev = None
    for tag, clk in ev:
        if prev is not None:
            key = (names[phase % len(names)], f"{prev[0]}->{tag}")
            seg[key].append((clk - prev[1]) / mhz)
        if tag == 2:
            phase += 1
        prev = (tag, clk)
    tot = sum(np.sum(v) for v in seg.values()) / 4.0
    print(f"CTA slot {c}: {len(ev)} stamps, {tot:.1f} us/step accounted")
    if c == 0 or os.environ.get("VX_TIMELINE_ALL") == "1":   # VX_TIMELINE_ALL=1: CTAs grid/2 and grid-1 too (skew between CTAs)
        agg = collections.defaultdict(list)
        for (nm, k), v in seg.items():
            agg[(nm, k)] += v
        for nm in ("S0",) + per_layer + ("PF",):
            parts = "  ".join(f"{k}:{np.mean(v):5.2f}" for (n2, k), v in sorted(agg.items()) if n2 == nm)
            tot_nm = sum(np.mean(v) for (n2, k), v in agg.items() if n2 == nm)
            print(f"  {nm:5s} {parts}   | {tot_nm:5.2f} us per phase")
