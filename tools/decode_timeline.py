#!/usr/bin/env python
"""Per-phase timeline of the persistent decode kernel (VX_DECODE_PROF=<first step>): for CTAs 0 / grid/2 / grid-1 and 4
consecutive steps, clock64 at every grid barrier's arrival and release.  Prints mean work time (release of the previous
barrier -> arrival) and wait time (arrival -> release) per phase in microseconds."""
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
print("us/step", 1e3 * t["ms_decode"] / (t["decode_steps"] - 1))
raw = eng.debug_read("decode_prof", 3 * 4 * 128 * 2).view(np.int64).reshape(3, 4 * 128)
mhz = 1965.0
names = ["S0"] + [f"L{l}.{n}" for l in range(12) for n in ("qkv", "attn", "out", "ff1", "ff2")] + ["PF"]
for c in range(3):
    st = raw[c]
    n = int((st != 0).sum()) // 2
    arr, rel = st[0:2 * n:2], st[1:2 * n:2]
    work = (arr[1:] - rel[:-1]) / mhz
    wait = (rel - arr) / mhz
    per = 62
    print(f"CTA slot {c}: stamped barriers {n}; mean wait {wait.mean():.2f} us, mean work {work.mean():.2f} us")
    if c == 0:
        k = (n // per) * per
        w = work[:k - 1]
        for i, nm in enumerate(names):
            idx = [j for j in range(len(w)) if (j + 1) % per == i]
            widx = [j for j in range(k) if j % per == i]
            if i in (0, 1, 2, 3, 4, 5, 61) or i % 5 == 1 and i < 12:
                print(f"  {nm:8s} work {np.mean([w[j] for j in idx]) if idx else float('nan'):6.2f} us   wait {wait[widx].mean():6.2f} us")
        tot_work = w.reshape(-1)[:].sum() / (k / per)
        print(f"  per step: work {work[:k-1].sum() / (k / per):.1f} us, wait {wait[:k].sum() / (k / per):.1f} us")
