import os, sys
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "vall-e-x_b200")):
    sys.path.insert(0, p)
from vallex_b200.engine import Engine
from vallex_b200.weights import synthetic_state_dict
dtype = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
max_ctx = int(sys.argv[4]) if len(sys.argv) > 4 else 256
ignore = int(sys.argv[5]) if len(sys.argv) > 5 else 1
eng = Engine(0, B, max_ctx, dtype)
eng.load_valle_state_dict(synthetic_state_dict(0))
eng.finalize()
rng = np.random.default_rng(0)
x = rng.integers(5, 70, size=6).astype(np.int32)
pr = rng.integers(0, 1024, size=(5, 8)).astype(np.int32)
eng.set_batch([x] * B, [np.zeros(6, np.int32)] * B, [pr] * B)
eng.ar_prefill()
toks, lens, _ = eng.ar_decode(max_new=steps, top_k=1, ignore_eos=bool(ignore))
print("ok", dtype, B, toks[:, -6:], lens)
