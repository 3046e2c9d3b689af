#!/usr/bin/env python
"""Vocos decode alone (B utterances x T frames) -- for `ncu --metrics gpu__time_duration.sum` launch lists."""
import os, sys
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "vall-e-x_b200")):
    sys.path.insert(0, p)
from vallex_b200.engine import Engine
from vallex_b200.weights import synthetic_vocos_state_dict
B, T = int(sys.argv[1]), int(sys.argv[2])
eng = Engine(0, 1, 64, "bf16")
eng.load_vocos_state_dict(synthetic_vocos_state_dict(0))
eng.finalize()
codes = np.random.default_rng(0).integers(0, 1024, size=(8, B, T)).astype(np.int32)
for _ in range(2):
    eng.vocos_decode(codes)
    print("ms_vocos", eng.timing()["ms_vocos"])
