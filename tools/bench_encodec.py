#!/usr/bin/env python
"""Prompt making: EnCodec encoder + RVQ of one 15-second mono 24 kHz clip (the reference's longest allowed
prompt, utils/prompt_making.py:60-62) through the public mirror `AudioTokenizer.encode` (host waveform in, codes out).
Prints one JSON line: GPU ms per clip (CUDA events inside the C-ABI call and wall clock around the Python call), frames/s,
and -- with --cpu -- the oracle port on the host cores for comparison (test infrastructure, never the product path)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "vall-e-x_b200"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=15.0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args()
    import torch
    from make_encodec_golden import synth_wave
    from vallex_b200.tokenizer import AudioTokenizer
    from vallex_b200.weights import synthetic_encodec_state_dict
    sd = synthetic_encodec_state_dict(0)
    n = int(a.seconds * 24000)
    wav = torch.from_numpy(synth_wave(np.random.default_rng(5), n)).view(1, 1, n)
    tok = AudioTokenizer("cuda:0", state_dict=sd)
    tok.encode(wav)
    wall, dev = [], []
    for _ in range(a.reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        codes = tok.encode(wav)[0][0]
        torch.cuda.synchronize()
        wall.append(time.perf_counter() - t0)
        dev.append(tok.last_timing()["ms"])
    T = codes.shape[-1]
    out = {"workload": f"EnCodec 24 kHz encoder + 8-level RVQ, one {a.seconds:g} s mono clip ({n} samples -> {T} frames)",
           "device_ms": min(dev), "e2e_ms": 1e3 * min(wall), "frames_per_s": T / min(wall), "x_realtime": a.seconds / min(wall),
           "kernel_launches": tok.last_timing()["kernel_launches"], "dtype": "f32"}
    if a.cpu:
        from oracle.encodec_oracle import OracleEncodec
        ora = OracleEncodec(sd)
        t0 = time.perf_counter()
        want = ora.encode(wav)[0][0]
        out["cpu_oracle_s"] = time.perf_counter() - t0
        out["cpu_threads"] = torch.get_num_threads()
        out["codes_equal_frac"] = float((want == codes.cpu()).float().mean())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
