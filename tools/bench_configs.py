#!/usr/bin/env python
"""Secondary measurements for BASELINE.json configs 3 and 4 (not the driver's bench line; results are kept under
profiles/ as evidence for the north-star roofline targets).

  --config 3 : batch-32 AR decode only, 512-phoneme synthetic text, mixed en/zh/ja prompts (Lp 18..160, P 161..758),
               forced steps with EOS ignored (context sweeps ~0.7k -> ~2.9k)       -> HBM roofline fraction
  --config 4 : NAR 7-codebook pass + Vocos, 1500-frame sequences (S=314, P=225, T=1275), batch B on one GPU
                                                                                  -> tensor-pipe fraction
  --config 5 : long text through the PUBLIC API (generation.generate_audio_from_long_text, mode='fixed-prompt'):
               one request = 16 synthetic "sentences" of 250 phoneme ids on one prompt (one request stream per GPU in
               the 8-GPU setting), every sentence capped at --t-new frames; batched (one ragged batch per max_batch
               sentences) against the reference's sentence-by-sentence order      -> frames/s, RTF per request
Prompts are synthetic with the preset shapes' ranges (the real presets live in the reference tree, which does not
exist on the GPU box)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "vall-e-x_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d["bf16_tflops_sustained"]), "measured"
    return 6650.0, 1400.0, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[3, 4, 5])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--steps", type=int, default=1500, help="config 3: forced decode steps")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--t-new", type=int, default=750, help="config 5: frames per sentence (random weights never emit EOS)")
    ap.add_argument("--sentences", type=int, default=16, help="config 5")
    a = ap.parse_args()
    import torch
    from vallex_b200.engine import Engine
    from vallex_b200.weights import synthetic_state_dict, synthetic_vocos_state_dict
    rng = np.random.default_rng(7)
    hbm, tfl, how = peaks()
    if a.config == 3:
        import bench
        print(json.dumps(dict(bench.cfg3_result(a.batch or 32, a.steps, a.dtype, reps=a.reps), config=3)))
    elif a.config == 5:
        from vallex_b200 import generation as g
        from vallex_b200.vallex import VALLE
        from vallex_b200.vocos import Vocos
        B = a.batch or 8                     # sentences decoded together
        Lp, P, Ln = 58, 225, 250
        m = VALLE(1024, 16, 12, norm_first=True, add_prenet=False, prefix_mode=1, share_embedding=True, nar_scale_factor=1.0,
                  prepend_bos=True, num_quantizers=8, dtype=a.dtype, max_batch=B, max_ctx=Lp + Ln + P + a.t_new + 8)
        m.load_state_dict(synthetic_state_dict(0))
        m.to("cuda:0").eval()
        voc = Vocos(synthetic_vocos_state_dict(0)).to("cuda:0").share_engine(m)
        g.model, g.vocos = m, voc
        import tempfile
        d = tempfile.mkdtemp()
        np.savez(os.path.join(d, "p.npz"), audio_tokens=rng.integers(0, 1024, size=(1, P, 8)).astype(np.int64),
                 text_tokens=rng.integers(5, 70, size=(1, Lp)).astype(np.int64), lang_code=np.int32(2))
        sents = [rng.integers(5, 70, size=Ln).astype(np.int64) for _ in range(a.sentences)]
        out = {}
        for name, kw in (("batched", {}), ("sequential", {"batch_sentences": False})):
            best = 1e30
            for _ in range(a.reps + 1):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                wav = g.generate_audio_from_long_text(sents, prompt=os.path.join(d, "p.npz"), language="en", mode="fixed-prompt",
                                                      top_k=1, max_new_tokens=a.t_new, quiet=True, **kw)
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            frames = wav.shape[0] / 320
            out[name] = {"s_per_request": best, "frames": frames, "frames_per_s": frames / best, "rtf": best / (frames / 75.0)}
        print(json.dumps({"config": 5, "sentences": a.sentences, "max_batch": B, "t_new": a.t_new, "dtype": a.dtype,
                          "api": "generation.generate_audio_from_long_text(mode='fixed-prompt')", **out,
                          "speedup_batched": out["sequential"]["s_per_request"] / out["batched"]["s_per_request"]}))
    else:
        import bench
        print(json.dumps(dict(bench.cfg4_result(a.batch or 8, a.dtype, reps=a.reps), config=4)))


if __name__ == "__main__":
    main()
