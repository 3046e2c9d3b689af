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
        B = a.batch or 32
        Lp = rng.integers(18, 161, size=B)
        P = rng.integers(161, 759, size=B)
        S = Lp + 512
        max_ctx = int((S + P).max()) + a.steps + 8
        eng = Engine(0, B, max_ctx, a.dtype)
        eng.load_valle_state_dict(synthetic_state_dict(0))
        eng.finalize()
        text = [rng.integers(5, 70, size=int(s)).astype(np.int32) for s in S]
        lang = [np.concatenate([np.full(int(lp), l % 3, np.int32), np.full(512, (l + 1) % 3, np.int32)]) for l, lp in enumerate(Lp)]
        prm = [rng.integers(0, 1024, size=(int(p), 8)).astype(np.int32) for p in P]
        res = []
        for _ in range(a.reps + 1):
            eng.set_batch(text, lang, prm)
            eng.ar_prefill()
            toks, lens, _ = eng.ar_decode(max_new=a.steps, top_k=1, ignore_eos=True)
            t = eng.timing()
            res.append(t)
        t = res[-1]
        fw = t["decode_steps"] - 1
        t_step = t["ms_decode"] / 1e3 / fw
        L0 = (S + P + 1).astype(np.float64)
        Lmean = L0 + fw / 2.0
        kvb = 49152 if a.dtype == "bf16" else 98304
        wb = 304412672 if a.dtype == "bf16" else 608825344
        bytes_step = wb + kvb * float((Lmean + 1).sum())
        ach = bytes_step / t_step / 1e9
        print(json.dumps({"config": 3, "B": B, "dtype": a.dtype, "decode_steps": t["decode_steps"], "ms_decode": t["ms_decode"],
                          "ms_prefill": t["ms_prefill"], "us_per_step": t_step * 1e6, "ar_tokens_per_s": B / t_step,
                          "mean_ctx": float(Lmean.mean()), "bytes_per_step": bytes_step,
                          "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "peak_source": how},
                          "lens_ok": bool((lens == np.minimum(a.steps, 16 * S)).all())}))
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
        B = a.batch or 8
        S, P, T = 314, 225, 1275
        eng = Engine(0, B, S + P + T + 8, a.dtype)
        eng.load_valle_state_dict(synthetic_state_dict(0))
        eng.load_vocos_state_dict(synthetic_vocos_state_dict(0))
        eng.finalize()
        text = [rng.integers(5, 70, size=S).astype(np.int32) for _ in range(B)]
        lang = [np.zeros(S, np.int32) for _ in range(B)]
        prm = [rng.integers(0, 1024, size=(P, 8)).astype(np.int32) for _ in range(B)]
        cb0 = rng.integers(0, 1024, size=(B, T)).astype(np.int32)
        tl = np.full((B,), T, np.int32)
        for _ in range(a.reps + 1):
            eng.set_batch(text, lang, prm)
            codes = eng.nar(cb0=cb0, t_lens=tl)
            t_nar = eng.timing()["ms_nar"]
            wave = eng.vocos_decode(np.ascontiguousarray(codes.transpose(2, 0, 1)))
            t_voc = eng.timing()["ms_vocos"]
        L = S + P + T
        flops = 7 * (2 * L * 150994944 + 49152 * L * L + 2 * T * 1048576) * B
        ach = flops / (t_nar / 1e3) / 1e12
        print(json.dumps({"config": 4, "B": B, "dtype": a.dtype, "ms_nar": t_nar, "ms_vocos": t_voc,
                          "frames_per_s": B * T / ((t_nar + t_voc) / 1e3), "nar_tflops": ach,
                          "roofline": {"bound": "tensor", "achieved": ach, "peak": tfl, "unit": "TFLOP/s", "frac": ach / tfl,
                                       "peak_source": how + " (bf16_tflops_sustained)"},
                          "vocos_frames_per_s": B * T / (t_voc / 1e3)}))


if __name__ == "__main__":
    main()
