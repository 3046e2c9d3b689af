#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: codec-tokens/s (and RTF) of the
generate_audio hot path, BASELINE.json configs[1]: batch-1 AR + NAR + Vocos end to end, 256-phoneme text,
librispeech_1 prompt (Lp=58, P=225), bf16, T = 1275 generated frames (P+T = 1500; random-init weights never
emit EOS, so the length is forced with max_new_tokens exactly as SURVEY.md 8(d) prescribes).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one whole utterance through the hot path.  N>1: every rank serves its own utterances (request
sharding, no data-path collective; weak scaling); the checkpoint is generated on rank 0 and broadcast over NCCL.
One JSON line on rank 0.  `value` = frames/s with inputs resident in HBM (device pointers through the C ABI);
`e2e` = the same through the public Python API with host buffers (H2D of ids/prompt and D2H of codes + waveform
inside the timed region).  `roofline` is for the AR decode step (HBM bound): algorithmic bytes/step =
304,412,672 + 49,152 * (L+1) at bf16 over the CUDA-event duration of the decode loop.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "vall-e-x_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

LN_NEW = 256          # new phonemes (configs[1])
T_NEW = 1275          # forced generated frames: P + T = 1500
W_AR_BYTES = 304_412_672
KV_BYTES_PER_KEY = 49_152
METRIC = "codec_tokens_per_s"
UNIT = "AR codec tokens (frames)/s, end-to-end AR+NAR+Vocos"


def workload_inputs():
    """librispeech_1 prompt tokens come from the committed golden fixture (inputs of the reference run);
    the 256 new phoneme ids are seeded uniform over the 5..69 phoneme vocabulary."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "cfg1_librispeech.npz"))
    Lp = int(g["Lp"])
    xp = g["x"][:Lp].astype(np.int32)
    prompt = g["prompt"].astype(np.int32)
    xn = np.random.default_rng(2024).integers(5, 70, size=(LN_NEW,), dtype=np.int64).astype(np.int32)
    return np.concatenate([xp, xn]), Lp, prompt


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.idx = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()  # exact PID we started
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


_CPU_CACHE = {}


def _cpu_sample_once(o, ov, xt, yt, Lp, new_tokens):
    import torch
    t0 = time.perf_counter()
    codes = o.inference(xt, torch.tensor([xt.shape[1]]), yt, Lp, top_k=1, temperature=1.0, prompt_language="en",
                        text_language="en", max_new_tokens=new_tokens)
    t1 = time.perf_counter()
    ov.decode(ov.codes_to_features(codes.permute(2, 0, 1)), torch.tensor([2]))
    t2 = time.perf_counter()
    return int(codes.shape[1]), t1 - t0, t2 - t1


def cpu_reference_sample(new_tokens: int = 48, threads: int = 0) -> dict:
    """The reference's CPU PyTorch path, restated by the oracle port (oracle/vallex_oracle.py, pinned to the
    reference's goldens), on a bounded sample of the same workload: same prompt and 256-phoneme text, AR prefill +
    `new_tokens` greedy steps + the 7 NAR passes over those frames + the Vocos restatement."""
    import torch
    from oracle.vallex_oracle import OracleVALLE
    from oracle.vocos_oracle import OracleVocos
    from vallex_b200.weights import synthetic_state_dict, synthetic_vocos_state_dict
    ncpu = os.cpu_count() or 1
    x, Lp, prompt = workload_inputs()
    if "o" not in _CPU_CACHE:
        _CPU_CACHE["o"] = OracleVALLE(synthetic_state_dict(0))
        _CPU_CACHE["ov"] = OracleVocos(synthetic_vocos_state_dict(0))
    o, ov = _CPU_CACHE["o"], _CPU_CACHE["ov"]
    xt = torch.from_numpy(x.astype(np.int64))[None]
    yt = torch.from_numpy(prompt.astype(np.int64))[None]
    # The reference would run with torch's default (all cores); tiny per-token ops make that oversubscribed on a
    # many-core host, so a short probe picks the fastest thread count and the baseline is reported at ITS best.
    cands = [threads] if threads else sorted({min(ncpu, c) for c in (8, 16, 32, ncpu)})
    if "threads" in _CPU_CACHE and not threads:
        cands = [_CPU_CACHE["threads"]]
    best, best_rate = cands[0], -1.0
    if len(cands) > 1:
        for c in cands:
            torch.set_num_threads(c)
            T, ta, tv = _cpu_sample_once(o, ov, xt, yt, Lp, 4)
            if T / (ta + tv) > best_rate:
                best, best_rate = c, T / (ta + tv)
    torch.set_num_threads(best)
    _CPU_CACHE["threads"] = best
    T, ta, tv = _cpu_sample_once(o, ov, xt, yt, Lp, new_tokens)
    return {"value": T / (ta + tv), "unit": UNIT, "cores": best, "kind": "port",
            "sample": f"S=314 (Lp=58+256), P=225, AR prefill + {T} greedy steps + 7 NAR passes + Vocos on {T} frames "
                      f"(AR+NAR {ta:.2f}s, Vocos {tv:.2f}s); fp32 torch CPU, {best} of {ncpu} threads (fastest of {cands}); "
                      f"reference cannot batch (models/vallex.py:491)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-sample-tokens", type=int, default=48)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    ap.add_argument("--t-new", type=int, default=T_NEW, help="profiling runs only: shorter forced length")
    ap.add_argument("--min-warmup", type=int, default=3, help="profiling runs only (ncu replays every launch)")
    a = ap.parse_args()
    globals()["T_NEW"] = a.t_new
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != max(a.gpus, 1) and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")

    config = {"workload": "BASELINE.json configs[1]: batch-1 AR+NAR+Vocos end-to-end, 256-phoneme text, librispeech_1 prompt "
                          "(S=314, P=225), T=1275 forced (P+T=1500), greedy, one utterance per step per GPU",
              "S": 58 + LN_NEW, "P": 225, "T": T_NEW, "utterances_per_step_per_gpu": 1, "parallelism": f"request-sharded x{world}",
              "l2": "per-step working set (0.74 GB weights + KV) exceeds the 126 MB L2; no explicit flush needed"}

    if a.impl == "reference":
        if rank != 0:
            return
        W, K = max(a.warmup, 0), max(a.steps, 1)
        vals = []
        for i in range(min(W, 1) + K):  # each step = one bounded sample (~10-20 s); one warm-up is plenty on CPU
            r = cpu_reference_sample(a.cpu_sample_tokens)
            if i >= min(W, 1):
                vals.append(r)
        v = float(np.mean([r["value"] for r in vals]))
        cb = dict(vals[-1], value=v)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": K,
                          "warmup": W, "ms_per_step": 1e3 * T_NEW / v, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded random-init weights, seeded phonemes, librispeech_1 prompt tokens)",
                          "config": config, "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return

    import torch
    import torch.distributed as dist

    from vallex_b200 import _capi
    from vallex_b200.dist import broadcast_state_dict, max_over_ranks
    from vallex_b200.vallex import VALLE
    from vallex_b200.vocos import Vocos
    from vallex_b200.weights import synthetic_state_dict, synthetic_vocos_state_dict

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
        sd = broadcast_state_dict(synthetic_state_dict(0) if rank == 0 else None, src=0, device=dev)
        vsd = broadcast_state_dict(synthetic_vocos_state_dict(0) if rank == 0 else None, src=0, device=dev)
    else:
        sd, vsd = synthetic_state_dict(0), synthetic_vocos_state_dict(0)

    m = VALLE(1024, 16, 12, norm_first=True, add_prenet=False, prefix_mode=1, share_embedding=True, nar_scale_factor=1.0,
              prepend_bos=True, num_quantizers=8, dtype=a.dtype, max_batch=1, max_ctx=2048)
    m.load_state_dict(sd)
    m.to(dev).eval()
    voc = Vocos(vsd).to(dev).share_engine(m)
    eng = m.engine
    lib, h = eng.raw()
    x, Lp, prompt = workload_inputs()
    S, P = len(x), prompt.shape[0]
    lang = np.zeros((S,), np.int32)

    # ---- device-resident leg: raw C ABI with device pointers --------------------------------------------
    d_text = torch.from_numpy(x)[None].to(dev); d_lang = torch.from_numpy(lang)[None].to(dev)
    d_prompt = torch.from_numpy(prompt)[None].contiguous().to(dev)
    d_S = torch.tensor([S], dtype=torch.int32, device=dev); d_P = torch.tensor([P], dtype=torch.int32, device=dev)
    d_tok = torch.zeros((1, T_NEW), dtype=torch.int32, device=dev); d_len = torch.zeros((1,), dtype=torch.int32, device=dev)
    d_codes = torch.zeros((1, T_NEW, 8), dtype=torch.int32, device=dev)
    d_codes_t = torch.zeros((8, 1, T_NEW), dtype=torch.int32, device=dev)
    d_wave = torch.zeros((1, 320 * T_NEW), dtype=torch.float32, device=dev)
    ck, ptr, DEV = _capi.check, _capi.ptr, _capi.VX_DEVICE
    phase = {"prefill": [], "decode": [], "nar": [], "vocos": [], "steps": []}

    def device_step(record: bool):
        ck(lib.vx_set_batch(h, 1, ptr(d_text), ptr(d_S), ptr(d_lang), S, ptr(d_prompt), ptr(d_P), P, DEV, None), "set_batch")
        ck(lib.vx_ar_prefill(h, None), "prefill")
        ck(lib.vx_ar_decode(h, T_NEW, 1, 1.0, 1.0, 0, 1, None, ptr(d_tok), ptr(d_len), None, DEV, None), "decode")
        ck(lib.vx_nar(h, None, None, T_NEW, ptr(d_codes), DEV, None), "nar")
        d_codes_t.copy_(d_codes.permute(2, 0, 1))
        torch.cuda.current_stream().synchronize()
        ck(lib.vx_vocos_decode(h, ptr(d_codes_t), 1, T_NEW, 2, ptr(d_wave), DEV, None), "vocos")
        if record:
            t = eng.timing()
            for k in ("prefill", "decode", "nar", "vocos"):
                phase[k].append(t["ms_" + k])
            phase["steps"].append(t["decode_steps"])

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(a.warmup, a.min_warmup)):
        device_step(False)
    launches0 = eng.timing()["kernel_launches"]
    clocks = ClockSampler(local)
    sync_all()
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(a.steps):
        device_step(True)
    ev1.record()
    sync_all()
    wall = time.perf_counter() - t0
    assert int(d_len[0]) == T_NEW
    # the engine runs on its own stream and every call ends synchronised, so host wall time between the two
    # device-wide synchronisations brackets exactly the K steps; the per-phase numbers are CUDA-event times
    # taken on the engine's stream inside each call.
    dev_ms = float(np.sum(phase["prefill"]) + np.sum(phase["decode"]) + np.sum(phase["nar"]) + np.sum(phase["vocos"]))
    launches = eng.timing()["kernel_launches"] - launches0
    elapsed = max_over_ranks(wall, dev) if world > 1 else wall
    value = world * a.steps * T_NEW / elapsed

    # ---- end-to-end leg: public API, host buffers ---------------------------------------------------------
    xt = torch.from_numpy(x.astype(np.int64))[None]
    yt = torch.from_numpy(prompt.astype(np.int64))[None]
    pin = lambda t: t.pin_memory()
    xt, yt = pin(xt), pin(yt)

    def e2e_step():
        codes = m.inference(xt, torch.tensor([S]), yt, Lp, top_k=1, temperature=1.0, prompt_language="en",
                            text_language="en", max_new_tokens=T_NEW, quiet=True)
        frames = codes.permute(2, 0, 1)
        wave = voc.decode(voc.codes_to_features(frames), bandwidth_id=torch.tensor([2]))
        return codes, wave

    for _ in range(0 if a.no_e2e else 3):
        codes, wave = e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(0 if a.no_e2e else a.steps):
        codes, wave = e2e_step()
    sync_all()
    e2e_wall = max(time.perf_counter() - t0, 1e-9)
    e2e_elapsed = max_over_ranks(e2e_wall, dev) if world > 1 else e2e_wall
    e2e_value = world * a.steps * T_NEW / e2e_elapsed
    h2d = int(S * 4 * 2 + P * 8 * 4 + 8 + 8 * T_NEW * 4)   # ids, lang ids, prompt, lens; codes re-uploaded for Vocos
    d2h = int(T_NEW * 4 + 4 + 4 + T_NEW * 8 * 4 + 320 * T_NEW * 4)  # cb0 tokens, len, logprob, codes, waveform
    clk = clocks.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the AR decode (dominant phase, HBM bound) ---------------------------------------------
    steps = float(np.mean(phase["steps"]))
    t_step = float(np.mean(phase["decode"])) / 1e3 / max(steps - 1, 1)   # forwards = steps - 1 (last token needs none)
    L_mean = S + P + 1 + (steps - 1) / 2.0
    kvb = 1.0 if a.dtype == "bf16" else 2.0
    wbytes = W_AR_BYTES if a.dtype == "bf16" else 2 * W_AR_BYTES - 0
    bytes_step = wbytes + KV_BYTES_PER_KEY * kvb * (L_mean + 1)
    peak, how = peaks()
    ach = bytes_step / t_step / 1e9
    traffic = None  # measured DRAM bytes per decode step from the committed ncu --set full capture of this kernel
    tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.isfile(tp) and a.dtype == "bf16":
        traffic = json.load(open(tp))
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, a.min_warmup),
        "ms_per_step": 1e3 * elapsed / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": a.dtype, "data": "synthetic (seeded random-init weights, seeded phonemes; librispeech_1 prompt tokens from the reference's preset)",
        "config": config,
        "rtf": (elapsed / a.steps) / (T_NEW / 75.0),
        "ar_decode_tokens_per_s": 1.0 / t_step,
        "phase_ms": {k: float(np.mean(phase[k])) for k in ("prefill", "decode", "nar", "vocos")},
        "device_ms_per_step": dev_ms / a.steps,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_elapsed / a.steps, "rtf": (e2e_elapsed / a.steps) / (T_NEW / 75.0)},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"kernel": ("ar_decode_persistent_kernel" if os.environ.get("VX_DECODE_FUSED") == "0" or a.dtype != "bf16" else "ar_decode_fused_kernel")
                               + " (one launch = all decode steps; per-step figures)", "bound": "hbm",
                     "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": (traffic or {}).get("dram_bytes_per_step"),
                     "traffic_note": (None if traffic is None else f"per decode step, {traffic['capture']} ({traffic['source']}), kernel {traffic['kernel']}; algorithmic at that context: {traffic['algorithmic_bytes_per_step']:.0f} B"),
                     "peak_source": how, "bytes_per_step": bytes_step, "avg_step_us": t_step * 1e6, "mean_ctx": L_mean},
    }
    if not a.no_cpu_baseline and world == 1:   # the CPU baseline leg is timed on rank 0 at N=1 only
        out["cpu_baseline"] = cpu_reference_sample(a.cpu_sample_tokens)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
