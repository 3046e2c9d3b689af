"""Generate the committed golden fixtures by running the UNMODIFIED reference.

Run HERE (authoring container, /root/reference present):  python tests/golden/make_golden.py [case ...]

For every case the reference `models.vallex.VALLE` (loaded through oracle/ref_shim.py) is given a
synthetic checkpoint (`vallex_b200.weights.synthetic_state_dict(seed)`, strict load) and run greedy
(`top_k=1, temperature=1.0`) exactly like models/vallex.py:458-686.  Stored per case:
  inputs:  x [S], Lp, prompt [P,8], prompt_language, text_language (str or per-token list), seed
  outputs: codes [T,8] (fp32 weights), ar_top2 [T+1,2] / nar_top2 [7,T,2] (top-1/top-2 logit per step,
           captured with forward hooks on the predict layers -- the reference itself is untouched),
           codes_bf16w + margins: the same run with every matrix weight rounded to bf16 (the values the
           engine's bf16 mode holds), computed in fp32 by the reference.
Real preset prompts (inputs only) are read from /root/reference/presets/*.npz and embedded in the
fixture because the GPU box has no reference tree.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vall-e-x_b200"))

from oracle import ref_shim  # noqa: E402
from vallex_b200.weights import round_state_dict_bf16, synthetic_state_dict  # noqa: E402

OUT = os.path.dirname(__file__)
LANGS = ["en", "zh", "ja"]

CASES = {
    # name: (seed, Lp, Ln, P, prompt_language, text_language, preset)
    "tiny_en": (0, 2, 4, 5, "en", "en", None),
    "noprompt_en": (0, 0, 5, 0, "en", "en", None),
    "mixlang": (1, 3, 5, 12, "zh", ["en", "en", "zh", "ja", "ja"], None),
    "mid_ja": (2, 6, 10, 40, "ja", "ja", None),
    "cfg1_librispeech": (0, None, 64, None, "en", "en", "librispeech_1"),
    # context > 2048 keys and audio PE positions > 4000 (modules/embedding.py:68-91 auto-extension): a 3900-frame prompt,
    # 10 phonemes -> the reference's own 16*S cap ends the run at T = 160 (context 4071)
    "longctx_en": (0, 4, 6, 3900, "en", "en", None),
}


def build_inputs(name):
    seed, Lp, Ln, P, pl, tl, preset = CASES[name]
    rng = np.random.default_rng(1234 + seed)
    if preset is not None:
        d = np.load(os.path.join(ref_shim.REF_ROOT, "presets", preset + ".npz"))
        prompt = d["audio_tokens"][0].astype(np.int64)
        xp = d["text_tokens"][0].astype(np.int64)
        Lp, P = len(xp), prompt.shape[0]
    else:
        prompt = rng.integers(0, 1024, size=(P, 8), dtype=np.int64)
        xp = rng.integers(5, 70, size=(Lp,), dtype=np.int64)
    xn = rng.integers(5, 70, size=(Ln,), dtype=np.int64)
    x = np.concatenate([xp, xn])
    return seed, x, Lp, prompt, pl, tl


def run_reference(model, x, Lp, prompt, pl, tl):
    ar, nar = [], [[] for _ in range(7)]
    hooks = [model.ar_predict_layer.register_forward_hook(lambda m, i, o: ar.append(o.detach().clone()))]
    for j in range(7):
        hooks.append(model.nar_predict_layers[j].register_forward_hook(
            lambda m, i, o, j=j: nar[j].append(o.detach().clone())))
    with torch.no_grad():
        codes = model.inference(torch.from_numpy(x)[None], torch.tensor([len(x)]), torch.from_numpy(prompt)[None],
                                enroll_x_lens=Lp, top_k=1, temperature=1.0, prompt_language=pl,
                                text_language=tl, best_of=1)
    for h in hooks:
        h.remove()
    ar_top2 = torch.stack([torch.topk(l[0], 2)[0] for l in ar]).numpy()           # [T+1, 2]
    nar_top2 = torch.stack([torch.topk(n[0][0], 2, dim=-1)[0] for n in nar]).numpy()  # [7, T, 2]
    ar_logits_head = torch.stack([l[0] for l in ar[:4]]).numpy()                  # first 4 steps, full rows
    return codes[0].numpy(), ar_top2, nar_top2, ar_logits_head


def main(names):
    torch.set_num_threads(int(os.environ.get("GOLDEN_THREADS", os.cpu_count())))
    by_seed = {}
    for name in names:
        seed, x, Lp, prompt, pl, tl = build_inputs(name)
        if seed not in by_seed:
            by_seed[seed] = synthetic_state_dict(seed)
        sd = by_seed[seed]
        out = dict(x=x, Lp=np.int64(Lp), prompt=prompt, prompt_language=np.array(pl),
                   text_language=np.array(tl if isinstance(tl, list) else [tl]),
                   text_language_is_list=np.bool_(isinstance(tl, list)), seed=np.int64(seed))
        for tag, rounding in (("", None), ("_bf16w", "bf16")):
            sdr = sd
            if rounding:
                sdr = round_state_dict_bf16(sd)
            t0 = time.time()
            model = ref_shim.load_reference_valle(sdr)
            codes, ar_top2, nar_top2, head = run_reference(model, x, Lp, prompt, pl, tl)
            print(f"{name}{tag}: T={codes.shape[0]} in {time.time() - t0:.1f}s", flush=True)
            out["codes" + tag] = codes.astype(np.int16)
            out["ar_top2" + tag] = ar_top2.astype(np.float32)
            out["nar_top2" + tag] = nar_top2.astype(np.float32)
            if tag == "":
                out["ar_logits_head"] = head.astype(np.float32)
            del model
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def make_eos_cases():
    """True EOS termination (models/vallex.py:574): random-init weights never emit EOS, so row 1024 of
    ar_predict_layer.weight is crafted from the reference's own final hidden state at step t* (w = c * h / |h|^2 makes the
    EOS logit exactly c there) and the REFERENCE is re-run with the patched checkpoint; a candidate is kept only if the
    reference really stops at T == t*.  t* = 0 is the "first sample is EOS" edge (T = 0, NAR on an empty target)."""
    torch.set_num_threads(os.cpu_count())
    seed, x, Lp, prompt, pl, tl = build_inputs("tiny_en")
    sd = synthetic_state_dict(seed)
    model = ref_shim.load_reference_valle(sd)
    hs, lg = [], []
    h1 = model.ar_decoder.norm.register_forward_hook(lambda m, i, o: hs.append(o.detach()[0, -1].clone()))
    h2 = model.ar_predict_layer.register_forward_hook(lambda m, i, o: lg.append(o.detach()[0].clone()))
    with torch.no_grad():
        model.inference(torch.from_numpy(x)[None], torch.tensor([len(x)]), torch.from_numpy(prompt)[None], enroll_x_lens=Lp,
                        top_k=1, temperature=1.0, prompt_language=pl, text_language=tl)
    h1.remove(); h2.remove()
    for name, cands in (("eos_mid", [37, 23, 41, 29, 53]), ("eos_first", [0])):
        for t in cands:
            h = hs[t]
            c = float(lg[t].max()) + 1.0
            row = (c * h / (h * h).sum()).float()
            sd2 = dict(sd)
            w = sd["ar_predict_layer.weight"].clone()
            w[1024] = row
            sd2["ar_predict_layer.weight"] = w
            m2 = ref_shim.load_reference_valle(sd2)
            with torch.no_grad():
                codes = m2.inference(torch.from_numpy(x)[None], torch.tensor([len(x)]), torch.from_numpy(prompt)[None],
                                     enroll_x_lens=Lp, top_k=1, temperature=1.0, prompt_language=pl, text_language=tl)
            T = codes.shape[1]
            print(f"{name}: t*={t} -> reference T={T}", flush=True)
            if T == t:
                np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x, Lp=np.int64(Lp), prompt=prompt,
                                    prompt_language=np.array(pl), text_language=np.array([tl]),
                                    text_language_is_list=np.bool_(False), seed=np.int64(seed),
                                    eos_row=row.numpy().astype(np.float32), codes=codes[0].numpy().astype(np.int16))
                break
        else:
            raise SystemExit(f"no EOS candidate worked for {name}")


def make_cfg2_case(T_stop=1275):
    """BASELINE config 2 / the bench.py workload (S = 58 + 256, librispeech_1 prompt P = 225, T = 1275 so P + T = 1500).
    With random weights the unmodified reference would run to its 16*S = 5024 cap, so -- as in make_eos_cases -- row 1024
    of ar_predict_layer.weight is crafted from the reference's own hidden states so that EOS wins exactly at step 1275:
    pass 1 runs the reference and is cut short by a forward hook once the hidden state of step 1275 has been seen (the
    reference code is untouched); the row is the weighted least-squares solution that gives EOS the winning logit at that
    step and a low one at every earlier step; pass 2 is a COMPLETE run of the unmodified reference (AR stops by itself at T = 1275, then the 7 NAR stages) and
    is kept only if T == 1275.  Stored for fp32 weights and for bf16-rounded matrix weights (each with its own EOS row).
    Inputs are exactly bench.py's workload_inputs()."""
    torch.set_num_threads(int(os.environ.get("GOLDEN_THREADS", os.cpu_count())))
    d = np.load(os.path.join(ref_shim.REF_ROOT, "presets", "librispeech_1.npz"))
    prompt = d["audio_tokens"][0].astype(np.int64)
    xp = d["text_tokens"][0].astype(np.int64)
    xn = np.random.default_rng(2024).integers(5, 70, size=(256,), dtype=np.int64)
    x, Lp = np.concatenate([xp, xn]), len(xp)
    pl = tl = "en"
    sd = synthetic_state_dict(0)
    out = dict(x=x, Lp=np.int64(Lp), prompt=prompt, prompt_language=np.array(pl), text_language=np.array([tl]),
               text_language_is_list=np.bool_(False), seed=np.int64(0), T=np.int64(T_stop))

    class _Stop(Exception):
        pass

    for tag, sdr in (("", sd), ("_bf16w", round_state_dict_bf16(sd))):
        t0 = time.time()
        model = ref_shim.load_reference_valle(sdr)
        hs, lg = [], []

        def on_norm(m, i, o):
            hs.append(o.detach()[0, -1].clone())

        def on_pred(m, i, o):
            lg.append(o.detach()[0].clone())
            if len(lg) == T_stop + 1:
                raise _Stop()
        h1 = model.ar_decoder.norm.register_forward_hook(on_norm)
        h2 = model.ar_predict_layer.register_forward_hook(on_pred)
        try:
            with torch.no_grad():
                model.inference(torch.from_numpy(x)[None], torch.tensor([len(x)]), torch.from_numpy(prompt)[None],
                                enroll_x_lens=Lp, top_k=1, temperature=1.0, prompt_language=pl, text_language=tl)
        except _Stop:
            pass
        h1.remove(); h2.remove()
        print(f"cfg2{tag}: pass 1 ({len(lg)} steps) in {time.time() - t0:.0f}s", flush=True)
        H = torch.stack(hs)                                   # [T_stop + 1, 1024]
        hstar = H[T_stop]
        c = float(lg[T_stop].max()) + 1.0
        # EOS row: weighted least squares over the hidden states the reference itself produced -- logit c at step T_stop
        # (exact after the final rescale), about -2 at every earlier step (1276 equations, 1024 unknowns: small residuals)
        wts = torch.ones(T_stop + 1, dtype=torch.float64)
        wts[T_stop] = 1e3
        tgt = torch.full((T_stop + 1,), -2.0, dtype=torch.float64)
        tgt[T_stop] = c
        A = H.double() * wts[:, None]
        row = torch.linalg.lstsq(A, (tgt * wts)[:, None]).solution[:, 0].float()
        row = row * (c / float((row * hstar).sum()))
        if tag:   # the bf16 engine mode stores this matrix in bf16: the reference sees exactly that value
            row = row.to(torch.bfloat16).to(torch.float32)
        eos_all = H @ row
        top = torch.stack([l[:1024].max() for l in lg])
        assert bool((eos_all[:T_stop] < top[:T_stop]).all()), "crafted EOS row fires early"
        sd2 = dict(sdr)
        wmat = sdr["ar_predict_layer.weight"].clone()
        wmat[1024] = row
        sd2["ar_predict_layer.weight"] = wmat
        m2 = ref_shim.load_reference_valle(sd2)
        codes, ar_top2, nar_top2, head = run_reference(m2, x, Lp, prompt, pl, tl)
        print(f"cfg2{tag}: complete reference run T={codes.shape[0]} in {time.time() - t0:.0f}s", flush=True)
        assert codes.shape[0] == T_stop, codes.shape
        out["eos_row" + tag] = row.numpy().astype(np.float32)
        out["codes" + tag] = codes.astype(np.int16)
        # margins over the 1024 audio logits + EOS as the reference saw them (top-2 of the full 1025 row)
        out["ar_top2" + tag] = ar_top2.astype(np.float32)
        out["nar_top2" + tag] = nar_top2.astype(np.float32)
        del model, m2
    np.savez_compressed(os.path.join(OUT, "cfg2_librispeech.npz"), **out)


if __name__ == "__main__":
    if sys.argv[1:] == ["eos"]:
        make_eos_cases()
        sys.exit(0)
    if sys.argv[1:] == ["cfg2"]:
        make_cfg2_case()
        sys.exit(0)
    main(sys.argv[1:] or [n for n in CASES if n != "cfg1_librispeech"])
