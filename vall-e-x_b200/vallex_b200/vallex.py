"""Drop-in for the reference's `models.vallex.VALLE` inference surface (models/vallex.py:400-686),
backed by the B200 engine.  Same constructor call (utils/generation.py:67-78), `load_state_dict`
with the reference's 374 checkpoint keys, `.eval()`, `.to(device)` and
`.inference(x, x_lens, y, enroll_x_lens, top_k, temperature, prompt_language, text_language,
best_of, length_penalty, return_worst) -> LongTensor[1,T,8]`, same asserts and error types.
`forward` raises NotImplementedError exactly like the reference (models/vallex.py:447-457).
"""
from __future__ import annotations

from collections import namedtuple
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from .engine import Engine
from .macros import LANGUAGE_ID, N_DIM, NUM_AUDIO_TOKENS, NUM_HEAD, NUM_LAYERS, NUM_QUANTIZERS, NUM_TEXT_TOKENS

_IncompatibleKeys = namedtuple("_IncompatibleKeys", ["missing_keys", "unexpected_keys"])


def _device_index(device) -> int:
    d = torch.device(device) if not isinstance(device, torch.device) else device
    if d.type != "cuda":
        raise RuntimeError("vallex_b200 runs on CUDA (B200, sm_100a) only; there is no CPU path")
    return 0 if d.index is None else d.index


def _resolve_seed(seed: Optional[int]) -> int:
    if seed is None:
        return int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
    return int(seed)


def _check_ids(text: np.ndarray, prompt: np.ndarray):
    """The reference's embedding lookups raise IndexError on an out-of-range id (F.embedding; models/vallex.py:497,
    529,605-632); the engine indexes its tables with these ids, so the same error is raised before anything runs."""
    if text.size and (int(text.min()) < 0 or int(text.max()) >= NUM_TEXT_TOKENS):
        raise IndexError("index out of range in self (text token id outside the 2048-entry text embedding)")
    if prompt.size and (int(prompt.min()) < 0 or int(prompt.max()) >= NUM_AUDIO_TOKENS):
        raise IndexError("index out of range in self (audio prompt code outside 0..1023)")


class VALLE:
    def __init__(self, d_model: int, nhead: int, num_layers: int, norm_first: bool = True, add_prenet: bool = False,
                 prefix_mode: int = 0, share_embedding: bool = True, nar_scale_factor: float = 1.0,
                 prepend_bos: bool = True, num_quantizers: int = 8, *, dtype: str = "bf16", max_batch: int = 8,
                 max_ctx: int = 6144, **kwargs):
        cfg = (d_model, nhead, num_layers, norm_first, add_prenet, prefix_mode, share_embedding, nar_scale_factor,
               prepend_bos, num_quantizers)
        if cfg != (N_DIM, NUM_HEAD, NUM_LAYERS, True, False, 1, True, 1.0, True, NUM_QUANTIZERS):
            raise NotImplementedError(
                "the B200 engine implements the released VALL-E X configuration only "
                "(utils/generation.py:67-78): VALLE(1024, 16, 12, norm_first=True, add_prenet=False, "
                f"prefix_mode=1, share_embedding=True, nar_scale_factor=1.0, prepend_bos=True, num_quantizers=8); got {cfg}")
        self.language_ID = dict(LANGUAGE_ID)
        self.num_quantizers = num_quantizers
        self.prefix_mode = prefix_mode
        self._dtype, self._max_batch, self._max_ctx = dtype, max_batch, max_ctx
        self._device = 0
        self._sd: Optional[Dict[str, torch.Tensor]] = None
        self._engine: Optional[Engine] = None
        self.last_timing: dict = {}

    # -- nn.Module-like surface ----------------------------------------------------------------
    def load_state_dict(self, state_dict, strict: bool = True):
        """torch.nn.Module.load_state_dict semantics (utils/generation.py:79-83 calls it with strict=True and asserts
        `not missing_keys`): returns (missing_keys, unexpected_keys); with strict=True any missing / unexpected key or a
        shape mismatch raises RuntimeError right here, not at first engine use."""
        from .weights import expected_state_dict_shapes
        want = expected_state_dict_shapes()
        missing = [k for k in want if k not in state_dict]
        unexpected = [k for k in state_dict if k not in want]
        errors = []
        for k, shp in want.items():
            if k in state_dict and tuple(state_dict[k].shape) != shp:
                errors.append(f"size mismatch for {k}: copying a param with shape {tuple(state_dict[k].shape)} from checkpoint, "
                              f"the shape in current model is {shp}.")
        if strict:
            if unexpected:
                errors.insert(0, "Unexpected key(s) in state_dict: " + ", ".join(f'"{k}"' for k in unexpected) + ". ")
            if missing:
                errors.insert(0, "Missing key(s) in state_dict: " + ", ".join(f'"{k}"' for k in missing) + ". ")
        if errors:
            raise RuntimeError("Error(s) in loading state_dict for VALLE:\n\t" + "\n\t".join(errors))
        if missing and self._sd is None:
            raise RuntimeError("load_state_dict(strict=False) with missing keys needs a previously loaded checkpoint")
        base = dict(self._sd) if (missing and self._sd is not None) else {}
        base.update({k: v for k, v in state_dict.items() if k in want})
        self._sd = base
        if self._engine is not None:
            self._engine.close()
            self._engine = None
        return _IncompatibleKeys(missing, unexpected)

    def eval(self):
        return self

    def to(self, device):
        idx = _device_index(device)
        if self._engine is not None and idx != self._device:
            self._engine.close()
            self._engine = None
        self._device = idx
        return self

    def cuda(self, device=0):
        return self.to(torch.device("cuda", device))

    def forward(self, *args, **kwargs):
        raise NotImplementedError

    __call__ = forward

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            if self._sd is None:
                raise RuntimeError("load_state_dict() must be called before inference")
            eng = Engine(self._device, self._max_batch, self._max_ctx, self._dtype)
            eng.load_valle_state_dict(self._sd)  # strict: vx_finalize_weights rejects missing keys / bad shapes
            vs = getattr(self, "_vocos_sd", None)
            if vs is not None:
                eng.load_vocos_state_dict(vs)
            eng.finalize()
            self._engine = eng
        return self._engine

    def attach_vocos(self, vocos_state_dict):
        """Share this model's engine with the Vocos decoder (one engine per GPU)."""
        self._vocos_sd = vocos_state_dict
        if self._engine is not None:
            self._engine.load_vocos_state_dict(vocos_state_dict)
            self._engine.finalize()

    # -- inference -----------------------------------------------------------------------------
    def _lang_ids(self, S: int, Lp: int, prompt_language: str, text_language: Union[str, Sequence[str]]) -> np.ndarray:
        ids = np.empty((S,), np.int32)
        ids[:Lp] = self.language_ID[prompt_language]
        if isinstance(text_language, str):
            ids[Lp:] = self.language_ID[text_language]
        else:
            tl = [self.language_ID[t] for t in text_language]
            if len(tl) == 1:
                ids[Lp:] = tl[0]
            else:
                if len(tl) != S - Lp:  # the reference's broadcast would raise here too (models/vallex.py:505)
                    raise RuntimeError(f"text_language has {len(tl)} entries for {S - Lp} text tokens")
                ids[Lp:] = tl
        return ids

    def inference(self, x: torch.Tensor, x_lens: torch.Tensor, y: torch.Tensor, enroll_x_lens, top_k: int = -100,
                  temperature: float = 1.0, prompt_language: str = None, text_language=None, best_of: int = 1,
                  length_penalty: float = 1.0, return_worst: bool = False, *, max_new_tokens: int = 0,
                  seed: Optional[int] = None, quiet: bool = False) -> torch.Tensor:
        """models/vallex.py:458-686.  `max_new_tokens`, `seed`, `quiet` are engine extensions (keyword-only).
        `seed=None` draws the device sampler's Philox seed from torch's global CPU generator, so `torch.manual_seed`
        governs reproducibility of sampled output as it does for the reference."""
        assert x.ndim == 2, x.shape
        assert x_lens.ndim == 1, x_lens.shape
        assert y.ndim == 3, y.shape
        assert y.shape[0] == 1, y.shape
        assert torch.all(x_lens > 0)
        S = int(x_lens.max())
        Lp = int(enroll_x_lens)
        P = int(y.shape[1])
        text = x[0, :S].detach().to("cpu").numpy().astype(np.int32)
        prompt = y[0].detach().to("cpu").numpy().astype(np.int32).reshape(P, NUM_QUANTIZERS)
        lang = self._lang_ids(S, Lp, prompt_language, text_language)
        _check_ids(text, prompt)
        eng = self.engine
        seed = _resolve_seed(seed)
        nb = int(best_of)
        if nb > self._max_batch:
            raise ValueError(f"best_of={nb} exceeds the engine's max_batch={self._max_batch}")
        eng.set_batch([text], [lang], [prompt])
        eng.ar_prefill()                      # ONE prefill; the reference recomputes it best_of times (models/vallex.py:525-527)
        t_pre = eng.timing()["ms_prefill"]
        if nb > 1:
            eng.ar_fork(nb)                   # KV rows / token buffer / first logits replicated on the device
        toks, lens, slp = eng.ar_decode(max_new=max_new_tokens, top_k=top_k, top_p=1.0, temperature=temperature, seed=seed)
        if int(lens.max()) == 0 and P + 1 == P:  # unreachable with prepend_bos (models/vallex.py:579-582)
            raise SyntaxError("well trained model shouldn't reach here.")
        if eng.ctx_clamped() and int(lens.max()) >= self._max_ctx - (S + P + 1) - 1:
            # the reference has no context limit (it would run on to EOS or its 16*S cap, models/vallex.py:574-578): never
            # hand back a silently truncated utterance
            raise RuntimeError(f"KV cache exhausted: the utterance filled max_ctx={self._max_ctx} rows (S={S}, P={P}, {int(lens.max())} new "
                               f"frames) before EOS / the 16*S cap; build VALLE(..., max_ctx=...) larger or pass max_new_tokens")
        # best / worst beam by sum_logprobs / len**length_penalty (models/vallex.py:583-594), chosen on the device; the
        # NAR stages run on the chosen beam only (batch 1), like the reference.  (At the 16*S cap the reference draws one
        # more sample whose logprob enters the ranking before it breaks, models/vallex.py:569-578; the engine stops at the
        # cap without that extra forward pass -- only the ranking of beams that ALL hit the cap can differ.)
        pick = eng.ar_select_beam(length_penalty, return_worst) if nb > 1 else 0
        self.last_beam = dict(pick=pick, lens=lens.copy(), sum_logprobs=slp.copy())
        T = int(lens[pick])
        if not quiet:
            print(f"VALL-E EOS [{P} -> {P + 1 + T}]")
        codes = eng.nar(Tmax=T)
        self.last_timing = dict(eng.timing(), ms_prefill=t_pre)
        out = torch.from_numpy(codes[:1, :T].astype(np.int64))
        return out.to(x.device) if x.is_cuda else out

    def inference_batch(self, items, top_k: int = -100, temperature: float = 1.0, *, max_new_tokens: int = 0,
                        seed: Optional[int] = None, quiet: bool = False, continuous: Optional[bool] = None) -> List[torch.Tensor]:
        """Engine extension (the reference is batch-1 only, models/vallex.py:491): many independent
        utterances through `max_batch` decode slots.  `items` = sequence of `(x [1,S], y [1,P,8], enroll_x_lens,
        prompt_language, text_language)` with the meaning of `.inference`'s arguments; returns one LongTensor[1,T_b,8] per item,
        in input order.  Every utterance keeps its own EOS / 16*S cap and (sampling mode) its own Philox stream, so the result
        is a vmap of `.inference` calls.
        AR decode is CONTINUOUSLY batched when there are more items than slots (`continuous`, default on): the persistent
        decode kernel hands a slot back as soon as its utterance finishes (vx_ar_decode_ex stop_on_finish), the next pending
        utterance is prefilled alone into that slot (vx_slot_prefill) and the batch resumes -- nobody waits for the slowest
        utterance of a chunk.  The NAR stages then run over the finished utterances in ragged batches of `max_batch`."""
        items = list(items)
        if not items:
            return []
        seed = _resolve_seed(seed)
        eng = self.engine
        MB = self._max_batch
        prepared = []
        for x, y, enroll_x_lens, prompt_language, text_language in items:
            assert x.ndim == 2 and x.shape[0] == 1, x.shape
            assert y.ndim == 3 and y.shape[0] == 1, y.shape
            S, P = int(x.shape[1]), int(y.shape[1])
            assert S > 0
            text = x[0].detach().to("cpu").numpy().astype(np.int32)
            prompt = y[0].detach().to("cpu").numpy().astype(np.int32).reshape(P, NUM_QUANTIZERS)
            _check_ids(text, prompt)
            prepared.append((text, self._lang_ids(S, int(enroll_x_lens), prompt_language, text_language), prompt))
        if continuous is None:
            continuous = len(items) > MB
        n = len(items)
        cb0: List[Optional[np.ndarray]] = [None] * n
        timing = {"ms_prefill": 0.0, "ms_decode": 0.0, "ms_nar": 0.0, "decode_steps": 0, "decode_launches": 0}

        def check_room(i, T):
            text, _, prompt = prepared[i]
            if T >= self._max_ctx - (len(text) + prompt.shape[0] + 1) - 1 and (not max_new_tokens or T < max_new_tokens) and T < 16 * len(text):
                raise RuntimeError(f"KV cache exhausted for utterance {i}: max_ctx={self._max_ctx} rows filled before EOS / the 16*S cap; "
                                   f"build VALLE(..., max_ctx=...) larger or pass max_new_tokens")

        # ---- AR decode --------------------------------------------------------------------------------------------------
        if continuous:
            Smax = max(len(t) for t, _, _ in prepared)
            Pmax = max(p.shape[0] for _, _, p in prepared)
            first = list(range(min(MB, n)))
            nxt = len(first)
            slot_item = list(first)                      # item index held by each slot (-1: idle)
            acc = [[] for _ in first]                    # tokens gathered so far for the slot's current utterance
            eng.set_batch([prepared[i][0] for i in first], [prepared[i][1] for i in first], [prepared[i][2] for i in first],
                          Smax=Smax, Pmax=Pmax)
            eng.ar_prefill()
            timing["ms_prefill"] += eng.timing()["ms_prefill"]
            resume, launch = False, 0
            have = [0] * len(first)                      # tokens already copied out per slot
            while any(i >= 0 for i in slot_item):
                toks, lens, _, fin = eng.ar_decode_ex(max_new=max_new_tokens, top_k=top_k, top_p=1.0, temperature=temperature,
                                                      seed=seed + 7919 * launch, resume=resume, stop_on_finish=nxt < n)
                t = eng.timing()
                timing["ms_decode"] += t["ms_decode"]; timing["decode_steps"] += t["decode_steps"]; timing["decode_launches"] += 1
                resume, launch = True, launch + 1
                for s_, i in enumerate(slot_item):
                    if i < 0:
                        continue
                    T = int(lens[s_])
                    acc[s_].append(toks[s_, have[s_]:T].copy())      # the device buffer is cleared every launch
                    have[s_] = T
                    if fin[s_]:
                        cb0[i] = np.concatenate(acc[s_]) if acc[s_] else np.zeros((0,), np.int32)
                        check_room(i, T)
                        if nxt < n:                                   # refill the slot
                            eng.slot_prefill(s_, *prepared[nxt])
                            slot_item[s_], acc[s_], have[s_] = nxt, [], 0
                            nxt += 1
                        else:
                            slot_item[s_] = -1
                timing["ms_prefill"] = eng.timing()["ms_prefill"] if False else timing["ms_prefill"]
        else:
            for c0 in range(0, n, MB):
                idx = list(range(c0, min(c0 + MB, n)))
                eng.set_batch([prepared[i][0] for i in idx], [prepared[i][1] for i in idx], [prepared[i][2] for i in idx])
                eng.ar_prefill()
                timing["ms_prefill"] += eng.timing()["ms_prefill"]
                toks, lens, _ = eng.ar_decode(max_new=max_new_tokens, top_k=top_k, top_p=1.0, temperature=temperature, seed=seed + c0)
                t = eng.timing()
                timing["ms_decode"] += t["ms_decode"]; timing["decode_steps"] += t["decode_steps"]; timing["decode_launches"] += 1
                for b, i in enumerate(idx):
                    cb0[i] = toks[b, :int(lens[b])].copy()
                    check_room(i, int(lens[b]))
        # ---- NAR: the finished utterances in ragged batches -----------------------------------------------------------------
        outs: List[Optional[torch.Tensor]] = [None] * n
        for c0 in range(0, n, MB):
            idx = list(range(c0, min(c0 + MB, n)))
            Tm = max(int(len(cb0[i])) for i in idx)
            tok = np.zeros((len(idx), max(Tm, 1)), np.int32)
            for b, i in enumerate(idx):
                tok[b, :len(cb0[i])] = cb0[i]
            eng.set_batch([prepared[i][0] for i in idx], [prepared[i][1] for i in idx], [prepared[i][2] for i in idx])
            codes = eng.nar(cb0=tok[:, :Tm], t_lens=np.asarray([len(cb0[i]) for i in idx], np.int32))
            timing["ms_nar"] += eng.timing()["ms_nar"]
            for b, i in enumerate(idx):
                T = int(len(cb0[i]))
                if not quiet:
                    P = int(prepared[i][2].shape[0])
                    print(f"VALL-E EOS [{P} -> {P + 1 + T}]")
                out = torch.from_numpy(codes[b:b + 1, :T].astype(np.int64))
                x = items[i][0]
                outs[i] = out.to(x.device) if x.is_cuda else out
        self.last_timing = timing
        return outs
