"""Python handle on one vx_engine (one per GPU).  Marshals numpy/torch buffers into the C ABI
(include/vallex_b200.h); holds no model arithmetic of its own."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _capi
from ._capi import VX_BF16, VX_DEVICE, VX_F32, VX_HOST, check, ptr
from .macros import NUM_QUANTIZERS
from .weights import sine_position_table

_DTYPES = {"fp32": VX_F32, "f32": VX_F32, "float32": VX_F32, "bf16": VX_BF16, "bfloat16": VX_BF16}


def _np_f32(t) -> np.ndarray:
    if hasattr(t, "detach"):
        t = t.detach().to("cpu").float().contiguous().numpy()
    return np.ascontiguousarray(t, dtype=np.float32)


class Engine:
    def __init__(self, device: int = 0, max_batch: int = 1, max_ctx: int = 4096, dtype: str = "bf16"):
        self._lib = _capi.lib()
        if dtype not in _DTYPES:
            raise ValueError(f"dtype must be one of {sorted(_DTYPES)}")
        self.dtype = "fp32" if _DTYPES[dtype] == VX_F32 else "bf16"
        self.device, self.max_batch, self.max_ctx = int(device), int(max_batch), int(max_ctx)
        h = C.c_void_p()
        check(self._lib.vx_create(self.device, self.max_batch, self.max_ctx, _DTYPES[dtype], C.byref(h)), "vx_create")
        self._h = h
        self.B = 0
        self.S: List[int] = []
        self.P: List[int] = []
        self.has_valle = False
        self.has_vocos = False

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights -------------------------------------------------------------------------
    def _load(self, key: str, value):
        if getattr(value, "is_cuda", False) and value.device.index == self.device:
            # already on this engine's GPU (e.g. the NCCL receive buffer of dist.broadcast_state_dict): device-to-device ingest
            t = value.detach().float().contiguous()
            if t.ndim == 0:
                t = t.reshape(1)
            shape = (C.c_int64 * t.ndim)(*t.shape)
            check(self._lib.vx_load_weight_mem(self._h, key.encode(), ptr(t), shape, t.ndim, VX_DEVICE), f"vx_load_weight_mem({key})")
            return
        a = _np_f32(value)
        if a.ndim == 0:
            a = a.reshape(1)
        shape = (C.c_int64 * a.ndim)(*a.shape)
        check(self._lib.vx_load_weight(self._h, key.encode(), ptr(a), shape, a.ndim), f"vx_load_weight({key})")

    def load_valle_state_dict(self, sd: Dict[str, "object"]):
        for k, v in sd.items():
            self._load(k, v)
        self._load("pe_table", sine_position_table(max(self.max_ctx + 8, 4000)))
        self.has_valle = True

    def load_vocos_state_dict(self, sd: Dict[str, "object"]):
        for k, v in sd.items():
            if k.startswith("feature_extractor.encodec."):
                continue  # the EnCodec model itself is not on the decode path
            self._load("vocos." + k, v)
        self.has_vocos = True

    def finalize(self):
        check(self._lib.vx_finalize_weights(self._h), "vx_finalize_weights")

    # ---- batch ---------------------------------------------------------------------------
    def set_batch(self, text_ids: Sequence[np.ndarray], lang_ids: Sequence[np.ndarray],
                  prompts: Sequence[np.ndarray], Smax: int = 0, Pmax: int = 0):
        """text_ids[b] int [S_b]; lang_ids[b] int [S_b] (model numbering en=0 zh=1 ja=2);
        prompts[b] int [P_b, 8].  Smax / Pmax: capacity to reserve per slot (continuous batching refills slots with
        other utterances through slot_prefill; default: the largest of this batch)."""
        B = len(text_ids)
        self.S = [int(len(t)) for t in text_ids]
        self.P = [int(p.shape[0]) for p in prompts]
        Smax, Pmax = max(max(self.S), int(Smax)), max(max(self.P), int(Pmax))
        txt = np.zeros((B, Smax), np.int32)
        lng = np.zeros((B, Smax), np.int32)
        prm = np.zeros((B, max(Pmax, 1), NUM_QUANTIZERS), np.int32)
        for b in range(B):
            txt[b, :self.S[b]] = text_ids[b]
            lng[b, :self.S[b]] = lang_ids[b]
            if self.P[b]:
                prm[b, :self.P[b]] = prompts[b]
        sl = np.asarray(self.S, np.int32)
        pl = np.asarray(self.P, np.int32)
        check(self._lib.vx_set_batch(self._h, B, ptr(txt), ptr(sl), ptr(lng), Smax, ptr(prm) if Pmax else None,
                                     ptr(pl), Pmax, VX_HOST, None), "vx_set_batch")
        self.B = B

    def ar_prefill(self):
        check(self._lib.vx_ar_prefill(self._h, None), "vx_ar_prefill")

    def ar_fork(self, n_beams: int):
        """best_of beams from ONE prefill (models/vallex.py:525-527): replicate utterance 0's decode state on the device."""
        check(self._lib.vx_ar_fork(self._h, int(n_beams), None), "vx_ar_fork")
        self.B = int(n_beams)
        self.S = self.S[:1] * self.B
        self.P = self.P[:1] * self.B

    def ar_select_beam(self, length_penalty: float = 1.0, return_worst: bool = False) -> int:
        """models/vallex.py:583-594 on the device; the chosen beam becomes the batch's only utterance."""
        pick = C.c_int()
        check(self._lib.vx_ar_select_beam(self._h, float(length_penalty), int(bool(return_worst)), C.byref(pick), None),
              "vx_ar_select_beam")
        self.B = 1
        self.S, self.P = self.S[:1], self.P[:1]
        return pick.value

    def ar_budget(self, max_new: int = 0) -> int:
        n = C.c_int()
        check(self._lib.vx_ar_budget(self._h, int(max_new), C.byref(n)), "vx_ar_budget")
        return n.value

    def slot_prefill(self, slot: int, text_ids: np.ndarray, lang_ids: np.ndarray, prompt: np.ndarray):
        """Continuous batching: put a new utterance into a finished slot (prefilled alone into the slot's KV region)."""
        t = np.ascontiguousarray(text_ids, np.int32)
        l = np.ascontiguousarray(lang_ids, np.int32)
        pr = np.ascontiguousarray(prompt, np.int32).reshape(-1, NUM_QUANTIZERS)
        check(self._lib.vx_slot_prefill(self._h, int(slot), ptr(t), int(len(t)), ptr(l), ptr(pr) if pr.shape[0] else None,
                                        int(pr.shape[0]), VX_HOST, None), "vx_slot_prefill")
        self.S[slot], self.P[slot] = int(len(t)), int(pr.shape[0])

    def ar_decode_ex(self, max_new: int = 0, top_k: int = -100, top_p: float = 1.0, temperature: float = 1.0, seed: int = 0,
                     resume: bool = False, stop_on_finish: bool = False):
        """vx_ar_decode_ex: returns (tokens [B,max_steps], lens [B], sum_logprobs [B], finished [B]) of every slot's CURRENT
        utterance; stop_on_finish returns as soon as one more slot is done, resume continues the paused batch."""
        steps = self.ar_budget(max_new)
        toks = np.zeros((self.B, steps), np.int32)
        lens = np.zeros((self.B,), np.int32)
        slp = np.zeros((self.B,), np.float32)
        fin = np.zeros((self.B,), np.int32)
        check(self._lib.vx_ar_decode_ex(self._h, int(max_new), int(top_k), float(top_p), float(temperature), int(seed) & (2 ** 64 - 1),
                                        0, None, ptr(toks), ptr(lens), ptr(slp), int(bool(resume)), int(bool(stop_on_finish)),
                                        ptr(fin), VX_HOST, None), "vx_ar_decode_ex")
        return toks, lens, slp, fin

    def ar_decode(self, max_new: int = 0, top_k: int = -100, top_p: float = 1.0, temperature: float = 1.0,
                  seed: int = 0, ignore_eos: bool = False, forced: Optional[np.ndarray] = None):
        """Returns (tokens int32 [B,max_steps] (sampled), lens int32 [B], sum_logprobs float32 [B])."""
        steps = self.ar_budget(max_new)
        toks = np.zeros((self.B, steps), np.int32)
        lens = np.zeros((self.B,), np.int32)
        slp = np.zeros((self.B,), np.float32)
        f = None
        if forced is not None:
            f = np.zeros((self.B, steps), np.int32)
            n = min(steps, forced.shape[1])
            f[:, :n] = forced[:, :n]
        check(self._lib.vx_ar_decode(self._h, int(max_new), int(top_k), float(top_p), float(temperature),
                                     int(seed) & (2 ** 64 - 1), int(bool(ignore_eos)), ptr(f), ptr(toks), ptr(lens),
                                     ptr(slp), VX_HOST, None), "vx_ar_decode")
        return toks, lens, slp

    def nar(self, cb0: Optional[np.ndarray] = None, t_lens: Optional[np.ndarray] = None, Tmax: Optional[int] = None):
        """cb0 int [B,Tmax] (None: use the AR output on the device).  Returns codes int32 [B,Tmax,8]."""
        if cb0 is not None:
            cb0 = np.ascontiguousarray(cb0, np.int32)
            t_lens = np.ascontiguousarray(t_lens, np.int32)
            Tmax = cb0.shape[1]
        assert Tmax is not None
        out = np.zeros((self.B, max(Tmax, 1), NUM_QUANTIZERS), np.int32)
        check(self._lib.vx_nar(self._h, ptr(cb0), ptr(t_lens), int(Tmax), ptr(out), VX_HOST, None), "vx_nar")
        return out[:, :Tmax]

    def vocos_decode(self, codes: np.ndarray, bandwidth_id: int = 2) -> np.ndarray:
        """codes int [8,B,T] -> wave float32 [B, 320*T]."""
        codes = np.ascontiguousarray(codes, np.int32)
        K, B, T = codes.shape
        assert K == NUM_QUANTIZERS
        wave = np.zeros((B, 320 * T), np.float32)
        check(self._lib.vx_vocos_decode(self._h, ptr(codes), B, T, int(bandwidth_id), ptr(wave), VX_HOST, None),
              "vx_vocos_decode")
        return wave

    # raw-pointer variants for device-resident buffers (bench `value` leg)
    def raw(self):
        return self._lib, self._h

    def timing(self) -> dict:
        a, b, c, d = C.c_float(), C.c_float(), C.c_float(), C.c_float()
        s, k = C.c_int64(), C.c_int64()
        check(self._lib.vx_last_timing(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(s), C.byref(k)),
              "vx_last_timing")
        return dict(ms_prefill=a.value, ms_decode=b.value, ms_nar=c.value, ms_vocos=d.value, decode_steps=s.value,
                    kernel_launches=k.value)

    def ctx_clamped(self) -> bool:
        """True if max_ctx (not EOS / 16*S / max_new) limited some utterance's budget in the last ar_decode."""
        return bool(self._lib.vx_ar_ctx_clamped(self._h))

    def last_decode_kernel(self) -> str:
        """Engine-reported name of the kernel the last ar_decode launched."""
        v = self._lib.vx_last_decode_kernel(self._h)
        return v.decode() if v else ""

    def debug_read(self, what: str, max_elems: int) -> np.ndarray:
        buf = np.zeros((max_elems,), np.float32)
        n = C.c_int64()
        check(self._lib.vx_debug_read(self._h, what.encode(), ptr(buf), max_elems, C.byref(n)), "vx_debug_read")
        return buf[:n.value]
