"""Drop-in for the audio half of the reference's `data.tokenizer` (data/tokenizer.py:63-111):
`AudioTokenizer(device)`, `.encode(wav[B,1,N]) -> [(codes[B,8,T], None)]`, `tokenize_audio(tokenizer, audio)`.

The reference builds `encodec.EncodecModel.encodec_model_24khz()` at 6 kbps (8 codebooks) with weight norm removed and runs
its SEANet encoder + residual vector quantiser in PyTorch.  Here the encoder, the 2-layer LSTM and the RVQ run as sm_100a
kernels behind `vx_encodec_*` (csrc/encodec.cu); there is no CPU path.  `decode` (EnCodec's own decoder) is not on the
`generate_audio` path -- the reference decodes with Vocos (utils/generation.py:148-150) -- and raises.

Weights: the `encodec` package downloads `encodec_24khz-d7cc33bc.th` from the hub; offline this class takes a state dict,
a local file (ENCODEC_B200_WEIGHTS / ./checkpoints/encodec_24khz-d7cc33bc.th) or, with VALLEX_B200_SYNTHETIC=1, seeded
random weights of the same shapes (tests / benchmarks).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _capi
from ._capi import VX_HOST, check
from .weights import synthetic_encodec_state_dict

SAMPLE_RATE = 24000
CHANNELS = 1
TARGET_BANDWIDTH = 6.0           # data/tokenizer.py:71 -> 8 codebooks
N_Q = 8


def _encoder_keys(sd):
    return {k: v for k, v in sd.items() if k.startswith("encoder.") or k.startswith("quantizer.vq.layers.")}


class AudioTokenizer:
    """EnCodec audio (data/tokenizer.py:63-96)."""

    def __init__(self, device: Any = None, state_dict: Optional[Dict[str, torch.Tensor]] = None) -> None:
        if state_dict is None:
            state_dict = self._find_weights()
        if not device:
            device = torch.device("cuda:0")       # the reference's default when CUDA is present (data/tokenizer.py:75-80)
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("vallex_b200 runs on CUDA (B200, sm_100a) only; there is no CPU path")
        self._device = device
        self.sample_rate = SAMPLE_RATE
        self.channels = CHANNELS
        self._h = C.c_void_p()
        lib = _capi.lib()
        check(lib.vx_encodec_create(device.index or 0, C.byref(self._h)), "vx_encodec_create")
        sd = _encoder_keys(state_dict)
        for k, v in sd.items():
            if "_codebook." in k and not k.endswith("._codebook.embed"):
                continue                            # cluster_size / embed_avg / inited: training-time EMA state
            if k.startswith("quantizer.vq.layers.") and int(k.split(".")[3]) >= 32:
                continue
            a = np.ascontiguousarray(v.detach().to("cpu", torch.float32).numpy())
            shape = (C.c_int64 * a.ndim)(*a.shape)
            check(lib.vx_encodec_load_weight(self._h, k.encode(), a.ctypes.data_as(C.c_void_p), shape, a.ndim), f"load {k}")
        check(lib.vx_encodec_finalize(self._h), "vx_encodec_finalize")     # strict: raises on a missing / mis-shaped tensor
        if lib.vx_encodec_num_codebooks(self._h) < N_Q:
            raise _capi.VallexB200Error("EnCodec state dict holds fewer than the 8 codebooks of 6 kbps")

    @staticmethod
    def _find_weights():
        for c in (os.environ.get("ENCODEC_B200_WEIGHTS", ""), "./checkpoints/encodec_24khz-d7cc33bc.th"):
            if c and os.path.isfile(c):
                sd = torch.load(c, map_location="cpu")
                return sd.get("state_dict", sd)
        if os.environ.get("VALLEX_B200_SYNTHETIC", "0") == "1":
            return synthetic_encodec_state_dict(0)
        raise FileNotFoundError("EnCodec weights not found (offline: set ENCODEC_B200_WEIGHTS to encodec_24khz-d7cc33bc.th)")

    @property
    def device(self):
        return self._device

    def close(self):
        if self._h:
            _capi.lib().vx_encodec_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def encode(self, wav: torch.Tensor) -> List[Tuple[torch.Tensor, None]]:
        """`EncodecModel.encode`: wav [B, 1, N] float at 24 kHz -> one frame `[(codes int64 [B, 8, T], None)]`."""
        if wav.dim() != 3 or wav.shape[1] != CHANNELS:
            raise ValueError(f"expected wav of shape [B, {CHANNELS}, N], got {tuple(wav.shape)}")   # encodec asserts the same
        B, _, N = wav.shape
        lib = _capi.lib()
        T = int(lib.vx_encodec_frames(N))
        out = np.empty((B, N_Q, T), dtype=np.int32)
        w = np.ascontiguousarray(wav.detach().to("cpu", torch.float32).numpy())
        for b in range(B):                      # utterances are independent (the reference passes B = 1)
            check(lib.vx_encodec_encode(self._h, w[b, 0].ctypes.data_as(C.c_void_p), N, N_Q, out[b].ctypes.data_as(C.c_void_p),
                                        VX_HOST, None), "vx_encodec_encode")
        return [(torch.from_numpy(out.astype(np.int64)).to(self._device), None)]

    def embeddings(self) -> torch.Tensor:
        """Encoder output [128, T] of the last encoded waveform (parity tests)."""
        lib = _capi.lib()
        ms, frames, launches = C.c_float(), C.c_int64(), C.c_int64()
        check(lib.vx_encodec_last_timing(self._h, C.byref(ms), C.byref(frames), C.byref(launches)), "vx_encodec_last_timing")
        buf = np.empty((128, int(frames.value)), dtype=np.float32)
        n = C.c_int64()
        check(lib.vx_encodec_embeddings(self._h, buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(n)), "vx_encodec_embeddings")
        return torch.from_numpy(buf)

    def last_timing(self):
        ms, frames, launches = C.c_float(), C.c_int64(), C.c_int64()
        check(_capi.lib().vx_encodec_last_timing(self._h, C.byref(ms), C.byref(frames), C.byref(launches)), "vx_encodec_last_timing")
        return {"ms": float(ms.value), "frames": int(frames.value), "kernel_launches": int(launches.value)}

    def decode(self, frames):
        raise NotImplementedError("EnCodec's decoder is not on the generate_audio path (the reference decodes with Vocos)")


def convert_audio(wav: torch.Tensor, sr: int, target_sr: int, target_channels: int) -> torch.Tensor:
    """`encodec.utils.convert_audio`: channel mix-down / expansion, then resampling (torchaudio, host side)."""
    assert wav.dim() >= 2, "Audio tensor must have at least 2 dimensions"
    assert wav.shape[-2] in [1, 2], "Audio must be mono or stereo."
    *shape, channels, length = wav.shape
    if target_channels == 1:
        wav = wav.mean(-2, keepdim=True)
    elif target_channels == 2:
        wav = wav.expand(*shape, target_channels, length)
    elif channels == 1:
        wav = wav.expand(target_channels, -1)
    else:
        raise RuntimeError(f"Impossible to convert from {channels} to {target_channels}")
    if sr != target_sr:
        import torchaudio

        wav = torchaudio.transforms.Resample(sr, target_sr)(wav)
    return wav


def tokenize_audio(tokenizer: AudioTokenizer, audio):
    """data/tokenizer.py:98-111: `audio` is a path or a (wav [C,N], sample_rate) pair."""
    if isinstance(audio, str):
        import torchaudio

        wav, sr = torchaudio.load(audio)
    else:
        wav, sr = audio
    wav = convert_audio(wav, sr, tokenizer.sample_rate, tokenizer.channels)
    wav = wav.unsqueeze(0)
    with torch.no_grad():
        encoded_frames = tokenizer.encode(wav)
    return encoded_frames
