"""Drop-in for the slice of the third-party `vocos.Vocos` API the reference uses
(utils/generation.py:4,89,148-150): `Vocos.from_pretrained(path_or_id)`, `.to(device)`,
`.codes_to_features(codes[8,B,T])`, `.decode(features, bandwidth_id=tensor([2])) -> [B, 320*T]`.

On the B200 engine the codebook gather, ConvNeXt backbone and iSTFT head run as one call
(`vx_vocos_decode`), so `codes_to_features` returns a light `CodesFeatures` handle (it reports the
feature shape [B,128,T]) that `decode` consumes.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np
import torch

from .engine import Engine
from .weights import synthetic_vocos_state_dict


class CodesFeatures:
    def __init__(self, codes: np.ndarray):
        self.codes = codes  # int32 [8,B,T]

    @property
    def shape(self):
        return torch.Size([self.codes.shape[1], 128, self.codes.shape[2]])


class Vocos:
    def __init__(self, state_dict: Dict[str, torch.Tensor], dtype: str = "bf16"):
        self._sd = state_dict
        self._device = 0
        self._engine: Optional[Engine] = None
        self._owner = None
        self._dtype = dtype

    @classmethod
    def from_pretrained(cls, repo_or_path: str = "charactr/vocos-encodec-24khz", **kw) -> "Vocos":
        """Loads `pytorch_model.bin`-style weights from a local path.  There is no network here: a hub id
        resolves only through VOCOS_B200_WEIGHTS / ./checkpoints/vocos-encodec-24khz.pt; with
        VALLEX_B200_SYNTHETIC=1 seeded random weights of the same shapes are used (benchmarks)."""
        cands = [repo_or_path, os.environ.get("VOCOS_B200_WEIGHTS", ""), "./checkpoints/vocos-encodec-24khz.pt"]
        for c in cands:
            if c and os.path.isfile(c):
                sd = torch.load(c, map_location="cpu")
                return cls(sd.get("state_dict", sd), **kw)
        if os.environ.get("VALLEX_B200_SYNTHETIC", "0") == "1":
            return cls(synthetic_vocos_state_dict(0), **kw)
        raise FileNotFoundError(f"Vocos weights not found for {repo_or_path!r} (offline: set VOCOS_B200_WEIGHTS)")

    def share_engine(self, valle) -> "Vocos":
        """Run on the VALLE model's engine (one engine per GPU) instead of a private one."""
        self._owner = valle
        valle.attach_vocos(self._sd)
        return self

    def to(self, device):
        d = torch.device(device) if not isinstance(device, torch.device) else device
        if d.type != "cuda":
            raise RuntimeError("vallex_b200 runs on CUDA (B200, sm_100a) only; there is no CPU path")
        self._device = 0 if d.index is None else d.index
        return self

    def eval(self):
        return self

    @property
    def engine(self) -> Engine:
        if self._owner is not None:
            return self._owner.engine
        if self._engine is None:
            eng = Engine(self._device, 1, 64, self._dtype)
            eng.load_vocos_state_dict(self._sd)
            eng.finalize()
            self._engine = eng
        return self._engine

    def codes_to_features(self, codes: torch.Tensor) -> CodesFeatures:
        if codes.dim() == 2:
            codes = codes.unsqueeze(1)
        if codes.numel() and (int(codes.min()) < 0 or int(codes.max()) >= 1024):   # F.embedding raises the same in the vocos package
            raise IndexError("index out of range in self (EnCodec code outside 0..1023)")
        return CodesFeatures(np.ascontiguousarray(codes.detach().to("cpu").numpy().astype(np.int32)))

    def decode(self, features_input, bandwidth_id: Optional[torch.Tensor] = None, **kwargs) -> torch.Tensor:
        if not isinstance(features_input, CodesFeatures):
            raise NotImplementedError("decode() takes the handle returned by codes_to_features()")
        bw = 0 if bandwidth_id is None else int(torch.as_tensor(bandwidth_id).reshape(-1)[0])
        wave = self.engine.vocos_decode(features_input.codes, bw)
        return torch.from_numpy(wave)
