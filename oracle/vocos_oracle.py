"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  ** PARITY UNPINNED **

Plain-PyTorch (CPU, fp32) restatement of the Vocos decode used by the reference at
`utils/generation.py:89,148-150` (`Vocos.from_pretrained('charactr/vocos-encodec-24khz')`,
`codes_to_features`, `decode(features, bandwidth_id=tensor([2]))`).

The arithmetic lives in the third-party pip package `vocos` (gemelo-ai/vocos, ex charactr), which is
UNPINNED in the reference (`requirements.txt:23`), is not vendored under /root/reference, and is not
installed here; its pretrained weights are not reachable either (no network).  No reference test
pins its output.  This file therefore restates the package's PUBLISHED algorithm from its public
source layout -- `pretrained.py` (codes_to_features, decode), `feature_extractors.py`
(EncodecFeatures.codebook_weights), `models.py` (VocosBackbone), `modules.py` (ConvNeXtBlock,
AdaLayerNorm), `heads.py` (ISTFTHead), `spectral_ops.py` (ISTFT, padding="same") -- with the
config of the `charactr/vocos-encodec-24khz` checkpoint (input 128, dim 384, intermediate 1152,
8 layers, 4 bandwidth embeddings, n_fft 1280, hop 320).  Waveform parity of the B200 engine is
checked against THIS restatement only; re-verify against the real package where `import vocos`
works.  One stage has an independent pin: the iSTFT (`istft_same`) is checked against `torch.istft` in
tests/test_oracle_golden.py.  Only tests/, smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

N_FFT = 1280
HOP = 320
DIM = 384
LAYERS = 8
BINS = 1024


class OracleVocos:
    def __init__(self, state_dict: Dict[str, torch.Tensor]):
        self.sd = {k: v.detach().to(torch.float32) for k, v in state_dict.items()}

    @torch.no_grad()
    def codes_to_features(self, codes: torch.Tensor) -> torch.Tensor:
        """vocos pretrained.py `codes_to_features`: codes int [K,B,T] -> features [B,128,T].
        Row offset k*1024 into the concatenated codebook table, sum over the K codebooks."""
        if codes.dim() == 2:
            codes = codes.unsqueeze(1)
        offsets = torch.arange(0, BINS * len(codes), BINS)
        idx = codes.long() + offsets.view(-1, 1, 1)
        feats = F.embedding(idx, self.sd["feature_extractor.codebook_weights"]).sum(dim=0)
        return feats.transpose(1, 2)

    def _adanorm(self, prefix: str, x: torch.Tensor, bw: torch.Tensor) -> torch.Tensor:
        """vocos modules.py `AdaLayerNorm.forward`: LN without affine (eps 1e-6) * scale[bw] + shift[bw]."""
        scale = F.embedding(bw, self.sd[f"{prefix}.scale.weight"])
        shift = F.embedding(bw, self.sd[f"{prefix}.shift.weight"])
        x = F.layer_norm(x, (DIM,), eps=1e-6)
        return x * scale + shift

    @torch.no_grad()
    def backbone(self, feats: torch.Tensor, bandwidth_id: torch.Tensor) -> torch.Tensor:
        """vocos models.py `VocosBackbone.forward` + modules.py `ConvNeXtBlock.forward`.  -> [B,T,384]."""
        sd = self.sd
        x = F.conv1d(feats, sd["backbone.embed.weight"], sd["backbone.embed.bias"], padding=3)
        x = self._adanorm("backbone.norm", x.transpose(1, 2), bandwidth_id).transpose(1, 2)
        for i in range(LAYERS):
            p = f"backbone.convnext.{i}"
            res = x
            h = F.conv1d(x, sd[f"{p}.dwconv.weight"], sd[f"{p}.dwconv.bias"], padding=3, groups=DIM)
            h = self._adanorm(f"{p}.norm", h.transpose(1, 2), bandwidth_id)
            h = F.linear(h, sd[f"{p}.pwconv1.weight"], sd[f"{p}.pwconv1.bias"])
            h = F.gelu(h)
            h = F.linear(h, sd[f"{p}.pwconv2.weight"], sd[f"{p}.pwconv2.bias"])
            h = sd[f"{p}.gamma"] * h
            x = res + h.transpose(1, 2)
        return F.layer_norm(x.transpose(1, 2), (DIM,), sd["backbone.final_layer_norm.weight"],
                            sd["backbone.final_layer_norm.bias"], 1e-6)

    @torch.no_grad()
    def head(self, x: torch.Tensor) -> torch.Tensor:
        """vocos heads.py `ISTFTHead.forward` + spectral_ops.py `ISTFT.forward` (padding='same')."""
        sd = self.sd
        x = F.linear(x, sd["head.out.weight"], sd["head.out.bias"]).transpose(1, 2)
        mag, p = x.chunk(2, dim=1)
        mag = torch.clip(torch.exp(mag), max=1e2)
        S = mag * (torch.cos(p) + 1j * torch.sin(p))
        return self.istft_same(S)

    @torch.no_grad()
    def istft_same(self, S: torch.Tensor) -> torch.Tensor:
        """vocos spectral_ops.py `ISTFT.forward`, padding='same': complex [B,641,T] -> [B, 320*T].  (Pinned on its own by
        tests/test_oracle_golden.py against torch.istft, which differs only in how much it trims from both ends.)"""
        window = self.sd["head.istft.window"]
        pad = (N_FFT - HOP) // 2
        B, N, T = S.shape
        ifft = torch.fft.irfft(S, N_FFT, dim=1, norm="backward") * window[None, :, None]
        out_size = (T - 1) * HOP + N_FFT
        y = F.fold(ifft, output_size=(1, out_size), kernel_size=(1, N_FFT), stride=(1, HOP))[:, 0, 0, pad:-pad]
        wsq = window.square().expand(1, T, -1).transpose(1, 2)
        env = F.fold(wsq, output_size=(1, out_size), kernel_size=(1, N_FFT), stride=(1, HOP)).squeeze()[pad:-pad]
        assert (env > 1e-11).all()
        return y / env

    @torch.no_grad()
    def decode(self, features: torch.Tensor, bandwidth_id: torch.Tensor) -> torch.Tensor:
        """vocos pretrained.py `decode`: [B,128,T] -> [B, 320*T]."""
        return self.head(self.backbone(features, bandwidth_id))
