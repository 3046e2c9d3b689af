"""CPU restatement of the EnCodec 24 kHz ENCODER + residual vector quantiser -- TEST INFRASTRUCTURE, not product.

Path: SURVEY.md 8(f4), the audio half of prompt making.  Reference call sites: `data/tokenizer.py:63-111`
(`AudioTokenizer.encode` = `EncodecModel.encodec_model_24khz()` with `set_target_bandwidth(6.0)` -> 8 codebooks, weight
norm removed, `tokenize_audio`), `utils/prompt_making.py:57-84` (`encoded_frames[0][0].transpose(2, 1)` -> the
`audio_tokens [1,T,8]` array of a prompt .npz).

The arithmetic lives in the third-party `encodec` package (requirements.txt:7, version unpinned, not vendored, not
installable here).  This file restates its published algorithm (SEANet encoder: `encodec/modules/seanet.py`
`SEANetEncoder`, `encodec/modules/conv.py` `SConv1d` causal reflect padding, `encodec/modules/lstm.py` `SLSTM` with skip,
`encodec/quantization/core_vq.py` `EuclideanCodebook.quantize` / `ResidualVectorQuantization.encode`) for the
`encodec_24khz` configuration: channels 1, n_filters 32, ratios [8,5,4,2] (encoder runs them reversed), one residual
layer per stage (kernel 3 then 1, compress 2, true-skip off = convolutional shortcut), ELU(alpha 1), causal, reflect
padding, 2 LSTM layers, dimension 128, 1024 bins, 75 frames/s, `normalize=False`, no chunking (`segment=None`), so
`encode(wav)` returns ONE frame `[(codes[B, n_q, T], None)]`.

Pinning: **pinned to transformers 5.5 `EncodecModel`** (an independent port of the same model that ships in this image):
tests/golden/make_encodec_golden.py loads the synthetic state dict of `vallex_b200.weights.synthetic_encodec_state_dict`
into `transformers.EncodecModel(EncodecConfig())`, runs its `encode(..., bandwidth=6.0)` and commits inputs, codes and
embedding checksums under tests/golden/encodec_*.npz; tests/test_oracle_golden.py requires this restatement to reproduce
those codes exactly and checks it live against transformers when importable.  The reference's own `encodec` package is
absent, so parity with IT is by construction of the two ports from the same published source.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

RATIOS = (2, 4, 5, 8)
FILTERS = 32
DIM = 128
FRAME_RATE = 75
BINS = 1024


def n_q_for_bandwidth(bandwidth_kbps: float) -> int:
    """`ResidualVectorQuantizer.get_num_quantizers_for_bandwidth`: floor(bw*1000 / (log2(bins) * frame_rate)), >= 1."""
    return int(max(1, math.floor(bandwidth_kbps * 1000 / (math.log2(BINS) * FRAME_RATE))))


def _weight(sd, prefix):
    """Plain `weight` (the reference strips weight norm, data/tokenizer.py:31-60) or `weight_g`/`weight_v`
    (w = g * v / ||v||, norm over all dims but 0) or torch's parametrization names."""
    if f"{prefix}.weight" in sd:
        return sd[f"{prefix}.weight"].float()
    for gk, vk in ((f"{prefix}.weight_g", f"{prefix}.weight_v"),
                   (f"{prefix}.parametrizations.weight.original0", f"{prefix}.parametrizations.weight.original1")):
        if gk in sd:
            g, v = sd[gk].float(), sd[vk].float()
            return g * v / v.flatten(1).norm(dim=1).view(-1, 1, 1)
    raise KeyError(prefix)


def pad1d_reflect(x: torch.Tensor, left: int, right: int) -> torch.Tensor:
    """`encodec.modules.conv.pad1d(mode='reflect')`: inputs not longer than the padding get zeros appended first."""
    length = x.shape[-1]
    max_pad = max(left, right)
    extra = 0
    if length <= max_pad:
        extra = max_pad - length + 1
        x = F.pad(x, (0, extra))
    y = F.pad(x, (left, right), mode="reflect")
    return y[..., : y.shape[-1] - extra]


def sconv1d(x, w, b, stride=1, dilation=1):
    """Causal `SConv1d`: left pad (k-1)*d+1-stride, right pad up to a whole last frame, reflect."""
    k_eff = (w.shape[-1] - 1) * dilation + 1
    pad_total = k_eff - stride
    length = x.shape[-1]
    n_frames = math.ceil((length - k_eff + pad_total) / stride + 1) - 1
    extra = n_frames * stride + k_eff - pad_total - length
    x = pad1d_reflect(x, pad_total, extra)
    return F.conv1d(x, w, b, stride=stride, dilation=dilation)


def lstm_layer(x, w_ih, w_hh, b_ih, b_hh):
    """One `nn.LSTM` layer, time-major x [T,B,H]; gate order i, f, g, o."""
    T, B, H = x.shape
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    pre = x @ w_ih.t() + b_ih
    out = []
    for t in range(T):
        gates = pre[t] + h @ w_hh.t() + b_hh
        i, f, g, o = gates.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        out.append(h)
    return torch.stack(out)


class OracleEncodec:
    def __init__(self, state_dict, bandwidth: float = 6.0):
        self.sd = {k: v for k, v in state_dict.items()}
        self.n_q = n_q_for_bandwidth(bandwidth)

    def _conv(self, x, prefix, stride=1, dilation=1):
        return sconv1d(x, _weight(self.sd, prefix), self.sd[f"{prefix}.bias"].float(), stride, dilation)

    @torch.no_grad()
    def encoder(self, wav: torch.Tensor) -> torch.Tensor:
        """wav [B,1,N] fp32 at 24 kHz -> embeddings [B,128,T], T = ceil(N/320)."""
        x = self._conv(wav.float(), "encoder.model.0.conv.conv")
        idx = 1
        for r in RATIOS:
            y = self._conv(F.elu(x), f"encoder.model.{idx}.block.1.conv.conv")
            y = self._conv(F.elu(y), f"encoder.model.{idx}.block.3.conv.conv")
            x = self._conv(x, f"encoder.model.{idx}.shortcut.conv.conv") + y
            x = self._conv(F.elu(x), f"encoder.model.{idx + 2}.conv.conv", stride=r)
            idx += 3
        s = x.permute(2, 0, 1)
        y = s
        for layer in range(2):
            p = f"encoder.model.{idx}.lstm"
            y = lstm_layer(y, self.sd[f"{p}.weight_ih_l{layer}"].float(), self.sd[f"{p}.weight_hh_l{layer}"].float(),
                           self.sd[f"{p}.bias_ih_l{layer}"].float(), self.sd[f"{p}.bias_hh_l{layer}"].float())
        x = (y + s).permute(1, 2, 0)
        return self._conv(F.elu(x), f"encoder.model.{idx + 2}.conv.conv")

    @torch.no_grad()
    def quantize(self, emb: torch.Tensor, return_margins: bool = False):
        """emb [B,128,T] -> codes [B,n_q,T] int64 (and, on request, the top-1/top-2 gap of every decision)."""
        B, D, T = emb.shape
        res = emb.permute(0, 2, 1).reshape(-1, D)
        codes, margins = [], []
        for q in range(self.n_q):
            embed = self.sd[f"quantizer.vq.layers.{q}._codebook.embed"].float()
            et = embed.t()
            dist = -(res.pow(2).sum(1, keepdim=True) - 2 * res @ et + et.pow(2).sum(0, keepdim=True))
            ind = dist.max(dim=-1).indices
            if return_margins:
                top2 = dist.topk(2, dim=-1).values
                margins.append((top2[:, 0] - top2[:, 1]).view(B, T))
            res = res - embed[ind]
            codes.append(ind.view(B, T))
        codes = torch.stack(codes, dim=1)
        return (codes, torch.stack(margins, dim=1)) if return_margins else codes

    @torch.no_grad()
    def encode(self, wav: torch.Tensor):
        """`EncodecModel.encode(wav)` for the 24 kHz model: a one-element list of (codes [B,n_q,T], scale None)."""
        return [(self.quantize(self.encoder(wav)), None)]
