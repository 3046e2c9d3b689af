"""Synthetic (seeded, random-init) checkpoints in the reference's key layout.

There is no network in the build/bench environment, so the real
`vallex-checkpoint.pt` / `charactr/vocos-encodec-24khz` weights cannot be
fetched.  These generators produce state dicts with exactly the key names,
shapes and dtypes the reference model exposes (reference
`models/vallex.py:55-264,400-445`, `modules/transformer.py:83-108,178-260`),
so the same dict can be loaded strict into the reference `VALLE` (golden
generation, `tests/golden/make_golden.py`), into the CPU oracle, and into the
B200 engine.  Scales follow torch's default initialisers so activations have
realistic magnitudes; LayerNorm affines are perturbed away from (1, 0) so a
kernel that forgets them fails parity.

The Vocos dict follows the public `vocos` package layout (third party, not in
the reference tree; see oracle/vocos_oracle.py header -- parity unpinned).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

from .macros import N_DIM, NUM_HEAD, NUM_LAYERS, NUM_QUANTIZERS, NUM_TEXT_TOKENS, NUM_AUDIO_TOKENS

D = N_DIM
FF = 4 * N_DIM


def _uniform(g, shape, bound):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound


def _normal(g, shape, std=1.0):
    return torch.randn(shape, generator=g, dtype=torch.float32) * std


def _encoder_layer(sd, g, prefix, adaptive):
    sd[f"{prefix}.self_attn.in_proj_weight"] = _uniform(g, (3 * D, D), math.sqrt(6.0 / (D + 3 * D)))
    sd[f"{prefix}.self_attn.in_proj_bias"] = _uniform(g, (3 * D,), 0.02)
    sd[f"{prefix}.self_attn.out_proj.weight"] = _uniform(g, (D, D), 1.0 / math.sqrt(D))
    sd[f"{prefix}.self_attn.out_proj.bias"] = _uniform(g, (D,), 0.02)
    sd[f"{prefix}.linear1.weight"] = _uniform(g, (FF, D), 1.0 / math.sqrt(D))
    sd[f"{prefix}.linear1.bias"] = _uniform(g, (FF,), 1.0 / math.sqrt(D))
    sd[f"{prefix}.linear2.weight"] = _uniform(g, (D, FF), 1.0 / math.sqrt(FF))
    sd[f"{prefix}.linear2.bias"] = _uniform(g, (D,), 1.0 / math.sqrt(FF))
    for n in ("norm1", "norm2"):
        if adaptive:
            sd[f"{prefix}.{n}.project_layer.weight"] = _uniform(g, (2 * D, D), 1.0 / math.sqrt(D))
            sd[f"{prefix}.{n}.project_layer.bias"] = _uniform(g, (2 * D,), 1.0 / math.sqrt(D))
            sd[f"{prefix}.{n}.norm.weight"] = 1.0 + _normal(g, (D,), 0.1)
            sd[f"{prefix}.{n}.norm.bias"] = _normal(g, (D,), 0.1)
        else:
            sd[f"{prefix}.{n}.weight"] = 1.0 + _normal(g, (D,), 0.1)
            sd[f"{prefix}.{n}.bias"] = _normal(g, (D,), 0.1)


def synthetic_state_dict(seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """374-key fp32 state dict in the order the reference's `state_dict()` yields."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000003 * (seed + 1))
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    sd["ar_text_embedding.word_embeddings.weight"] = _normal(g, (NUM_TEXT_TOKENS, D))
    sd["nar_text_embedding.word_embeddings.weight"] = _normal(g, (NUM_TEXT_TOKENS, D))
    sd["ar_audio_embedding.word_embeddings.weight"] = _normal(g, (NUM_AUDIO_TOKENS + 2, D))
    sd["ar_text_position.alpha"] = torch.tensor([1.0 + 0.25 * float(_normal(g, (1,))[0])])
    sd["ar_audio_position.alpha"] = torch.tensor([1.0 + 0.25 * float(_normal(g, (1,))[0])])
    for i in range(NUM_LAYERS):
        _encoder_layer(sd, g, f"ar_decoder.layers.{i}", adaptive=False)
    sd["ar_decoder.norm.weight"] = 1.0 + _normal(g, (D,), 0.1)
    sd["ar_decoder.norm.bias"] = _normal(g, (D,), 0.1)
    sd["ar_predict_layer.weight"] = _uniform(g, (NUM_AUDIO_TOKENS + 1, D), 1.0 / math.sqrt(D))
    sd["nar_audio_embeddings.0.word_embeddings.weight"] = _normal(g, (NUM_AUDIO_TOKENS + 1, D))
    for j in range(1, NUM_QUANTIZERS):
        sd[f"nar_audio_embeddings.{j}.word_embeddings.weight"] = _normal(g, (NUM_AUDIO_TOKENS, D))
    sd["nar_text_position.alpha"] = torch.ones(1)
    sd["nar_audio_position.alpha"] = torch.ones(1)
    for i in range(NUM_LAYERS):
        _encoder_layer(sd, g, f"nar_decoder.layers.{i}", adaptive=True)
    sd["nar_decoder.norm.project_layer.weight"] = _uniform(g, (2 * D, D), 1.0 / math.sqrt(D))
    sd["nar_decoder.norm.project_layer.bias"] = _uniform(g, (2 * D,), 1.0 / math.sqrt(D))
    sd["nar_decoder.norm.norm.weight"] = 1.0 + _normal(g, (D,), 0.1)
    sd["nar_decoder.norm.norm.bias"] = _normal(g, (D,), 0.1)
    # nar_predict_layers.{0..5} alias nar_audio_embeddings.{2..7} (models/vallex.py:261-264)
    for j in range(NUM_QUANTIZERS - 2):
        sd[f"nar_predict_layers.{j}.weight"] = sd[f"nar_audio_embeddings.{j + 2}.word_embeddings.weight"]
    sd[f"nar_predict_layers.{NUM_QUANTIZERS - 2}.weight"] = _uniform(g, (NUM_AUDIO_TOKENS, D), 1.0 / math.sqrt(D))
    for j in range(NUM_QUANTIZERS - 1):
        sd[f"nar_stage_embeddings.{j}.word_embeddings.weight"] = _normal(g, (1, D))
    sd["ar_language_embedding.word_embeddings.weight"] = _normal(g, (3, D))
    sd["nar_language_embedding.word_embeddings.weight"] = _normal(g, (3, D))
    return sd


# ---- Vocos (charactr/vocos-encodec-24khz shapes; third-party, from the public package layout) ----
VOCOS_IN = 128
VOCOS_DIM = 384
VOCOS_FF = 1152
VOCOS_LAYERS = 8
VOCOS_NUM_BW = 4
VOCOS_NFFT = 1280
VOCOS_HOP = 320
VOCOS_CODEBOOKS = 16  # 12 kbps table; the path uses the first 8 (6 kbps)


def synthetic_vocos_state_dict(seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    g = torch.Generator(device="cpu")
    g.manual_seed(7000003 * (seed + 1))
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    sd["feature_extractor.codebook_weights"] = _normal(g, (VOCOS_CODEBOOKS * 1024, VOCOS_IN), 0.5)
    sd["backbone.embed.weight"] = _uniform(g, (VOCOS_DIM, VOCOS_IN, 7), 1.0 / math.sqrt(VOCOS_IN * 7))
    sd["backbone.embed.bias"] = _uniform(g, (VOCOS_DIM,), 1.0 / math.sqrt(VOCOS_IN * 7))
    sd["backbone.norm.scale.weight"] = 1.0 + _normal(g, (VOCOS_NUM_BW, VOCOS_DIM), 0.1)
    sd["backbone.norm.shift.weight"] = _normal(g, (VOCOS_NUM_BW, VOCOS_DIM), 0.1)
    for i in range(VOCOS_LAYERS):
        p = f"backbone.convnext.{i}"
        sd[f"{p}.dwconv.weight"] = _uniform(g, (VOCOS_DIM, 1, 7), 1.0 / math.sqrt(7))
        sd[f"{p}.dwconv.bias"] = _uniform(g, (VOCOS_DIM,), 1.0 / math.sqrt(7))
        sd[f"{p}.norm.scale.weight"] = 1.0 + _normal(g, (VOCOS_NUM_BW, VOCOS_DIM), 0.1)
        sd[f"{p}.norm.shift.weight"] = _normal(g, (VOCOS_NUM_BW, VOCOS_DIM), 0.1)
        sd[f"{p}.pwconv1.weight"] = _uniform(g, (VOCOS_FF, VOCOS_DIM), 1.0 / math.sqrt(VOCOS_DIM))
        sd[f"{p}.pwconv1.bias"] = _uniform(g, (VOCOS_FF,), 1.0 / math.sqrt(VOCOS_DIM))
        sd[f"{p}.pwconv2.weight"] = _uniform(g, (VOCOS_DIM, VOCOS_FF), 1.0 / math.sqrt(VOCOS_FF))
        sd[f"{p}.pwconv2.bias"] = _uniform(g, (VOCOS_DIM,), 1.0 / math.sqrt(VOCOS_FF))
        sd[f"{p}.gamma"] = torch.full((VOCOS_DIM,), 1.0 / VOCOS_LAYERS) + _normal(g, (VOCOS_DIM,), 0.02)
    sd["backbone.final_layer_norm.weight"] = 1.0 + _normal(g, (VOCOS_DIM,), 0.1)
    sd["backbone.final_layer_norm.bias"] = _normal(g, (VOCOS_DIM,), 0.1)
    # small head weights keep exp(mag) well inside the clip range, like a trained head
    sd["head.out.weight"] = _uniform(g, (VOCOS_NFFT + 2, VOCOS_DIM), 0.5 / math.sqrt(VOCOS_DIM))
    sd["head.out.bias"] = _uniform(g, (VOCOS_NFFT + 2,), 0.5 / math.sqrt(VOCOS_DIM))
    sd["head.istft.window"] = torch.hann_window(VOCOS_NFFT)
    return sd


# ---- EnCodec 24 kHz encoder + RVQ (reference data/tokenizer.py:63-111) ----
# Key layout of `encodec.EncodecModel.encodec_model_24khz().state_dict()` after the reference's
# `remove_encodec_weight_norm` (data/tokenizer.py:31-60): `encoder.model.<i>` is the SEANet nn.Sequential
# (conv, 4 x [resnet block, ELU, strided conv], LSTM, ELU, conv), `quantizer.vq.layers.<q>._codebook.embed` the RVQ tables.
# The `encodec` package is third party and not installed here: names follow its public source layout (from memory);
# the architecture is cross-checked against transformers' port of the same model (tests/golden/make_encodec_golden.py).
ENCODEC_RATIOS = (2, 4, 5, 8)  # encoder order = reversed SEANet `ratios` [8,5,4,2]
ENCODEC_FILTERS = 32
ENCODEC_DIM = 128
ENCODEC_CODEBOOKS = 32  # 24 kbps table; 6 kbps (data/tokenizer.py:71) uses the first 8
ENCODEC_BINS = 1024


def encodec_conv_specs():
    """(state-dict prefix, C_out, C_in, kernel, stride, dilation) of every encoder convolution, in execution order;
    resnet blocks contribute (block.1 k3, block.3 k1, shortcut k1)."""
    specs = [("encoder.model.0.conv.conv", ENCODEC_FILTERS, 1, 7, 1, 1)]
    dim, idx = ENCODEC_FILTERS, 1
    for r in ENCODEC_RATIOS:
        specs.append((f"encoder.model.{idx}.block.1.conv.conv", dim // 2, dim, 3, 1, 1))
        specs.append((f"encoder.model.{idx}.block.3.conv.conv", dim, dim // 2, 1, 1, 1))
        specs.append((f"encoder.model.{idx}.shortcut.conv.conv", dim, dim, 1, 1, 1))
        specs.append((f"encoder.model.{idx + 2}.conv.conv", 2 * dim, dim, 2 * r, r, 1))
        dim, idx = 2 * dim, idx + 3
    specs.append((f"encoder.model.{idx + 2}.conv.conv", ENCODEC_DIM, dim, 7, 1, 1))  # idx = 13 is the LSTM
    return specs


def synthetic_encodec_state_dict(seed: int = 0, codebooks: int = 8) -> "OrderedDict[str, torch.Tensor]":
    g = torch.Generator(device="cpu")
    g.manual_seed(9000011 * (seed + 1))
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for prefix, co, ci, k, _s, _d in encodec_conv_specs():
        sd[f"{prefix}.weight"] = _normal(g, (co, ci, k), math.sqrt(2.0 / (ci * k)))
        sd[f"{prefix}.bias"] = _uniform(g, (co,), 1.0 / math.sqrt(ci * k))
    h = ENCODEC_FILTERS * 16
    for layer in range(2):
        for n, shape in (("weight_ih", (4 * h, h)), ("weight_hh", (4 * h, h)), ("bias_ih", (4 * h,)), ("bias_hh", (4 * h,))):
            sd[f"encoder.model.13.lstm.{n}_l{layer}"] = _uniform(g, shape, 1.0 / math.sqrt(h))
    for q in range(codebooks):
        # residuals shrink level by level in a trained RVQ; mimic that so deep levels are exercised at small magnitudes
        sd[f"quantizer.vq.layers.{q}._codebook.embed"] = _normal(g, (ENCODEC_BINS, ENCODEC_DIM), 0.6 * 0.8 ** q)
    return sd


_EXPECTED = None


def expected_state_dict_shapes() -> "OrderedDict[str, tuple]":
    """Key -> shape of the reference model's `state_dict()` (374 keys); used by the strict
    `VALLE.load_state_dict` check (utils/generation.py:79-83)."""
    global _EXPECTED
    if _EXPECTED is None:
        _EXPECTED = OrderedDict((k, tuple(v.shape)) for k, v in synthetic_state_dict(0).items())
    return _EXPECTED


_BF16_PAT = __import__("re").compile(
    r"(self_attn\.in_proj_weight|self_attn\.out_proj\.weight|linear1\.weight|linear2\.weight"
    r"|^ar_predict_layer\.weight|^nar_predict_layers\.\d+\.weight|^nar_audio_embeddings\.[2-7]\.word_embeddings\.weight)$")


def is_bf16_matrix(key: str) -> bool:
    """Which checkpoint tensors the engine's bf16 mode stores in bf16: every GEMM operand (QKV / out-proj /
    FFN / predict layers).  nar_audio_embeddings.{2..7} are included because the reference ties them to
    nar_predict_layers.{0..5} (models/vallex.py:261-264) -- one tensor, one precision."""
    return _BF16_PAT.search(key) is not None


def round_state_dict_bf16(sd):
    """The fp32 state dict holding exactly the values the bf16 engine mode computes with."""
    return type(sd)((k, (v.to(torch.bfloat16).to(torch.float32) if is_bf16_matrix(k) else v)) for k, v in sd.items())


def sine_position_table(length: int, dim: int = D) -> torch.Tensor:
    """fp32 sinusoid table, computed with the same torch CPU ops (exp/sin/cos in fp32) the
    reference uses in `modules/embedding.py:68-91` so the uploaded table is bit-identical."""
    position = torch.arange(0, length, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * -(math.log(10000.0) / dim))
    pe = torch.zeros(length, dim)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe
