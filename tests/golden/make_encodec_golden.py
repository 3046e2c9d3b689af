"""Generate tests/golden/encodec_*.npz from transformers' EncodecModel (an independent port of the `encodec` package the
reference calls in data/tokenizer.py:63-111; the package itself is not installable here).

    python tests/golden/make_encodec_golden.py

The synthetic state dict (`vallex_b200.weights.synthetic_encodec_state_dict`, encodec-package key layout) is renamed into
the transformers layout and loaded STRICT for the encoder + the first 8 codebooks, `EncodecModel.encode(wav, bandwidth=6.0)`
is run on seeded waveforms, and inputs + codes + per-decision margins + embedding statistics are committed.  Weights are
not stored: the generator is seeded.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "vall-e-x_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def hf_model(sd):
    """transformers.EncodecModel (24 kHz defaults) carrying `sd` (encodec-package keys)."""
    from transformers import EncodecConfig, EncodecModel

    torch.manual_seed(0)
    model = EncodecModel(EncodecConfig()).eval()
    own = model.state_dict()
    new = {}
    for k, v in sd.items():
        if k.startswith("encoder.model."):
            k2 = k.replace("encoder.model.", "encoder.layers.").replace(".conv.conv.", ".conv.")
            if k2.endswith(".conv.weight"):
                # transformers keeps the weight-norm parametrisation: g = ||v||, v = w reproduces w
                base = k2[: -len("weight")]
                new[base + "parametrizations.weight.original0"] = v.flatten(1).norm(dim=1).view(-1, 1, 1)
                new[base + "parametrizations.weight.original1"] = v
            else:
                new[k2] = v
        elif k.startswith("quantizer.vq.layers."):
            new[k.replace("quantizer.vq.layers.", "quantizer.layers.").replace("._codebook.", ".codebook.")] = v
    missing = [k for k in own if k.startswith("encoder.") and k not in new]
    assert not missing, missing
    unexpected = [k for k in new if k not in own]
    assert not unexpected, unexpected
    for k, v in new.items():
        assert own[k].shape == v.shape, (k, own[k].shape, v.shape)
    model.load_state_dict(new, strict=False)
    return model


def synth_wave(rng, n):
    """Speech-like test signal in [-1, 1]: a few drifting harmonics under a slow envelope plus noise."""
    t = np.arange(n) / 24000.0
    f0 = 110.0 + 40.0 * np.sin(2 * np.pi * 0.7 * t + rng.uniform(0, 6.28))
    ph = 2 * np.pi * np.cumsum(f0) / 24000.0
    x = sum(a * np.sin(h * ph + rng.uniform(0, 6.28)) for h, a in ((1, 0.5), (2, 0.25), (3, 0.15), (5, 0.08)))
    env = 0.5 + 0.5 * np.sin(2 * np.pi * 2.3 * t + rng.uniform(0, 6.28))
    x = x * env + 0.02 * rng.standard_normal(n)
    return (x / max(1.0, np.abs(x).max())).astype(np.float32)


CASES = {  # name -> (samples, seed)
    "encodec_tiny": (1000, 1),      # shorter than one hop multiple; exercises the right "extra" padding
    "encodec_short": (7, 2),        # shorter than the first conv's padding: the zero-extend-then-reflect branch
    "encodec_3s": (72000, 3),
    "encodec_ragged": (50123, 4),   # not a multiple of 320
}


def main():
    from oracle.encodec_oracle import OracleEncodec
    from vallex_b200.weights import synthetic_encodec_state_dict

    sd = synthetic_encodec_state_dict(0)
    model = hf_model(sd)
    ora = OracleEncodec(sd)
    for name, (n, seed) in CASES.items():
        wav = torch.from_numpy(synth_wave(np.random.default_rng(seed), n)).view(1, 1, n)
        with torch.no_grad():
            emb = model.encoder(wav)
            out = model.encode(wav, bandwidth=6.0)
        codes = out.audio_codes[0]  # [B, n_q, T]
        assert codes.shape[1] == 8
        emb_o = ora.encoder(wav)
        codes_o, margins = ora.quantize(emb, return_margins=True)
        assert torch.equal(codes_o, codes), "oracle RVQ differs from transformers on transformers' own embeddings"
        codes_oo = ora.encode(wav)[0][0]
        print(f"{name}: N={n} T={codes.shape[-1]} max|emb - oracle emb| = {float((emb - emb_o).abs().max()):.3e}"
              f" codes equal (oracle end to end) = {float((codes_oo == codes).float().mean()):.5f}"
              f" min margin {float(margins.min()):.3e}")
        np.savez_compressed(os.path.join(HERE, name + ".npz"), wav=wav.numpy()[0, 0], codes=codes.numpy().astype(np.int16),
                            margins=margins.numpy().astype(np.float32), emb_mean=emb.mean(dim=2).numpy(),
                            emb_abs_mean=emb.abs().mean().numpy(), emb_head=emb[0, :, :4].numpy(), seed=0)


if __name__ == "__main__":
    main()
