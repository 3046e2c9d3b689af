"""CPU: the oracle (oracle/vallex_oracle.py) is pinned to the fixtures the UNMODIFIED reference produced
(tests/golden/make_golden.py), bit-exactly, and -- where /root/reference exists -- live against it."""
import numpy as np
import pytest
import torch

from conftest import load_golden, state_dict_for
from oracle import ref_shim
from oracle.vallex_oracle import OracleVALLE


def _run(oracle, g, **kw):
    x = torch.from_numpy(g["x"].astype(np.int64))[None]
    y = torch.from_numpy(g["prompt"].astype(np.int64))[None]
    return oracle.inference(x, torch.tensor([x.shape[1]]), y, int(g["Lp"]), top_k=1, temperature=1.0,
                            prompt_language=g["prompt_language_arg"], text_language=g["text_language_arg"], **kw)


@pytest.mark.parametrize("name", ["tiny_en", "noprompt_en", "mixlang"])
def test_oracle_reproduces_reference_tokens(name):
    g = load_golden(name)
    dbg = {}
    codes = _run(OracleVALLE(state_dict_for(g["seed"])), g, debug=dbg)[0].numpy()
    assert codes.shape == g["codes"].shape
    assert np.array_equal(codes, g["codes"].astype(np.int64))
    head = torch.stack([l[0] for l in dbg["logits"][:4]]).numpy()
    assert np.array_equal(head, g["ar_logits_head"])  # same floats, not just same argmax
    top2 = torch.stack([torch.topk(l[0], 2)[0] for l in dbg["logits"]]).numpy()
    np.testing.assert_allclose(top2, g["ar_top2"], rtol=0, atol=0)


def test_oracle_reproduces_reference_tokens_bf16_rounded_weights():
    g = load_golden("tiny_en")
    codes = _run(OracleVALLE(state_dict_for(g["seed"], rounded=True)), g)[0].numpy()
    assert np.array_equal(codes, g["codes_bf16w"].astype(np.int64))


def test_oracle_prefix_of_full_size_case():
    """cfg1 (librispeech_1 prompt, S=122, T=1952): the AR prefix under max_new_tokens equals the reference's
    tokens (greedy decoding is causal), without paying for the whole CPU run."""
    g = load_golden("cfg1_librispeech")
    o = OracleVALLE(state_dict_for(g["seed"]))
    x = torch.from_numpy(g["x"].astype(np.int64))[None]
    y = torch.from_numpy(g["prompt"].astype(np.int64))[None]
    yy = o.ar_decode(x, y[..., 0], int(g["Lp"]), g["prompt_language_arg"], g["text_language_arg"], top_k=1,
                     max_new_tokens=24)
    assert g["codes"].shape == (16 * len(g["x"]), 8)  # the 16*S cap (models/vallex.py:577)
    assert np.array_equal(yy[0, 1 + y.shape[1]:].numpy(), g["codes"][:24, 0].astype(np.int64))


def test_oracle_prefix_of_long_context_case():
    """longctx_en (3900-frame prompt, context 4071 > 2048, audio PE positions > 4000 = past the reference's initial table,
    modules/embedding.py:68-72): the oracle's first tokens equal the unmodified reference's."""
    g = load_golden("longctx_en")
    o = OracleVALLE(state_dict_for(g["seed"]))
    x = torch.from_numpy(g["x"].astype(np.int64))[None]
    y = torch.from_numpy(g["prompt"].astype(np.int64))[None]
    dbg = {}
    yy = o.ar_decode(x, y[..., 0], int(g["Lp"]), "en", "en", top_k=1, max_new_tokens=2, debug=dbg)
    assert g["codes"].shape == (160, 8) and g["prompt"].shape == (3900, 8)
    assert np.array_equal(yy[0, 1 + y.shape[1]:].numpy(), g["codes"][:2, 0].astype(np.int64))
    top2 = torch.stack([torch.topk(l[0], 2)[0] for l in dbg["logits"]]).numpy()
    np.testing.assert_allclose(top2, g["ar_top2"][:3], rtol=0, atol=0)


def test_oracle_prefix_of_bench_shape_case():
    """cfg2_librispeech = the bench.py workload (S = 314, P = 225; the unmodified reference stops by itself at T = 1275 thanks to
    the fixture's crafted EOS row): the oracle's AR prefix and logit margins equal the reference's, for fp32 weights and for
    bf16-rounded matrix weights."""
    g = load_golden("cfg2_librispeech")
    assert g["codes"].shape == g["codes_bf16w"].shape == (1275, 8) and len(g["x"]) == 314 and g["prompt"].shape == (225, 8)
    x = torch.from_numpy(g["x"].astype(np.int64))[None]
    y = torch.from_numpy(g["prompt"].astype(np.int64))[None]
    for tag in ("", "_bf16w"):
        sd = dict(state_dict_for(0, rounded=bool(tag)))
        w = sd["ar_predict_layer.weight"].clone()
        w[1024] = torch.from_numpy(g["eos_row" + tag])
        sd["ar_predict_layer.weight"] = w
        dbg = {}
        yy = OracleVALLE(sd).ar_decode(x, y[..., 0], int(g["Lp"]), "en", "en", top_k=1, max_new_tokens=6, debug=dbg)
        assert np.array_equal(yy[0, 1 + y.shape[1]:].numpy(), g["codes" + tag][:6, 0].astype(np.int64))
        top2 = torch.stack([torch.topk(l[0], 2)[0] for l in dbg["logits"]]).numpy()
        np.testing.assert_allclose(top2, g["ar_top2" + tag][:7], rtol=0, atol=0)


@pytest.mark.parametrize("name,T", [("eos_mid", 37), ("eos_first", 0)])
def test_oracle_true_eos_termination(name, T):
    """EOS sampled by the model (models/vallex.py:574): row 1024 of the predict layer is crafted in the fixture so the
    REFERENCE stops after T tokens (T = 0: the very first sample is EOS, NAR runs on an empty target)."""
    g = load_golden(name)
    sd = dict(state_dict_for(g["seed"]))
    w = sd["ar_predict_layer.weight"].clone()
    w[1024] = torch.from_numpy(g["eos_row"])
    sd["ar_predict_layer.weight"] = w
    codes = _run(OracleVALLE(sd), g)[0].numpy()
    assert codes.shape == (T, 8) == g["codes"].shape
    assert np.array_equal(codes, g["codes"].astype(np.int64))


def test_golden_edge_cases_present():
    assert load_golden("noprompt_en")["prompt"].shape == (0, 8)       # no prompt: P = 0, Lp = 0
    assert bool(load_golden("mixlang")["text_language_is_list"])      # per-token language list
    for n in ("tiny_en", "noprompt_en", "mixlang", "mid_ja"):
        g = load_golden(n)
        assert g["codes"].shape[0] == 16 * len(g["x"])                # random weights never emit EOS -> cap
        assert g["codes"].min() >= 0 and g["codes"].max() < 1024


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present (GPU box)")
def test_oracle_live_against_reference():
    g = load_golden("noprompt_en")
    sd = state_dict_for(g["seed"])
    ref = ref_shim.load_reference_valle(sd)
    x = torch.from_numpy(g["x"].astype(np.int64))[None]
    y = torch.from_numpy(g["prompt"].astype(np.int64))[None]
    with torch.no_grad():
        want = ref.inference(x, torch.tensor([x.shape[1]]), y, enroll_x_lens=int(g["Lp"]), top_k=1, temperature=1.0,
                             prompt_language=g["prompt_language_arg"], text_language=g["text_language_arg"])
    got = _run(OracleVALLE(sd), g)
    assert torch.equal(want, got)
    # the synthetic checkpoint has exactly the reference's key set / shapes (strict load above) and its
    # tied tensors stay tied
    assert set(ref.state_dict().keys()) == set(sd.keys())


def test_vocos_oracle_istft_against_torch_istft():
    """The iSTFT of the Vocos restatement (V4, 'same' padding: trim (n_fft - hop)/2 = 480 per side) against torch's own
    istft (center=True: trims n_fft/2 = 640 per side): identical overlap-add and window-envelope normalisation, so
    torch.istft(S) == oracle(S)[160 : 160 + hop*(T-1)].  An independent pin for that part of the unpinned oracle."""
    import torch
    from oracle.vocos_oracle import HOP, N_FFT, OracleVocos
    from vallex_b200.weights import synthetic_vocos_state_dict
    ov = OracleVocos(synthetic_vocos_state_dict(0))
    g = torch.Generator().manual_seed(3)
    for B, T in ((1, 9), (2, 40)):
        S = torch.complex(torch.randn(B, N_FFT // 2 + 1, T, generator=g), torch.randn(B, N_FFT // 2 + 1, T, generator=g))
        S[:, 0].imag.zero_()
        S[:, -1].imag.zero_()          # DC and Nyquist bins of a real signal are real (irfft ignores their imaginary part)
        mine = ov.istft_same(S)
        ref = torch.istft(S, n_fft=N_FFT, hop_length=HOP, win_length=N_FFT, window=ov.sd["head.istft.window"], center=True)
        assert mine.shape == (B, HOP * T) and ref.shape == (B, HOP * (T - 1))
        off = (N_FFT // 2) - (N_FFT - HOP) // 2
        torch.testing.assert_close(mine[:, off:off + ref.shape[1]], ref, rtol=1e-4, atol=1e-4)
