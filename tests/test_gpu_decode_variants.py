"""GPU tests of the decode kernels: both batch-1 kernels (cluster kernel = default, 3-barrier kernel = fallback), the batched tensor-core kernel (default for
5 <= B <= 32 at bf16: mma.sync GEMV phases with the batch as N and bf16 hi/lo activations, balanced warp-level attention
phase over the flattened key space; VX_DECODE_MMA=0 selects the CUDA-core GEMV kernel) and the tcgen05 attention paths.
(The round-1 experiments -- two-barrier exchange over L2 flags, an 8-CTA-cluster variant that could not launch, TMA-staged weights,
trimmed variants, weight ring -- were measured on a B200 at the start of round 2, lost against the then-default kernel
(profiles/r02_experiments_call1.log) and were deleted; the 4-CTA-cluster kernel of decode_cluster.cu came later in round 2.)
Same bar as tests/test_gpu_parity.py::test_fused_decode_variant: teacher-forced agreement with the reference's
bf16-weight tokens wherever the reference's top-1/top-2 margin is binding."""
import os

import numpy as np
import pytest

from conftest import load_golden, state_dict_for
from test_gpu_parity import MARGIN, _valle

pytestmark = pytest.mark.gpu


def _teacher_forced(m, gs, names):
    eng = m.engine
    langs = [m._lang_ids(len(g["x"]), int(g["Lp"]), g["prompt_language_arg"], g["text_language_arg"]) for g in gs]
    eng.set_batch([g["x"].astype(np.int32) for g in gs], langs, [g["prompt"].astype(np.int32) for g in gs])
    eng.ar_prefill()
    wants = [g["codes_bf16w"].astype(np.int64)[:, 0] for g in gs]
    Tmax = max(len(w) for w in wants)
    forced = np.zeros((len(gs), Tmax), np.int32)
    for b, w in enumerate(wants):
        forced[b, :len(w)] = w
    toks, lens, _ = eng.ar_decode(max_new=Tmax, top_k=1, forced=forced, ignore_eos=True)
    for b, (g, w) in enumerate(zip(gs, wants)):
        T = len(w)
        margin = g["ar_top2_bf16w"][:T, 0] - g["ar_top2_bf16w"][:T, 1]
        match = toks[b, :T] == w
        assert match[margin > MARGIN].all(), (names[b], margin[~match & (margin > MARGIN)][:5])
        assert match.mean() > 0.97, (names[b], match.mean())
    kernel = eng.last_decode_kernel()
    eng.close()
    return kernel


@pytest.mark.parametrize("B", [8, 16, 32])
def test_tensor_core_batched_decode_variant(monkeypatch, B):
    names = (["tiny_en", "noprompt_en"] * 16)[:B]
    gs = [load_golden(n) for n in names]
    monkeypatch.delenv("VX_DECODE_MMA", raising=False)   # default path for B >= 8
    m = _valle(state_dict_for(gs[0]["seed"]), "bf16", max_ctx=512, max_batch=B)
    _teacher_forced(m, gs, names)


@pytest.mark.parametrize("B", [5, 8, 13])
def test_tensor_core_batched_decode_ragged_long(monkeypatch, B):
    """Ragged batch with very different context lengths (BASELINE config 1's S=122, P=225, T=1952 next to 6-phoneme utterances
    that finish after 80 / 96 tokens): the balanced attention phase cuts the flattened key space of all (utterance, head)
    pairs into equal warp ranges, so heads are split into many pieces for the long utterances and several heads share one
    warp for the short ones; finished utterances drop out of the plan."""
    names = (["cfg1_librispeech", "tiny_en", "noprompt_en", "tiny_en", "cfg1_librispeech", "noprompt_en"] * 3)[:B]
    gs = [load_golden(n) for n in names]
    monkeypatch.delenv("VX_DECODE_MMA", raising=False)
    monkeypatch.delenv("VX_DECODE_ATTN", raising=False)
    m = _valle(state_dict_for(0), "bf16", max_ctx=2400, max_batch=B)
    _teacher_forced(m, gs, names)


@pytest.mark.parametrize("t5", ["0", "1"])
@pytest.mark.parametrize("name", ["tiny_en", "mid_ja"])
def test_nar_and_prefill_attention_kernels(monkeypatch, name, t5):
    """Prefill + NAR through both attention kernels: the tcgen05/TMEM single-pass kernel (default; head operand images written
    by the QKV GEMM epilogue) and the mma.sync kernel (VX_ATTN_T5=0).  Same bar as test_bf16_teacher_forced_tokens."""
    g = load_golden(name)
    monkeypatch.setenv("VX_ATTN_T5", t5)
    m = _valle(state_dict_for(g["seed"]), "bf16", max_ctx=1024, max_batch=1)
    want = g["codes_bf16w"].astype(np.int64)
    T = want.shape[0]
    eng = m.engine
    lang = m._lang_ids(len(g["x"]), int(g["Lp"]), g["prompt_language_arg"], g["text_language_arg"])
    eng.set_batch([g["x"].astype(np.int32)], [lang], [g["prompt"].astype(np.int32)])
    eng.ar_prefill()
    toks, lens, _ = eng.ar_decode(top_k=1, forced=want[None, :, 0].astype(np.int32), ignore_eos=True)
    ar_margin = g["ar_top2_bf16w"][:T, 0] - g["ar_top2_bf16w"][:T, 1]
    ar_match = toks[0, :T] == want[:, 0]
    assert ar_match[ar_margin > MARGIN].all() and ar_match.mean() > 0.97     # exercises the prefill attention (KV cache)
    codes = eng.nar(cb0=want[None, :, 0].astype(np.int32), t_lens=np.asarray([T], np.int32))[0]
    nar_margin = (g["nar_top2_bf16w"][..., 0] - g["nar_top2_bf16w"][..., 1]).T
    nar_match = codes[:, 1:] == want[:, 1:]
    upstream_ok = np.cumprod(np.concatenate([np.ones((T, 1), bool), nar_match[:, :-1]], axis=1), axis=1).astype(bool)
    binding = nar_margin > MARGIN
    assert nar_match.mean() > 0.90
    assert nar_match[binding & upstream_ok.all(axis=1, keepdims=True)].mean() > 0.995
    eng.close()



@pytest.mark.parametrize("flag,kernel", [("0", "ar_decode_fused_kernel<1>"), ("1", "ar_decode_cluster_kernel")])
def test_batch1_kernels(monkeypatch, flag, kernel):
    """The batch-1 default is the cluster kernel (decode_cluster.cu: a head = a 4-CTA cluster, two grid barriers per layer);
    VX_DECODE_CLUSTER=0 selects the 3-barrier kernel it replaced (also the fallback where 16 four-CTA clusters cannot be
    co-resident).  Same teacher-forced bar for both, and the engine reports which kernel ran."""
    monkeypatch.setenv("VX_DECODE_CLUSTER", flag)
    for name in ("tiny_en", "mid_ja"):
        g = load_golden(name)
        m = _valle(state_dict_for(g["seed"]), "bf16", max_ctx=1024)
        assert _teacher_forced(m, [g], [name]) == kernel
