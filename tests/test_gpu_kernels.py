"""GPU: isolated kernel checks against a plain PyTorch fp32 reference of the same op (operands rounded to bf16
exactly as the kernel sees them)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("M,N,K,epi", [(128, 128, 64, 3), (200, 256, 1024, 3), (1814, 3072, 1024, 0), (333, 1024, 4096, 1),
                                       (700, 4096, 1024, 2), (50, 1024, 1024, 3), (1280, 384, 128, 3)])
def test_tcgen05_gemm(M, N, K, epi):
    from vallex_b200 import _capi
    lib = _capi.lib()
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g)
    C0 = torch.randn(M, N, generator=g)
    C = C0.clone().contiguous()
    _capi.check(lib.vx_debug_gemm_tc(_capi.ptr(A), _capi.ptr(W), _capi.ptr(bias), _capi.ptr(C), M, N, K, epi), "gemm_tc")
    ref = (_bf(A).double() @ _bf(W).double().T + bias.double())
    if epi == 1:
        ref = ref + C0.double()
    if epi == 2:
        ref = ref.clamp_min(0)
    tol = 2e-2 if epi in (0, 2) else 2e-3  # bf16 output rounding vs fp32 output
    err = (C.double() - ref).abs().max().item()
    print(f"gemm_tc M={M} N={N} K={K} epi={epi}: max abs err {err:.3e}")
    assert err < tol * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("variant", ["tc", "t5"])
@pytest.mark.parametrize("S,length,mode", [(6, 12, 1), (30, 200, 1), (122, 348, 1), (20, 64, 0), (40, 300, 0), (314, 1814, 0)])
def test_tensor_core_attention(S, length, mode, variant):
    """variant t5: tcgen05/TMEM single-pass attention (default for prefill + NAR in bf16 mode); tc: the mma.sync flash attention
    (VX_ATTN_T5=0)."""
    from vallex_b200 import _capi
    lib = _capi.lib()
    g = torch.Generator().manual_seed(S * 1000 + length)
    qkv = torch.randn(length, 3072, generator=g)
    out = torch.zeros(length, 1024)
    fn = lib.vx_debug_attn_tc if variant == "tc" else lib.vx_debug_attn_t5
    _capi.check(fn(_capi.ptr(qkv), _capi.ptr(out), S, length, mode), "attn_" + variant)
    q, k, v = [_bf(t).view(length, 16, 64).transpose(0, 1).double() for t in qkv.split(1024, dim=-1)]
    att = (q @ k.transpose(-2, -1)) / 8.0
    if mode == 1:
        i = torch.arange(length)[:, None]
        j = torch.arange(length)[None, :]
        att = att.masked_fill(~((j < S) | (j <= i)), float("-inf"))
    ref = (torch.softmax(att, -1) @ v).transpose(0, 1).reshape(length, 1024)
    err = (out.double() - ref).abs().max().item()
    print(f"attn_{variant} S={S} len={length} mode={mode}: max abs err {err:.3e}")
    assert err < 3e-2


def test_engine_rejects_out_of_range_ids():
    """the kernels index embedding tables with the caller's ids; the C ABI must refuse ids the reference's
    F.embedding would raise IndexError for (models/vallex.py:497-507,605-632) instead of reading out of bounds."""
    from conftest import state_dict_for
    from vallex_b200._capi import VallexB200Error
    from vallex_b200.engine import Engine
    from vallex_b200.weights import synthetic_vocos_state_dict
    eng = Engine(0, 2, 128, "fp32")
    eng.load_valle_state_dict(state_dict_for(0))
    eng.load_vocos_state_dict(synthetic_vocos_state_dict(0))
    eng.finalize()
    good_t, good_l = np.arange(5, 11, dtype=np.int32), np.zeros(6, np.int32)
    good_p = np.zeros((4, 8), np.int32)
    for bad in ("text", "lang", "prompt"):
        t, l, p = good_t.copy(), good_l.copy(), good_p.copy()
        if bad == "text":
            t[3] = 2048
        elif bad == "lang":
            l[2] = 3
        else:
            p[1, 7] = 1024
        with pytest.raises(VallexB200Error, match="out of range"):
            eng.set_batch([good_t, t], [good_l, l], [good_p, p])
    eng.set_batch([good_t], [good_l], [good_p])          # padded tails beyond the lengths are not inspected
    eng.ar_prefill()
    eng.ar_decode(max_new=3, top_k=1)
    with pytest.raises(VallexB200Error, match="cb0"):
        eng.nar(cb0=np.asarray([[1, 5000, 2]], np.int32), t_lens=np.asarray([3], np.int32))
    codes = eng.nar(cb0=np.asarray([[1, 2, 3]], np.int32), t_lens=np.asarray([3], np.int32))
    assert codes.shape == (1, 3, 8) and codes.min() >= 0 and codes.max() < 1024
    bad_codes = np.ascontiguousarray(codes.transpose(2, 0, 1)).copy()
    bad_codes[5, 0, 1] = -7
    with pytest.raises(VallexB200Error, match="code out of range"):
        eng.vocos_decode(bad_codes)
    wave = eng.vocos_decode(np.ascontiguousarray(codes.transpose(2, 0, 1)))
    assert np.isfinite(wave).all()
    eng.close()
    with pytest.raises(VallexB200Error):
        Engine(0, 5000, 128, "fp32")                      # max_batch bound (pinned scratch holds max_batch + 1 ints)
