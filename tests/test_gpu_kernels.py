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


@pytest.mark.parametrize("variant", ["tc", "t5", "t5max"])
@pytest.mark.parametrize("S,length,mode", [(6, 12, 1), (30, 200, 1), (122, 348, 1), (20, 64, 0), (40, 300, 0), (314, 1814, 0)])
def test_tensor_core_attention(S, length, mode, variant, monkeypatch):
    """variant tc: mma.sync flash attention (default); t5: tcgen05/TMEM two-pass attention (VX_ATTN_T5=1); t5max: the same
    with a max-only statistics pass (VX_ATTN_T5=2; written after the round-1 GPU budget ran out -> VX_TEST_EXPERIMENTAL=1)."""
    import os
    if variant == "t5max":
        if os.environ.get("VX_TEST_EXPERIMENTAL") != "1":
            pytest.skip("not yet run on a B200; set VX_TEST_EXPERIMENTAL=1")
        monkeypatch.setenv("VX_ATTN_T5", "2")
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
