"""CPU: the C-ABI library builds, loads and exports every symbol include/vallex_b200.h declares; the
product path fails loudly (no CPU fallback) when no B200 is present."""
import os
import re

import pytest

from conftest import ROOT


def _declared():
    hdr = open(os.path.join(ROOT, "include", "vallex_b200.h")).read()
    return sorted(set(re.findall(r"VX_API\s+[\w\s\*]+?\b(vx_\w+)\s*\(", hdr)))


def test_header_declares_expected_entry_points():
    names = _declared()
    for n in ["vx_create", "vx_destroy", "vx_load_weight", "vx_finalize_weights", "vx_set_batch", "vx_ar_prefill",
              "vx_ar_decode", "vx_nar", "vx_vocos_decode", "vx_last_error", "vx_version"]:
        assert n in names


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()  # no-op when up to date
    from vallex_b200 import _capi
    lib = _capi.lib()
    for n in _declared():
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _capi.SIGNATURES, f"{n} has no ctypes signature"
    assert b"sm_100a" in lib.vx_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vallex_b200._capi import VallexB200Error
    from vallex_b200.engine import Engine
    with pytest.raises(VallexB200Error):
        Engine(0, 1, 64, "fp32")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "vall-e-x_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
