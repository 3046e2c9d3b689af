"""TEST INFRASTRUCTURE ONLY: import the UNMODIFIED reference `models.vallex.VALLE`.

Works only where the reference tree exists (this authoring container: /root/reference).
The GPU box has no /root/reference, so nothing under `-m gpu`, `smoke()` or `bench.py`
may call this; it is used by `tests/golden/make_golden.py` (fixture generation) and by the
CPU-only live cross-check in `tests/test_oracle_golden.py` (skipped when absent).

The only obstacle to importing the reference here is `models/visualizer.py:21`
(`import matplotlib.pyplot`), pulled in by `models/vallex.py:36`; a stub module is enough
(SURVEY.md 8(c)).  Nothing is copied from the reference tree.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("VALLEX_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "vallex.py"))


def load_reference_valle(state_dict):
    """Build the reference VALLE exactly as utils/generation.py:67-78 does and load `state_dict` strict."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from models.vallex import VALLE  # noqa: E402  (the reference's own class)

    model = VALLE(1024, 16, 12, norm_first=True, add_prenet=False, prefix_mode=1, share_embedding=True,
                  nar_scale_factor=1.0, prepend_bos=True, num_quantizers=8)
    missing, unexpected = model.load_state_dict(state_dict, strict=True)
    assert not missing and not unexpected
    model.eval()
    return model
