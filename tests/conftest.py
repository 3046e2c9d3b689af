import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "vall-e-x_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["tiny_en", "noprompt_en", "mixlang", "mid_ja", "cfg1_librispeech", "longctx_en"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    g["text_language_arg"] = ([str(s) for s in g["text_language"]] if bool(g["text_language_is_list"])
                              else str(g["text_language"][0]))
    g["prompt_language_arg"] = str(g["prompt_language"])
    return g


_SD_CACHE = {}


def state_dict_for(seed, rounded=False):
    from vallex_b200.weights import round_state_dict_bf16, synthetic_state_dict
    key = (int(seed), bool(rounded))
    if key not in _SD_CACHE:
        base = _SD_CACHE.get((int(seed), False))
        if base is None:
            base = synthetic_state_dict(int(seed))
            _SD_CACHE[(int(seed), False)] = base
        _SD_CACHE[key] = round_state_dict_bf16(base) if rounded else base
    return _SD_CACHE[key]


@pytest.fixture(scope="session")
def golden():
    return load_golden
