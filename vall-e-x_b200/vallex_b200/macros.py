"""Model constants and language maps for the generate_audio path.

Mirrors the values a caller of the reference sees in `macros.py:1-38` and
`models/macros.py:1-11` (same names, same numbers) so host code written against
the reference keeps working.
"""
NUM_LAYERS = 12
NUM_HEAD = 16
N_DIM = 1024
PREFIX_MODE = 1
NUM_QUANTIZERS = 8
SAMPLE_RATE = 24000

NUM_TEXT_TOKENS = 2048
NUM_AUDIO_TOKENS = 1024  # EnCodec RVQ bins; id 1024 = EOS/PAD, id 1025 = BOS

lang2token = {"zh": "[ZH]", "ja": "[JA]", "en": "[EN]", "mix": ""}
lang2code = {"zh": 0, "ja": 1, "en": 2}
token2lang = {"[ZH]": "zh", "[JA]": "ja", "[EN]": "en", "": "mix"}
code2lang = {0: "zh", 1: "ja", 2: "en"}
langdropdown2token = {"English": "[EN]", "中文": "[ZH]", "日本語": "[JA]", "Mix": ""}

# model-side language ids (models/vallex.py:439-443) -- NOT the npz lang_code numbering
LANGUAGE_ID = {"en": 0, "zh": 1, "ja": 2}

# derived sizes used by the engine
FFN_DIM = 4 * N_DIM
HEAD_DIM = N_DIM // NUM_HEAD
EOS_ID = NUM_AUDIO_TOKENS
BOS_ID = NUM_AUDIO_TOKENS + 1
FRAME_RATE = 75
HOP = 320
