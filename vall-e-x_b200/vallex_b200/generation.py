"""Drop-in for the reference's `utils.generation` API (utils/generation.py:50-276):
`preload_models()`, `generate_audio(text, prompt, language, accent)`,
`generate_audio_from_long_text(text, prompt, language, accent, mode)` with the same argument meaning,
prompt lookup order (path -> ./presets/{name}.npz -> ./customs/{name}.npz), npz format and error types.

Out of scope by design: the rule-based G2P front end and the sentence
splitter are CPU string processing with third-party dependencies.  `text_tokenizer` is therefore
pluggable: install the reference's `PhonemeBpeTokenizer` with `set_text_tokenizer(...)`, or pass
pre-tokenised phonemes -- `text` may be a `(phone_ids, langs)` pair or an int array of phoneme ids.
"""
from __future__ import annotations

import logging
import os
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .macros import *  # noqa: F401,F403  (same star-import surface as the reference module)
from .macros import NUM_QUANTIZERS, code2lang, lang2token, langdropdown2token, token2lang
from .vallex import VALLE
from .vocos import Vocos
from .weights import synthetic_state_dict

device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
checkpoints_dir = "./checkpoints/"
model_checkpoint_name = "vallex-checkpoint.pt"

model: Optional[VALLE] = None
codec = None  # the EnCodec encoder is only used to MAKE prompts (data/tokenizer.py); not on this path
vocos: Optional[Vocos] = None
text_tokenizer: Optional[Callable[[str], Tuple[List[int], List[str]]]] = None
sentence_splitter: Optional[Callable[[str], List[str]]] = None


def set_text_tokenizer(fn):
    """fn(text: str) -> (phone_ids, langs), e.g. the reference's PhonemeBpeTokenizer(...).tokenize."""
    global text_tokenizer
    text_tokenizer = fn


def set_sentence_splitter(fn):
    global sentence_splitter
    sentence_splitter = fn


def preload_models(dtype: str = "bf16", max_batch: int = 8, max_ctx: int = 6144):
    """utils/generation.py:50-89.  Reads ./checkpoints/vallex-checkpoint.pt (dict with key "model").
    No download is attempted (no network): a missing checkpoint raises, unless VALLEX_B200_SYNTHETIC=1
    asks for seeded random-init weights of the same architecture (benchmarks / tests)."""
    global model, codec, vocos
    path = os.path.join(checkpoints_dir, model_checkpoint_name)
    if os.path.isfile(path):
        sd = torch.load(path, map_location="cpu")["model"]
    elif os.environ.get("VALLEX_B200_SYNTHETIC", "0") == "1":
        sd = synthetic_state_dict(0)
    else:
        raise Exception(f"\n Model weights not found at {path}; download vallex-checkpoint.pt manually and put it there.")
    model = VALLE(N_DIM, NUM_HEAD, NUM_LAYERS, norm_first=True, add_prenet=False, prefix_mode=PREFIX_MODE,
                  share_embedding=True, nar_scale_factor=1.0, prepend_bos=True, num_quantizers=NUM_QUANTIZERS,
                  dtype=dtype, max_batch=max_batch, max_ctx=max_ctx).to(device)
    missing_keys, unexpected_keys = model.load_state_dict(sd, strict=True)
    assert not missing_keys
    model.eval()
    vocos = Vocos.from_pretrained("charactr/vocos-encodec-24khz").to(device).share_engine(model)


def _load_prompt(prompt):
    """utils/generation.py:103-123."""
    if prompt is not None:
        prompt_path = prompt
        if not os.path.exists(prompt_path):
            prompt_path = "./presets/" + prompt + ".npz"
        if not os.path.exists(prompt_path):
            prompt_path = "./customs/" + prompt + ".npz"
        if not os.path.exists(prompt_path):
            raise ValueError(f"Cannot find prompt {prompt}")
        d = np.load(prompt_path)
        audio_prompts = torch.tensor(d["audio_tokens"]).type(torch.int32)
        text_prompts = torch.tensor(d["text_tokens"]).type(torch.int32)
        return audio_prompts, text_prompts, code2lang[int(d["lang_code"])]
    return torch.zeros([1, 0, NUM_QUANTIZERS]).type(torch.int32), torch.zeros([1, 0]).type(torch.int32), None


def _tokenize(text, lang_token: str):
    """utils/generation.py:98-100,127: text = lang_token + text + lang_token; tokenize("_" + text)."""
    if isinstance(text, tuple):  # (phone_ids, langs) already tokenised
        return list(text[0]), list(text[1])
    if not isinstance(text, str):  # int array of phoneme ids
        ids = [int(t) for t in np.asarray(text).reshape(-1)]
        return ids, None
    if text_tokenizer is None:
        raise RuntimeError("no text tokenizer installed: call set_text_tokenizer(PhonemeBpeTokenizer(...).tokenize) "
                           "or pass phoneme ids (the G2P front end is outside this engine's scope)")
    phone_tokens, langs = text_tokenizer(f"_{lang_token + text + lang_token}".strip())
    if len(phone_tokens) == 0:
        raise ValueError("Empty text is given")  # utils/g2p/__init__.py:23-24
    return phone_tokens, langs


def _prepare_utterance(text, text_prompts, lang_pr, language, accent):
    """Token / language bookkeeping of utils/generation.py:112-146 -> (text_tokens [1,S], Lp, prompt_language, text_language)."""
    lang_token = lang2token[language]
    lang = token2lang[lang_token]
    if lang_pr is None:
        lang_pr = lang if lang != "mix" else "en"
    enroll_x_lens = text_prompts.shape[-1]
    if isinstance(text, str):
        logging.info(f"synthesize text: {lang_token + text + lang_token}")
    phone_tokens, langs = _tokenize(text, lang_token)
    if langs is None:
        langs = [lang if lang != "mix" else "en"] * len(phone_tokens)
    text_tokens = torch.cat([text_prompts.long(), torch.tensor([phone_tokens], dtype=torch.long)], dim=-1)
    lang = lang if accent == "no-accent" else token2lang[langdropdown2token[accent]]
    return text_tokens, enroll_x_lens, lang_pr, (langs if accent == "no-accent" else lang)


def _one_utterance(text, audio_prompts, text_prompts, lang_pr, language, accent, **infer_kw):
    text_tokens, enroll_x_lens, lang_pr, text_language = _prepare_utterance(text, text_prompts, lang_pr, language, accent)
    text_tokens_lens = torch.tensor([text_tokens.shape[-1]], dtype=torch.int32)
    kw = dict(top_k=-100, temperature=1)   # utils/generation.py:142-143
    kw.update(infer_kw)
    encoded_frames = model.inference(text_tokens, text_tokens_lens, audio_prompts, enroll_x_lens=enroll_x_lens,
                                     prompt_language=lang_pr, text_language=text_language, **kw)
    return encoded_frames, text_tokens, enroll_x_lens


def _detect_language(text) -> str:
    try:
        import langid  # optional, like the reference (utils/generation.py:6-7,96-97)
        langid.set_languages(["en", "zh", "ja"])
        return langid.classify(text)[0]
    except ImportError as e:
        raise RuntimeError("language='auto' needs the `langid` package; pass language explicitly") from e


@torch.no_grad()
def generate_audio(text, prompt=None, language="auto", accent="no-accent", **infer_kw):
    """utils/generation.py:91-152 -> float32 waveform [320*T] at 24 kHz."""
    if isinstance(text, str):
        text = text.replace("\n", "").strip(" ")
    if language == "auto":
        language = _detect_language(text)
    audio_prompts, text_prompts, lang_pr = _load_prompt(prompt)
    encoded_frames, _, _ = _one_utterance(text, audio_prompts, text_prompts, lang_pr, language, accent, **infer_kw)
    frames = encoded_frames.permute(2, 0, 1)
    features = vocos.codes_to_features(frames)
    samples = vocos.decode(features, bandwidth_id=torch.tensor([2]))
    return samples.squeeze().cpu().numpy()


@torch.no_grad()
def generate_audio_from_long_text(text, prompt=None, language="auto", accent="no-accent", mode="sliding-window",
                                  batch_sentences=None, **infer_kw):
    """utils/generation.py:154-276.  `text` may be a list of pre-split sentences (each a str or phoneme ids).
    `batch_sentences` (engine extension; default: on in 'fixed-prompt' mode when the model offers `inference_batch`):
    the sentences of a fixed-prompt request share nothing but the prompt (utils/generation.py:199-224), so they are
    decoded as one ragged batch instead of one after another.  'sliding-window' is sequential by construction."""
    if mode not in ("fixed-prompt", "sliding-window"):
        raise ValueError(f"No such mode {mode}")
    if prompt is None or prompt == "":
        mode = "sliding-window"  # utils/generation.py:162-163
    if isinstance(text, (list, tuple)) and not (len(text) == 2 and isinstance(text[1], (list, tuple)) and
                                                 all(isinstance(t, str) for t in text[1])):
        sentences = list(text)
    else:
        if sentence_splitter is None:
            raise RuntimeError("no sentence splitter installed: call set_sentence_splitter(split_text_into_sentences) "
                               "or pass a list of sentences")
        sentences = sentence_splitter(text)
    if language == "auto":
        language = _detect_language(text if isinstance(text, str) else str(sentences[0]))
    audio_prompts, text_prompts, lang_pr = _load_prompt(prompt if prompt else None)
    complete_tokens = torch.zeros([1, NUM_QUANTIZERS, 0]).type(torch.LongTensor)
    cleaned = []
    for sent in sentences:
        if isinstance(sent, str):
            sent = sent.replace("\n", "").strip(" ")
            if sent == "":
                continue
        cleaned.append(sent)
    sentences = cleaned
    if batch_sentences is None:
        batch_sentences = mode == "fixed-prompt" and hasattr(model, "inference_batch") and len(sentences) > 1
    if batch_sentences:
        if mode != "fixed-prompt":
            raise ValueError("batch_sentences needs mode='fixed-prompt' (sliding-window feeds each sentence's output to the next)")
        items = []
        for sent in sentences:
            text_tokens, enroll, lp, tl = _prepare_utterance(sent, text_prompts, lang_pr, language, accent)
            items.append((text_tokens, audio_prompts, enroll, lp, tl))
        kw = dict(top_k=-100, temperature=1)
        kw.update(infer_kw)
        for encoded_frames in model.inference_batch(items, **kw):
            complete_tokens = torch.cat([complete_tokens, encoded_frames.transpose(2, 1).cpu()], dim=-1)
        sentences = []
    if lang_pr is None:  # no prompt: utils/generation.py:192 (fixed for the whole request)
        lang_pr = language if language != "mix" else "en"
    original_audio_prompts, original_text_prompts = audio_prompts, text_prompts  # utils/generation.py:230-231
    for sent in sentences:
        encoded_frames, text_tokens, enroll = _one_utterance(sent, audio_prompts, text_prompts, lang_pr, language,
                                                             accent, **infer_kw)
        complete_tokens = torch.cat([complete_tokens, encoded_frames.transpose(2, 1).cpu()], dim=-1)
        if mode == "sliding-window":  # utils/generation.py:264-269: a coin flip per sentence, hand-off or reset
            if torch.rand(1) < 0.5:
                audio_prompts = encoded_frames[:, :, -NUM_QUANTIZERS:].type(torch.int32).cpu()
                text_prompts = text_tokens[:, enroll:].type(torch.int32)
            else:
                audio_prompts = original_audio_prompts
                text_prompts = original_text_prompts
    # one Vocos pass over the concatenation (utils/generation.py:225-227,271-273)
    features = vocos.codes_to_features(complete_tokens.permute(1, 0, 2))
    samples = vocos.decode(features, bandwidth_id=torch.tensor([2]))
    return samples.squeeze().cpu().numpy()
