"""Drop-in for the reference's `utils.prompt_making.make_prompt` (utils/prompt_making.py:57-84): turn an audio clip
(+ transcript) into the `.npz` prompt `generate_audio(prompt=...)` reads -- `audio_tokens` int64 [1,T,8] (EnCodec codes,
computed on the B200 by `tokenizer.AudioTokenizer`), `text_tokens` [1,L], `lang_code`.

Outside this engine's scope, and therefore pluggable exactly like `generation.text_tokenizer`:
the G2P front end (`set_text_tokenizer`, shared with `generation`), the language identifier the reference takes from `langid`
(`set_language_identifier`) and the Whisper transcriber used when no transcript is given (`set_transcriber`).
"""
from __future__ import annotations

import logging
import os
from typing import Callable, Optional, Tuple

import numpy as np
import torch

from . import generation
from .macros import lang2code, lang2token
from .tokenizer import AudioTokenizer, tokenize_audio

codec: Optional[AudioTokenizer] = None
language_identifier: Optional[Callable[[str], str]] = None          # text -> 'en' | 'zh' | 'ja'   (langid.classify)
transcriber: Optional[Callable[[torch.Tensor, int], Tuple[str, str]]] = None   # (wav, sr) -> (lang, text)  (whisper)


def set_language_identifier(fn):
    global language_identifier
    language_identifier = fn


def set_transcriber(fn):
    global transcriber
    transcriber = fn


def _codec() -> AudioTokenizer:
    global codec
    if codec is None:
        codec = AudioTokenizer(generation.device)
    return codec


def make_transcript(name, wav, sr, transcript=None):
    """utils/prompt_making.py:87-120: returns (lang_token + text + lang_token, lang)."""
    if not isinstance(wav, torch.FloatTensor):
        wav = torch.tensor(wav)
    if wav.abs().max() > 1:
        wav /= wav.abs().max()
    if wav.size(-1) == 2:
        wav = wav.mean(-1, keepdim=False)
    if wav.ndim == 1:
        wav = wav.unsqueeze(0)
    assert wav.ndim and wav.size(0) == 1
    if transcript is None or transcript == "":
        if transcriber is None:
            raise RuntimeError("no transcript given and no transcriber installed (the reference runs Whisper here, "
                               "utils/prompt_making.py:99-110): call set_transcriber(fn) or pass transcript=")
        lang, text = transcriber(wav, sr)
        if text.strip(" ")[-1] not in "?!.,。，？！。、":
            text += "."
    else:
        text = transcript
        if language_identifier is None:
            raise RuntimeError("no language identifier installed (the reference uses langid.classify, "
                               "utils/prompt_making.py:113): call set_language_identifier(fn)")
        lang = language_identifier(text)
    lang_token = lang2token[lang]
    return lang_token + text + lang_token, lang


def make_prompt(name, audio_prompt_path, transcript=None, save_dir: str = "./customs/"):
    """utils/prompt_making.py:57-84.  `audio_prompt_path` may also be a (wav [C,N], sample_rate) pair."""
    if isinstance(audio_prompt_path, str):
        import torchaudio

        wav_pr, sr = torchaudio.load(audio_prompt_path)
    else:
        wav_pr, sr = audio_prompt_path
    if wav_pr.size(-1) / sr > 15:
        raise ValueError(f"Prompt too long, expect length below 15 seconds, got {wav_pr.size(-1) / sr} seconds.")
    if wav_pr.size(0) == 2:
        wav_pr = wav_pr.mean(0, keepdim=True)
    if isinstance(transcript, tuple):          # (phoneme ids, 'en' | 'zh' | 'ja'): already through the G2P front end
        phonemes, lang_pr = [int(t) for t in transcript[0]], transcript[1]
    else:
        text_pr, lang_pr = make_transcript(name, wav_pr, sr, transcript)
        if generation.text_tokenizer is None:
            raise RuntimeError("no text tokenizer installed: call generation.set_text_tokenizer(PhonemeBpeTokenizer(...).tokenize) "
                               "or pass transcript=(phoneme_ids, language)")
        phonemes, _langs = generation.text_tokenizer(f"{text_pr}".strip())     # utils/prompt_making.py:72

    encoded_frames = tokenize_audio(_codec(), (wav_pr, sr))
    audio_tokens = encoded_frames[0][0].transpose(2, 1).cpu().numpy()

    text_tokens = np.asarray([phonemes], dtype=np.int64)       # get_text_token_collater(): no bos/eos, ids as they are

    os.makedirs(save_dir, exist_ok=True)
    save_path = os.path.join(save_dir, f"{name}.npz")
    np.savez(save_path, audio_tokens=audio_tokens, text_tokens=text_tokens, lang_code=lang2code[lang_pr])
    logging.info(f"Successful. Prompt saved to {save_path}")
    return save_path
