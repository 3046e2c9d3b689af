"""CPU: host-side mirror of the reference API (argument checks, error types, language-id mapping,
prompt lookup) -- no engine calls."""
import numpy as np
import pytest
import torch

from vallex_b200 import generation, macros
from vallex_b200.vallex import VALLE
from vallex_b200.weights import is_bf16_matrix, sine_position_table, synthetic_state_dict, synthetic_vocos_state_dict


def _model():
    return VALLE(1024, 16, 12, norm_first=True, add_prenet=False, prefix_mode=1, share_embedding=True,
                 nar_scale_factor=1.0, prepend_bos=True, num_quantizers=8)


def test_constructor_matches_reference_call_and_rejects_other_configs():
    m = _model()
    assert m.language_ID == {"en": 0, "zh": 1, "ja": 2}          # models/vallex.py:439-443
    assert m.eval() is m
    with pytest.raises(NotImplementedError):
        VALLE(512, 8, 6)
    with pytest.raises(NotImplementedError):
        m.forward()                                               # models/vallex.py:447-457
    with pytest.raises(RuntimeError):
        m.to("cpu")


def test_language_id_split():
    m = _model()
    ids = m._lang_ids(6, 2, "zh", "en")
    assert ids.tolist() == [1, 1, 0, 0, 0, 0]
    ids = m._lang_ids(5, 2, "en", ["ja", "zh", "en"])
    assert ids.tolist() == [0, 0, 2, 1, 0]
    with pytest.raises(RuntimeError):
        m._lang_ids(5, 2, "en", ["ja", "zh"])
    # npz lang_code numbering differs from the model numbering (macros.py:15-32 vs models/vallex.py:439-443)
    assert macros.code2lang[2] == "en" and macros.LANGUAGE_ID["en"] == 0


def test_inference_asserts_like_reference():
    m = _model()
    x = torch.zeros(1, 4, dtype=torch.long)
    with pytest.raises(AssertionError):
        m.inference(x[0], torch.tensor([4]), torch.zeros(1, 0, 8), 0, prompt_language="en", text_language="en")
    with pytest.raises(AssertionError):
        m.inference(x, torch.tensor([4]), torch.zeros(2, 3, 8), 0, prompt_language="en", text_language="en")
    with pytest.raises(AssertionError):
        m.inference(x, torch.tensor([0]), torch.zeros(1, 3, 8), 0, prompt_language="en", text_language="en")


def test_load_state_dict_is_strict_at_call_time():
    """utils/generation.py:79-83: `missing_keys, unexpected_keys = model.load_state_dict(checkpoint["model"], strict=True)`."""
    m = _model()
    sd = synthetic_state_dict(1)
    r = m.load_state_dict(sd, strict=True)
    assert r.missing_keys == [] and r.unexpected_keys == []
    bad = dict(sd)
    del bad["ar_predict_layer.weight"]
    bad["bogus.weight"] = torch.zeros(1)
    with pytest.raises(RuntimeError, match="Missing key"):
        m.load_state_dict(bad, strict=True)
    r = m.load_state_dict(bad, strict=False)                       # keeps the previously loaded tensor for the missing key
    assert r.missing_keys == ["ar_predict_layer.weight"] and r.unexpected_keys == ["bogus.weight"]
    assert m._sd["ar_predict_layer.weight"] is sd["ar_predict_layer.weight"] and "bogus.weight" not in m._sd
    wrong = dict(sd)
    wrong["ar_decoder.norm.weight"] = torch.zeros(7)
    with pytest.raises(RuntimeError, match="size mismatch"):
        m.load_state_dict(wrong, strict=False)


def test_out_of_range_ids_raise_index_error_like_embedding():
    """F.embedding raises IndexError in the reference (models/vallex.py:497,529,605-632); the engine must never see such ids."""
    m = _model()
    m.load_state_dict(synthetic_state_dict(1))
    kw = dict(prompt_language="en", text_language="en")
    with pytest.raises(IndexError):
        m.inference(torch.tensor([[5, 2048]]), torch.tensor([2]), torch.zeros(1, 3, 8, dtype=torch.long), 0, **kw)
    with pytest.raises(IndexError):
        m.inference(torch.tensor([[5, 6]]), torch.tensor([2]), torch.full((1, 3, 8), 1024, dtype=torch.long), 0, **kw)
    with pytest.raises(IndexError):
        m.inference(torch.tensor([[5, -1]]), torch.tensor([2]), torch.zeros(1, 3, 8, dtype=torch.long), 0, **kw)
    from vallex_b200.vocos import Vocos
    with pytest.raises(IndexError):
        Vocos(synthetic_vocos_state_dict(0)).codes_to_features(torch.full((8, 1, 4), 1024, dtype=torch.long))


def test_generation_errors(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    with pytest.raises(ValueError, match="Cannot find prompt"):
        generation._load_prompt("no_such_prompt")
    with pytest.raises(ValueError, match="No such mode"):
        generation.generate_audio_from_long_text([np.arange(5, 9)], prompt="x", language="en", mode="bogus")
    # npz prompt format is read unchanged (utils/generation.py:111-119)
    np.savez(tmp_path / "p.npz", audio_tokens=np.zeros((1, 7, 8), np.int64), text_tokens=np.full((1, 3), 9, np.int64),
             lang_code=np.int32(1))
    a, t, lang = generation._load_prompt(str(tmp_path / "p.npz"))
    assert a.shape == (1, 7, 8) and a.dtype == torch.int32 and t.shape == (1, 3) and lang == "ja"
    a, t, lang = generation._load_prompt(None)
    assert a.shape == (1, 0, 8) and t.shape == (1, 0) and lang is None


def test_synthetic_checkpoint_layout():
    sd = synthetic_state_dict(3)
    assert len(sd) == 374
    assert sum(v.numel() for k, v in sd.items() if not k.startswith("nar_predict_layers") or k.endswith("6.weight")) == 370539524
    assert sd["nar_predict_layers.0.weight"] is sd["nar_audio_embeddings.2.word_embeddings.weight"]
    assert is_bf16_matrix("ar_decoder.layers.3.linear2.weight") and is_bf16_matrix("nar_audio_embeddings.5.word_embeddings.weight")
    assert not is_bf16_matrix("ar_text_embedding.word_embeddings.weight") and not is_bf16_matrix("ar_decoder.layers.0.norm1.weight")
    v = synthetic_vocos_state_dict(0)
    assert v["head.out.weight"].shape == (1282, 384) and v["backbone.embed.weight"].shape == (384, 128, 7)
    pe = sine_position_table(16)
    assert pe.shape == (16, 1024) and float(pe[0, 1]) == 1.0 and float(pe[0, 0]) == 0.0


class _StubModel:
    """Records the calls generate_audio_from_long_text makes; emits T = number of new phonemes frames of constant codes."""
    def __init__(self, with_batch=True):
        self.calls = []
        if not with_batch:
            self.inference_batch = None
            del self.inference_batch

    def _frames(self, x, enroll):
        T = int(x.shape[1]) - int(enroll)
        return torch.full((1, T, 8), int(x[0, -1]), dtype=torch.long)

    def inference(self, x, x_lens, y, enroll_x_lens, **kw):
        self.calls.append(("single", int(x.shape[1]), int(y.shape[1]), kw.get("top_k"), kw.get("prompt_language")))
        return self._frames(x, enroll_x_lens)


class _StubBatchModel(_StubModel):
    def inference_batch(self, items, **kw):
        self.calls.append(("batch", len(items), kw.get("top_k")))
        return [self._frames(x, enroll) for x, y, enroll, lp, tl in items]


class _StubVocos:
    def codes_to_features(self, codes):
        self.codes = codes
        return codes

    def decode(self, features, bandwidth_id=None):
        return features[0].float().repeat_interleave(320, dim=-1)[:1]


def test_long_text_fixed_prompt_batches_sentences(tmp_path, monkeypatch):
    """utils/generation.py:199-227: fixed-prompt sentences are independent -> one ragged batch; the concatenated codes and
    the single Vocos pass are the same as the sequential path's."""
    np.savez(tmp_path / "p.npz", audio_tokens=np.zeros((1, 7, 8), np.int64), text_tokens=np.full((1, 3), 9, np.int64),
             lang_code=np.int32(2))
    sents = [np.arange(5, 9), np.arange(10, 16), np.arange(20, 23)]
    results = {}
    for name, mdl, kw in (("batched", _StubBatchModel(), {}), ("sequential", _StubBatchModel(), {"batch_sentences": False}),
                          ("no_batch_api", _StubModel(), {})):
        voc = _StubVocos()
        monkeypatch.setattr(generation, "model", mdl, raising=False)
        monkeypatch.setattr(generation, "vocos", voc, raising=False)
        wav = generation.generate_audio_from_long_text(sents, prompt=str(tmp_path / "p.npz"), language="en",
                                                       mode="fixed-prompt", **kw)
        results[name] = (mdl.calls, voc.codes.clone(), wav)
    assert results["batched"][0] == [("batch", 3, -100)]
    assert [c[0] for c in results["sequential"][0]] == ["single"] * 3
    assert results["sequential"][0][0] == ("single", 3 + 4, 7, -100, "en")
    assert [c[0] for c in results["no_batch_api"][0]] == ["single"] * 3
    for k in ("sequential", "no_batch_api"):
        assert torch.equal(results["batched"][1], results[k][1])
        assert np.array_equal(results["batched"][2], results[k][2])
    assert results["batched"][1].shape == (8, 1, 4 + 6 + 3)
    # sliding-window is sequential by construction
    monkeypatch.setattr(generation, "model", _StubBatchModel(), raising=False)
    with pytest.raises(ValueError, match="fixed-prompt"):
        generation.generate_audio_from_long_text(sents, prompt=str(tmp_path / "p.npz"), language="en",
                                                 mode="sliding-window", batch_sentences=True)


class _RecordingModel:
    """Records (x, y, enroll, prompt_language) of every inference call; output frames depend on the call index so a
    handed-off prompt is recognisable."""
    def __init__(self):
        self.calls = []

    def inference(self, x, x_lens, y, enroll_x_lens, **kw):
        k = len(self.calls)
        self.calls.append((x.clone(), y.clone(), int(enroll_x_lens), kw.get("prompt_language"), int(x_lens[0])))
        T = int(x.shape[1]) - int(enroll_x_lens) + 1
        return (torch.arange(T * 8, dtype=torch.long).reshape(1, T, 8) + 100 * (k + 1)) % 1024


def _reference_sliding_window(sentences, audio_prompts, text_prompts, model_fn, seed):
    """The control flow of utils/generation.py:228-269 restated: (x, y, enroll) each sentence receives."""
    torch.manual_seed(seed)
    original_audio_prompts, original_text_prompts = audio_prompts, text_prompts
    seen = []
    for k, phones in enumerate(sentences):
        enroll = text_prompts.shape[-1]
        text_tokens = torch.cat([text_prompts.long(), torch.tensor([list(phones)], dtype=torch.long)], dim=-1)
        seen.append((text_tokens, audio_prompts, enroll))
        frames = model_fn(k, text_tokens, enroll)
        if torch.rand(1) < 0.5:
            audio_prompts = frames[:, :, -8:]
            text_prompts = text_tokens[:, enroll:]
        else:
            audio_prompts = original_audio_prompts
            text_prompts = original_text_prompts
    return seen


@pytest.mark.parametrize("with_prompt", [True, False])
def test_long_text_sliding_window_prompt_sequence(tmp_path, monkeypatch, with_prompt):
    """utils/generation.py:228-269: after every sentence a coin flip either hands the sentence's output over as the next
    prompt or RESETS to the original prompt.  The exact (text, audio prompt, enroll length, prompt language) sequence is
    compared with a restatement of the reference's loop under the same torch seed."""
    np.savez(tmp_path / "p.npz", audio_tokens=(np.arange(7 * 8).reshape(1, 7, 8) % 1024).astype(np.int64),
             text_tokens=np.full((1, 3), 9, np.int64), lang_code=np.int32(0))
    sents = [np.arange(5, 9), np.arange(10, 16), np.arange(20, 23), np.arange(30, 41), np.arange(12, 14), np.arange(40, 47),
             np.arange(5, 12), np.arange(21, 25)]
    prompt = str(tmp_path / "p.npz") if with_prompt else None
    a0, t0, lang0 = generation._load_prompt(prompt)
    for seed in (0, 1, 2, 3):
        mdl = _RecordingModel()
        monkeypatch.setattr(generation, "model", mdl, raising=False)
        monkeypatch.setattr(generation, "vocos", _StubVocos(), raising=False)
        torch.manual_seed(seed)
        generation.generate_audio_from_long_text(sents, prompt=prompt, language="ja", mode="sliding-window")

        def model_fn(k, x, enroll):
            T = int(x.shape[1]) - enroll + 1
            return (torch.arange(T * 8, dtype=torch.long).reshape(1, T, 8) + 100 * (k + 1)) % 1024
        want = _reference_sliding_window(sents, a0, t0, model_fn, seed)
        assert len(mdl.calls) == len(want) == len(sents)
        resets = handoffs = 0
        for k, ((x, y, enroll, lp, xl), (wx, wy, wenroll)) in enumerate(zip(mdl.calls, want)):
            assert torch.equal(x, wx) and torch.equal(y.long(), wy.long()) and enroll == wenroll and xl == wx.shape[1], (seed, k)
            assert lp == ("zh" if with_prompt else "ja")        # lang_pr never changes (utils/generation.py:182-192)
            if k > 0:
                if torch.equal(y.long(), a0.long()) and enroll == t0.shape[-1]:
                    resets += 1
                else:
                    handoffs += 1
        if seed == 0:
            assert resets > 0 and handoffs > 0                   # both branches of utils/generation.py:264-269 exercised
    # fixed-prompt: every sentence sees the original prompt (utils/generation.py:199-224)
    if with_prompt:
        mdl = _RecordingModel()
        monkeypatch.setattr(generation, "model", mdl, raising=False)
        generation.generate_audio_from_long_text(sents, prompt=prompt, language="ja", mode="fixed-prompt")
        assert all(torch.equal(y.long(), a0.long()) and enroll == 3 for _, y, enroll, _, _ in mdl.calls)


class _FakeEngine:
    """Stands in for vallex_b200.engine.Engine: utterance with S phonemes 'generates' T = S frames (token t of an utterance whose
    first phoneme id is f is (10 f + t) % 1024), one token per slot per step; records the batches, refills and launches."""
    def __init__(self):
        self.batches, self.seeds, self.refills, self.launches = [], [], [], []

    def _tok(self, text, t):
        return (int(text[0]) * 10 + t) % 1024

    def set_batch(self, texts, langs, prompts, Smax=0, Pmax=0):
        self.texts = [np.asarray(t) for t in texts]
        self.T = [0] * len(texts)
        self.fin = [0] * len(texts)
        self.caps = (Smax, Pmax)
        self.caps_hist = getattr(self, 'caps_hist', []) + [self.caps]
        self.batches.append([int(len(t)) for t in texts])

    def ar_prefill(self):
        pass

    def slot_prefill(self, slot, text, lang, prompt):
        assert self.fin[slot] == 1 and len(text) <= self.caps[0]
        self.refills.append((slot, int(len(text))))
        self.texts[slot], self.T[slot], self.fin[slot] = np.asarray(text), 0, 0

    def ctx_clamped(self):
        return False

    def timing(self):
        return {"ms_prefill": 1.0, "ms_decode": 2.0, "ms_nar": 3.0, "decode_steps": 4, "kernel_launches": 5}

    def _run(self, max_new, stop_on_finish):
        B = len(self.texts)
        budget = [min(len(t), max_new) if max_new else len(t) for t in self.texts]
        steps = max(budget)
        toks = np.zeros((B, steps), np.int32)
        start_fin = sum(self.fin)
        while not all(self.fin):
            for b in range(B):
                if not self.fin[b]:
                    toks[b, self.T[b]] = self._tok(self.texts[b], self.T[b])
                    self.T[b] += 1
                    if self.T[b] >= budget[b]:
                        self.fin[b] = 1
            if stop_on_finish and sum(self.fin) > start_fin and not all(self.fin):
                break
        return toks, np.asarray(self.T, np.int32), np.zeros(B, np.float32), np.asarray(self.fin, np.int32)

    def ar_decode(self, max_new=0, top_k=-100, top_p=1.0, temperature=1.0, seed=0, **kw):
        self.seeds.append(seed)
        toks, lens, slp, _ = self._run(max_new, False)
        for b in range(len(self.texts)):     # a plain launch returns the whole rows
            toks[b, :lens[b]] = [self._tok(self.texts[b], t) for t in range(lens[b])]
        return toks, lens, slp

    def ar_decode_ex(self, max_new=0, top_k=-100, top_p=1.0, temperature=1.0, seed=0, resume=False, stop_on_finish=False):
        self.seeds.append(seed)
        self.launches.append((resume, stop_on_finish))
        return self._run(max_new, stop_on_finish)       # like the device buffer: only THIS launch's tokens are present

    def nar(self, cb0=None, t_lens=None, Tmax=None):
        B, Tm = cb0.shape
        codes = np.zeros((B, Tm, 8), np.int32)
        codes[:, :, :] = cb0[:, :, None]
        return codes


def test_inference_batch_chunks_orders_and_lengths():
    """VALLE.inference_batch: per-item outputs in input order with their own lengths; with more items than slots the AR decode is
    continuously batched (a finished slot is refilled at once, the batch resumes), otherwise / with continuous=False it runs in
    max_batch-sized ragged chunks; the NAR stages run over the finished utterances in ragged batches; same argument checks as
    .inference (models/vallex.py:485-493)."""
    m = VALLE(1024, 16, 12, norm_first=True, add_prenet=False, prefix_mode=1, share_embedding=True, nar_scale_factor=1.0,
              prepend_bos=True, num_quantizers=8, max_batch=2)
    S = [5, 3, 7, 4, 6]
    items = [(torch.full((1, s), 10 + i, dtype=torch.long), torch.zeros(1, 2, 8, dtype=torch.long), 1, "en", "zh")
             for i, s in enumerate(S)]

    def check(outs):
        assert [tuple(o.shape) for o in outs] == [(1, s, 8) for s in S]
        for i, o in enumerate(outs):
            assert o.dtype == torch.int64
            assert o[0, :, 0].tolist() == [((10 + i) * 10 + t) % 1024 for t in range(S[i])]

    fake = _FakeEngine()
    m._engine = fake
    check(m.inference_batch(items, top_k=1, seed=100, quiet=True))
    # continuous: slots start with items 0 (5 frames) and 1 (3 frames); slot 1 frees first -> item 2, then slot 0 -> item 3, ...
    assert fake.batches[0] == [5, 3] and fake.caps_hist[0] == (7, 2)
    assert fake.refills == [(1, 7), (0, 4), (0, 6)]
    assert fake.launches[0] == (False, True) and all(r for r, _ in fake.launches[1:]) and fake.launches[-1][1] is False
    assert fake.batches[1:] == [[5, 3], [7, 4], [6]]                   # the NAR batches, input order
    assert m.last_timing["decode_launches"] == len(fake.launches) == 4
    fake = _FakeEngine()
    m._engine = fake
    check(m.inference_batch(items, top_k=1, seed=100, quiet=True, continuous=False))
    assert fake.batches == [[5, 3], [7, 4], [6], [5, 3], [7, 4], [6]] and fake.seeds == [100, 102, 104] and not fake.refills
    assert m.last_timing["ms_decode"] == 6.0 and m.last_timing["ms_prefill"] == 3.0     # summed over the three chunks
    short = m.inference_batch(items[:2], top_k=1, max_new_tokens=2, quiet=True)
    assert [tuple(o.shape) for o in short] == [(1, 2, 8), (1, 2, 8)]
    with pytest.raises(AssertionError):
        m.inference_batch([(torch.zeros(4, dtype=torch.long), torch.zeros(1, 2, 8), 1, "en", "en")], quiet=True)
    with pytest.raises(AssertionError):
        m.inference_batch([(torch.zeros(1, 4, dtype=torch.long), torch.zeros(2, 2, 8), 1, "en", "en")], quiet=True)
    # seed=None follows torch's global generator (reproducible under torch.manual_seed, like the reference's sampling)
    torch.manual_seed(7)
    fake.seeds.clear()
    m.inference_batch(items[:1], quiet=True)
    a = fake.seeds[0]
    torch.manual_seed(7)
    m.inference_batch(items[:1], quiet=True)
    assert fake.seeds[1] == a


# ---- prompt making host logic (utils/prompt_making.py:57-120, data/tokenizer.py:98-111) with a stub codec: no GPU ----
class _StubCodec:
    sample_rate, channels = 24000, 1

    def __init__(self):
        self.seen = None

    def encode(self, wav):
        self.seen = wav
        T = (wav.shape[-1] + 319) // 320
        codes = (torch.arange(8)[None, :, None] * 100 + torch.arange(T)[None, None, :]) % 1024
        return [(codes.to(torch.int64), None)]


def test_make_prompt_npz_layout_and_errors(tmp_path, monkeypatch):
    from vallex_b200 import generation, prompt_making
    stub = _StubCodec()
    monkeypatch.setattr(prompt_making, "codec", stub)
    stereo = torch.stack([torch.linspace(-1, 1, 4800), torch.linspace(1, -1, 4800)])      # mean = 0 everywhere
    path = prompt_making.make_prompt("p0", (stereo, 24000), transcript=([7, 8, 9], "ja"), save_dir=str(tmp_path))
    assert tuple(stub.seen.shape) == (1, 1, 4800) and float(stub.seen.abs().max()) < 1e-6   # stereo -> mono mean, [B,1,N]
    d = np.load(path)
    assert d["audio_tokens"].shape == (1, 15, 8) and d["audio_tokens"].dtype == np.int64     # [1,T,8] = codes.transpose(2,1)
    assert d["audio_tokens"][0, 3, 2] == 203 and d["text_tokens"].tolist() == [[7, 8, 9]] and int(d["lang_code"]) == 1
    audio_prompts, text_prompts, lang_pr = generation._load_prompt(path)                     # what generate_audio reads back
    assert tuple(audio_prompts.shape) == (1, 15, 8) and lang_pr == "ja"
    with pytest.raises(ValueError, match="Prompt too long"):                                 # utils/prompt_making.py:60-62
        prompt_making.make_prompt("p1", (torch.zeros(1, 15 * 24000 + 1), 24000), transcript=([5], "en"), save_dir=str(tmp_path))


def test_make_transcript_plumbing(monkeypatch):
    from vallex_b200 import prompt_making
    monkeypatch.setattr(prompt_making, "language_identifier", None)
    monkeypatch.setattr(prompt_making, "transcriber", None)
    wav = torch.zeros(1, 100)
    with pytest.raises(RuntimeError, match="language identifier"):
        prompt_making.make_transcript("n", wav, 24000, "hello there")
    with pytest.raises(RuntimeError, match="transcriber"):
        prompt_making.make_transcript("n", wav, 24000, None)
    prompt_making.set_language_identifier(lambda t: "en")
    assert prompt_making.make_transcript("n", wav, 24000, "hello there") == ("[EN]hello there[EN]", "en")
    prompt_making.set_transcriber(lambda w, sr: ("zh", "ni hao"))
    assert prompt_making.make_transcript("n", wav, 24000, None) == ("[ZH]ni hao.[ZH]", "zh")    # punctuation appended like :55-56
    monkeypatch.setattr(prompt_making, "language_identifier", None)
    monkeypatch.setattr(prompt_making, "transcriber", None)


def test_convert_audio_and_tokenize_audio_plumbing():
    from vallex_b200.tokenizer import AudioTokenizer, convert_audio, tokenize_audio
    stereo = torch.stack([torch.ones(1000), -torch.ones(1000)])
    mono = convert_audio(stereo, 24000, 24000, 1)
    assert tuple(mono.shape) == (1, 1000) and float(mono.abs().max()) == 0.0
    up = convert_audio(torch.zeros(1, 1200), 12000, 24000, 1)                                # resampled by torchaudio on the host
    assert up.shape[-1] == 2400
    stub = _StubCodec()
    out = tokenize_audio(stub, (stereo, 24000))
    assert tuple(stub.seen.shape) == (1, 1, 1000) and out[0][0].shape == (1, 8, 4)
    with pytest.raises(RuntimeError, match="CUDA"):                                          # no CPU path
        AudioTokenizer("cpu", state_dict={})
