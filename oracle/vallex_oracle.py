"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product.

A plain-PyTorch (CPU, fp32) restatement of the reference's AR + NAR algorithm for
the `VALLE.inference` hot path.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module; the
product path (`vall-e-x_b200/`) never does and fails loudly without its CUDA library.

Pinned (not "parity unpinned"): `tests/golden/make_golden.py` runs the UNMODIFIED reference
`models.vallex.VALLE` from /root/reference (with the matplotlib stub described in SURVEY.md
8(c)) on the same synthetic checkpoints and commits its outputs to `tests/golden/*.npz`;
`tests/test_oracle_golden.py` requires this restatement to reproduce every committed token
bit-exactly, and (when /root/reference is present) checks it live against the reference.

Each function cites the reference file:line it restates.  The structure is deliberately
different from the reference (functional, explicit KV cache with a single growing buffer,
only the last y row re-embedded per step) -- the arithmetic per element is the same.
"""
from __future__ import annotations

import math
import time
from typing import Dict, List, Optional, Sequence, Union

import torch
import torch.nn.functional as F

D = 1024
H = 16
DH = 64
NL = 12
NQ = 8
EOS = 1024
BOS = 1025
LANGUAGE_ID = {"en": 0, "zh": 1, "ja": 2}  # models/vallex.py:439-443


def sine_table(length: int, dim: int = D) -> torch.Tensor:
    """modules/embedding.py:68-91 (non-reversed branch)."""
    position = torch.arange(0, length, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * -(math.log(10000.0) / dim))
    pe = torch.zeros(length, dim)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def add_position(x: torch.Tensor, alpha: torch.Tensor, pe: torch.Tensor, start: int = 0) -> torch.Tensor:
    """modules/embedding.py:93-97: out = x * 1.0 + alpha * pe[start:start+len] (x_scale = 1)."""
    n = x.shape[-2]
    return x * 1.0 + alpha * pe[start:start + n]


def top_k_top_p_filtering(logits, top_k=0, top_p=1.0, filter_value=-float("inf"), min_tokens_to_keep=1):
    """models/vallex.py:791-833."""
    if top_k > 0:
        k = min(max(top_k, min_tokens_to_keep), logits.size(-1))
        kth = torch.topk(logits, k)[0][..., -1, None]
        logits = logits.masked_fill(logits < kth, filter_value)
    if top_p < 1.0:
        sorted_logits, sorted_idx = torch.sort(logits, descending=True)
        cum = torch.cumsum(F.softmax(sorted_logits, dim=-1), dim=-1)
        remove = cum > top_p
        if min_tokens_to_keep > 1:
            remove[..., :min_tokens_to_keep] = 0
        remove[..., 1:] = remove[..., :-1].clone()
        remove[..., 0] = 0
        logits = logits.masked_fill(remove.scatter(1, sorted_idx, remove), filter_value)
    return logits


def topk_sampling(logits, top_k=10, top_p=1.0, temperature=1.0, generator=None):
    """models/vallex.py:836-853."""
    if temperature != 1.0:
        logits = logits / temperature
    logits = top_k_top_p_filtering(logits, top_k=top_k, top_p=top_p)
    token = torch.multinomial(F.softmax(logits, dim=-1), num_samples=1, generator=generator)
    logprobs = F.log_softmax(logits.float(), dim=-1)
    return token, logprobs[torch.arange(logprobs.shape[0]), token.squeeze(1)]


class OracleVALLE:
    """Functional restatement of `VALLE.inference` (models/vallex.py:458-686)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor]):
        """Pass `vallex_b200.weights.round_state_dict_bf16(sd)` to get the oracle the bf16 engine mode is
        compared against (same weight VALUES, fp32 arithmetic; SURVEY.md 7, hard part 1c)."""
        self.sd = {k: v.detach().to(torch.float32) for k, v in state_dict.items()}
        self.pe = sine_table(4000)

    def _pe(self, n: int) -> torch.Tensor:
        if self.pe.shape[0] < n:  # modules/embedding.py:68-72 auto-extension
            self.pe = sine_table(n)
        return self.pe

    # ---- shared blocks -------------------------------------------------------------------
    def _text(self, which: str, text: torch.Tensor, Lp: int, prompt_language: str,
              text_language: Union[str, Sequence[str]]) -> torch.Tensor:
        """models/vallex.py:497-507 (AR) and :622-632 (NAR)."""
        sd = self.sd
        x = F.embedding(text, sd[f"{which}_text_embedding.word_embeddings.weight"])
        lang = sd[f"{which}_language_embedding.word_embeddings.weight"]
        pl = torch.tensor([LANGUAGE_ID[prompt_language]])
        if isinstance(text_language, str):
            tl = torch.tensor([LANGUAGE_ID[text_language]])
        else:
            tl = torch.tensor([LANGUAGE_ID[t] for t in text_language])
        x = x.clone()
        x[:, :Lp, :] += F.embedding(pl, lang)
        x[:, Lp:, :] += F.embedding(tl, lang)
        return add_position(x, sd[f"{which}_text_position.alpha"], self._pe(x.shape[1]))

    def _ar_layer(self, i: int, x: torch.Tensor, kv, mask: torch.Tensor):
        """modules/transformer.py:337-347 + modules/activation.py:142-167 (pre-LN, KV cache)."""
        sd, p = self.sd, f"ar_decoder.layers.{i}"
        B, T, C = x.shape
        h = F.layer_norm(x, (D,), sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], 1e-5)
        q, k, v = F.linear(h, sd[f"{p}.self_attn.in_proj_weight"], sd[f"{p}.self_attn.in_proj_bias"]).chunk(3, dim=-1)
        q = q.view(B, T, H, DH).transpose(1, 2)
        k = k.view(B, T, H, DH).transpose(1, 2)
        v = v.view(B, T, H, DH).transpose(1, 2)
        if kv is not None:
            k = torch.cat((kv[0], k), dim=-2)
            v = torch.cat((kv[1], v), dim=-2)
        full = k.shape[-2]
        att = (q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(DH))
        att = att.masked_fill(mask[full - T:full, :full], float("-inf"))
        att = F.softmax(att, dim=-1)
        y = (att @ v).transpose(1, 2).contiguous().view(B, T, C)
        y = F.linear(y, sd[f"{p}.self_attn.out_proj.weight"], sd[f"{p}.self_attn.out_proj.bias"])
        x = x + y
        h = F.layer_norm(x, (D,), sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], 1e-5)
        h = F.linear(F.relu(F.linear(h, sd[f"{p}.linear1.weight"], sd[f"{p}.linear1.bias"])),
                     sd[f"{p}.linear2.weight"], sd[f"{p}.linear2.bias"])
        return x + h, (k, v)

    def _ar_stack(self, x, kvs, mask):
        new = []
        for i in range(NL):
            x, kv = self._ar_layer(i, x, None if kvs is None else kvs[i], mask)
            new.append(kv)
        x = F.layer_norm(x, (D,), self.sd["ar_decoder.norm.weight"], self.sd["ar_decoder.norm.bias"], 1e-5)
        return x, new

    @staticmethod
    def ar_mask(S: int, Ly: int) -> torch.Tensor:
        """models/vallex.py:534-549: text rows see text only; audio rows see all text + causal audio."""
        m = torch.zeros(S + Ly, S + Ly, dtype=torch.bool)
        m[:S, S:] = True
        m[S:, S:] = torch.triu(torch.ones(Ly, Ly, dtype=torch.bool), diagonal=1)
        return m

    # ---- AR ----------------------------------------------------------------------------
    @torch.no_grad()
    def ar_decode(self, x_ids, y_prompt0, Lp, prompt_language, text_language, top_k=-100,
                  temperature=1.0, best_of=1, length_penalty=1.0, return_worst=False,
                  max_new_tokens: Optional[int] = None, generator=None, debug: Optional[dict] = None):
        """models/vallex.py:496-598.  x_ids [1,S] int64, y_prompt0 [1,P] int64 (codebook 0).
        Returns y [1, 1+P+T] (BOS + prompt + generated, EOS stripped).
        `max_new_tokens` (engine extension, not in the reference) truncates the run: greedy tokens
        are a prefix of the untruncated run because the decoder is causal."""
        sd = self.sd
        S = x_ids.shape[1]
        P = y_prompt0.shape[1]
        x = self._text("ar", x_ids, Lp, prompt_language, text_language)
        y = F.pad(y_prompt0, (1, 0), value=BOS)
        x = x.repeat(best_of, 1, 1)
        y = y.repeat(best_of, 1)
        sum_logprobs = torch.zeros(best_of)
        cap = 16 * S
        aemb = sd["ar_audio_embedding.word_embeddings.weight"]
        alpha = sd["ar_audio_position.alpha"]
        kvs = None
        while True:
            if debug is not None and "times" in debug:   # bench.py's CPU baseline: wall clock at the start of every iteration
                debug["times"].append(time.perf_counter())
            Ly = y.shape[1]
            mask = self.ar_mask(S, Ly)
            if kvs is None:
                y_pos = add_position(F.embedding(y, aemb), alpha, self._pe(Ly))
                xy = torch.cat([x, y_pos], dim=1)
            else:  # only the newest row is consumed (models/vallex.py:552-553)
                xy = add_position(F.embedding(y[:, -1:], aemb), alpha, self._pe(Ly), start=Ly - 1)
            dec, kvs = self._ar_stack(xy, kvs, mask)
            logits = F.linear(dec[:, -1], sd["ar_predict_layer.weight"])
            if debug is not None:
                debug.setdefault("logits", []).append(logits.clone())
            samples, lp = topk_sampling(logits, top_k=top_k, top_p=1.0, temperature=temperature, generator=generator)
            sum_logprobs += lp * (y[:, -1] != EOS)
            samples[y[:, -1] == EOS] = EOS
            completed = bool((samples[:, -1] == EOS).all())
            n_new = Ly - 1 - P
            hit_cap = (Ly - P) > cap
            if max_new_tokens is not None and n_new >= max_new_tokens:
                hit_cap = True
            if completed or hit_cap:
                if P == Ly:
                    raise SyntaxError("well trained model shouldn't reach here.")
                lengths = torch.sum(y != EOS, dim=1)
                avg = sum_logprobs / lengths ** length_penalty
                best = y[torch.argmax(avg), :]
                worst = y[torch.argmin(avg), :]
                best = best[best != EOS]
                worst = worst[worst != EOS]
                y = (worst if return_worst else best).unsqueeze(0)
                break
            y = torch.cat([y, samples], dim=1)
        return y

    # ---- NAR ---------------------------------------------------------------------------
    def _adaln(self, prefix: str, x: torch.Tensor, stage_emb: torch.Tensor) -> torch.Tensor:
        """modules/transformer.py:93-108."""
        sd = self.sd
        w, b = torch.split(F.linear(stage_emb, sd[f"{prefix}.project_layer.weight"], sd[f"{prefix}.project_layer.bias"]),
                           D, dim=-1)
        return w * F.layer_norm(x, (D,), sd[f"{prefix}.norm.weight"], sd[f"{prefix}.norm.bias"], 1e-5) + b

    def _nar_layer(self, i: int, x: torch.Tensor, stage_emb: torch.Tensor) -> torch.Tensor:
        """modules/transformer.py:296-302 + modules/activation.py:566-585 (bidirectional MHA, no mask)."""
        sd, p = self.sd, f"nar_decoder.layers.{i}"
        B, L, _ = x.shape
        h = self._adaln(f"{p}.norm1", x, stage_emb)
        qkv = F.linear(h, sd[f"{p}.self_attn.in_proj_weight"], sd[f"{p}.self_attn.in_proj_bias"])
        q, k, v = qkv.view(B, L, 3, H, DH).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v)
        a = a.transpose(1, 2).reshape(B, L, D)
        x = x + F.linear(a, sd[f"{p}.self_attn.out_proj.weight"], sd[f"{p}.self_attn.out_proj.bias"])
        h = self._adaln(f"{p}.norm2", x, stage_emb)
        h = F.linear(F.relu(F.linear(h, sd[f"{p}.linear1.weight"], sd[f"{p}.linear1.bias"])),
                     sd[f"{p}.linear2.weight"], sd[f"{p}.linear2.bias"])
        return x + h

    @torch.no_grad()
    def nar_decode(self, x_ids, y_full0, prompts, Lp, prompt_language, text_language,
                   debug: Optional[dict] = None) -> torch.Tensor:
        """models/vallex.py:604-686, prefix_mode == 1 branch.  y_full0 [1,P+T] = prompt cb0 + generated cb0;
        prompts [1,P,8].  Returns codes [1,T,8]."""
        sd = self.sd
        S = x_ids.shape[1]
        P = prompts.shape[1]
        codes = [y_full0[:, P:]]
        y_emb = F.embedding(y_full0, sd["nar_audio_embeddings.0.word_embeddings.weight"]).clone()
        x = self._text("nar", x_ids, Lp, prompt_language, text_language)
        for j in range(1, NQ):
            y_emb[:, :P] += F.embedding(prompts[..., j], sd[f"nar_audio_embeddings.{j}.word_embeddings.weight"])
        for i in range(NQ - 1):
            y_pos = add_position(y_emb, sd["nar_audio_position.alpha"], self._pe(y_emb.shape[1]))
            xy = torch.cat([x, y_pos], dim=1)
            stage = sd[f"nar_stage_embeddings.{i}.word_embeddings.weight"]
            for l in range(NL):
                xy = self._nar_layer(l, xy, stage)
            xy = self._adaln("nar_decoder.norm", xy, stage)
            logits = F.linear(xy[:, S + P:], sd[f"nar_predict_layers.{i}.weight"])
            if debug is not None:
                debug.setdefault("nar_logits", []).append(logits.clone())
            samples = torch.argmax(logits, dim=-1)
            codes.append(samples)
            if i < NQ - 2:
                y_emb[:, P:] += F.embedding(samples, sd[f"nar_audio_embeddings.{i + 1}.word_embeddings.weight"])
        return torch.stack(codes, dim=-1)

    # ---- whole path ----------------------------------------------------------------------
    @torch.no_grad()
    def inference(self, x, x_lens, y, enroll_x_lens, top_k=-100, temperature=1.0, prompt_language=None,
                  text_language=None, best_of=1, length_penalty=1.0, return_worst=False,
                  max_new_tokens: Optional[int] = None, generator=None, debug: Optional[dict] = None):
        """Same signature/semantics as models/vallex.py:458-686; returns int64 [1,T,8]."""
        assert x.ndim == 2 and x_lens.ndim == 1 and y.ndim == 3 and y.shape[0] == 1
        x = x.long()
        y = y.long()
        Lp = int(enroll_x_lens)
        P = y.shape[1]
        yy = self.ar_decode(x, y[..., 0], Lp, prompt_language, text_language, top_k, temperature, best_of,
                            length_penalty, return_worst, max_new_tokens, generator, debug)
        return self.nar_decode(x, yy[:, 1:], y, Lp, prompt_language, text_language, debug)
