"""Structured (grammar-constrained) output.

Behavioural reference: vllm/v1/structured_output/* — the grammar is compiled and
advanced in the *scheduler* process, a per-request vocabulary bitmask is shipped
with the SchedulerOutput (scheduler.py:1735) and applied to the logits in the
worker (worker/gpu/structured_outputs.py).

Back-ends: ``xgrammar`` (json schema / regex / EBNF / json_object) when it is
importable and the tokenizer is HF-compatible; a built-in ``choice``/``regex``
back-end that walks a character-level matcher over the tokenizer's vocabulary
(works with the offline ByteTokenizer) otherwise.
"""
from __future__ import annotations

import json
import re
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Any, Optional

import numpy as np

from vllm_b200.config import VllmConfig
from vllm_b200.logger import init_logger
from vllm_b200.sampling_params import StructuredOutputsParams

logger = init_logger(__name__)


class Grammar:
    """Back-end interface."""

    def fill_bitmask(self, bitmask: np.ndarray, row: int) -> None: raise NotImplementedError
    def accept_tokens(self, request_id: str, tokens: list[int]) -> bool: raise NotImplementedError
    def validate_tokens(self, tokens: list[int]) -> list[int]: raise NotImplementedError
    def rollback(self, n: int) -> None: raise NotImplementedError
    def is_terminated(self) -> bool: raise NotImplementedError


class ChoiceGrammar(Grammar):
    """Finite set of allowed strings, matched over the token vocabulary by decoded text."""

    def __init__(self, choices: list[str], tokenizer, vocab_size: int, eos: Optional[int]) -> None:
        self.choices = choices
        self.vocab_size = vocab_size
        self.eos = eos
        self.prefix = ""
        self.history: list[str] = []
        self._tok_text = _vocab_texts(tokenizer, vocab_size)

    def _allowed(self) -> list[int]:
        rem = [c[len(self.prefix):] for c in self.choices if c.startswith(self.prefix)]
        out = []
        if any(r == "" for r in rem) and self.eos is not None:
            out.append(self.eos)
        rem = [r for r in rem if r]
        if rem:
            for i, t in enumerate(self._tok_text):
                if t and any(r.startswith(t) for r in rem):
                    out.append(i)
        return out

    def fill_bitmask(self, bitmask: np.ndarray, row: int) -> None:
        bitmask[row] = 0
        for i in self._allowed():
            bitmask[row, i >> 5] |= np.int32(np.uint32(1 << (i & 31)).view(np.int32)) if (i & 31) == 31 \
                else np.int32(1 << (i & 31))

    def accept_tokens(self, request_id: str, tokens: list[int]) -> bool:
        for t in tokens:
            if t == self.eos:
                continue
            self.history.append(self.prefix)
            self.prefix += self._tok_text[t]
        return True

    def validate_tokens(self, tokens: list[int]) -> list[int]:
        ok, p = [], self.prefix
        for t in tokens:
            s = self._tok_text[t] if t < len(self._tok_text) else ""
            if not s or not any(c.startswith(p + s) for c in self.choices):
                break
            p += s
            ok.append(t)
        return ok

    def rollback(self, n: int) -> None:
        for _ in range(min(n, len(self.history))):
            self.prefix = self.history.pop()

    def is_terminated(self) -> bool:
        return self.prefix in self.choices and not any(
            c != self.prefix and c.startswith(self.prefix) for c in self.choices)


class RegexGrammar(ChoiceGrammar):
    """Regex constraint by partial matching: a token is allowed if prefix+token can still be
    extended to a full match (checked with the ``regex`` module's partial matching)."""

    def __init__(self, pattern: str, tokenizer, vocab_size: int, eos: Optional[int]) -> None:
        import regex
        self.rx = regex.compile(pattern)
        self.vocab_size, self.eos = vocab_size, eos
        self.prefix, self.history = "", []
        self._tok_text = _vocab_texts(tokenizer, vocab_size)

    def _allowed(self) -> list[int]:
        out = []
        if self.rx.fullmatch(self.prefix) and self.eos is not None:
            out.append(self.eos)
        for i, t in enumerate(self._tok_text):
            if t and self.rx.fullmatch(self.prefix + t, partial=True):
                out.append(i)
        return out

    def validate_tokens(self, tokens: list[int]) -> list[int]:
        ok, p = [], self.prefix
        for t in tokens:
            s = self._tok_text[t] if t < len(self._tok_text) else ""
            if not s or not self.rx.fullmatch(p + s, partial=True):
                break
            p += s
            ok.append(t)
        return ok

    def is_terminated(self) -> bool:
        return bool(self.rx.fullmatch(self.prefix)) and not self._allowed()[1:]


_VOCAB_CACHE: dict[int, list[str]] = {}


def _vocab_texts(tokenizer, vocab_size: int) -> list[str]:
    key = id(tokenizer)
    if key not in _VOCAB_CACHE:
        special = set(getattr(tokenizer, "all_special_ids", []) or [])
        texts = []
        for i in range(vocab_size):
            if i in special:
                texts.append("")
                continue
            try:
                texts.append(tokenizer.decode([i]))
            except Exception:
                texts.append("")
        _VOCAB_CACHE[key] = texts
    return _VOCAB_CACHE[key]


class XGrammarGrammar(Grammar):
    def __init__(self, matcher, vocab_size: int) -> None:
        self.m = matcher
        self.vocab_size = vocab_size

    def fill_bitmask(self, bitmask: np.ndarray, row: int) -> None:
        import torch
        t = torch.from_numpy(bitmask)
        self.m.fill_next_token_bitmask(t, row)

    def accept_tokens(self, request_id: str, tokens: list[int]) -> bool:
        for t in tokens:
            if not self.m.accept_token(t):
                logger.warning("grammar rejected token %d for %s", t, request_id)
                return False
        return True

    def validate_tokens(self, tokens: list[int]) -> list[int]:
        ok = []
        for t in tokens:
            if self.m.accept_token(t):
                ok.append(t)
            else:
                break
        if ok:
            self.m.rollback(len(ok))
        return ok

    def rollback(self, n: int) -> None:
        self.m.rollback(n)

    def is_terminated(self) -> bool:
        return self.m.is_terminated()


class GuidanceGrammar(Grammar):
    """llguidance matcher (reference: v1/structured_output/backend_guidance.py): JSON schema / regex / Lark grammars
    compiled lazily by the Earley-based engine, bitmask filled straight into our int32 rows."""

    def __init__(self, matcher, vocab_size: int) -> None:
        import llguidance.numpy as lln
        self.m, self.vocab_size, self._lln = matcher, vocab_size, lln
        self._mask = lln.allocate_token_bitmask(1, vocab_size)
        self.n_accepted = 0

    def fill_bitmask(self, bitmask: np.ndarray, row: int) -> None:
        self._lln.fill_next_token_bitmask(self.m, self._mask, 0)
        w = min(bitmask.shape[1], self._mask.shape[1])
        bitmask[row, :w] = self._mask[0, :w]

    def accept_tokens(self, request_id: str, tokens: list[int]) -> bool:
        for t in tokens:
            if self.m.is_stopped():
                return True
            if not self.m.consume_token(int(t)):
                return False
            self.n_accepted += 1
        return True

    def validate_tokens(self, tokens: list[int]) -> list[int]:
        n = self.m.validate_tokens([int(t) for t in tokens])
        return list(tokens[:n])

    def rollback(self, n: int) -> None:
        if n > 0:
            self.m.rollback(n)
            self.n_accepted -= n

    def is_terminated(self) -> bool:
        return bool(self.m.is_stopped())


class OutlinesGrammar(Grammar):
    """outlines-core ``Guide`` over a regex ``Index`` (reference: v1/structured_output/backend_outlines.py): JSON schemas are
    lowered to a regex by the library, the index is a token-level DFA, the allowed set is written straight into our row."""

    def __init__(self, guide, vocab_size: int) -> None:
        self.g, self.vocab_size = guide, vocab_size
        self.n_accepted = 0
        self._done = False

    def fill_bitmask(self, bitmask: np.ndarray, row: int) -> None:
        r = bitmask[row]
        r[:] = 0
        self.g.write_mask_into(r.ctypes.data, r.size, 4)

    def accept_tokens(self, request_id: str, tokens: list[int]) -> bool:
        for t in tokens:
            if self.g.is_finished() and self._done:
                return True
            if int(t) not in set(self.g.get_tokens()):
                return False
            try:
                self.g.advance(int(t))
            except Exception:
                return False
            self.n_accepted += 1
        return True

    def validate_tokens(self, tokens: list[int]) -> list[int]:
        n = 0
        for t in tokens:
            if int(t) not in set(self.g.get_tokens()):
                break
            self.g.advance(int(t))
            n += 1
        if n:
            self.g.rollback_state(n)
        return list(tokens[:n])

    def rollback(self, n: int) -> None:
        if n > 0:
            self.g.rollback_state(n)
            self.n_accepted -= n

    def is_terminated(self) -> bool:
        # finished states may still allow more text (e.g. ``a+``): terminate only when nothing but EOS can follow
        return bool(self.g.is_finished()) and len(self.g.get_tokens()) <= 1


class LMFormatEnforcerGrammar(Grammar):
    """lm-format-enforcer ``TokenEnforcer`` (reference: v1/structured_output/backend_lm_format_enforcer.py): a character-
    level parser (JSON schema / regex / choices) walked over the token prefix tree; stateless per call, so the grammar
    keeps the token history and rollback is a truncation."""

    def __init__(self, enforcer, vocab_size: int, eos: Optional[int]) -> None:
        self.e, self.vocab_size, self.eos = enforcer, vocab_size, eos
        self.hist: list[int] = []

    def _allowed(self, hist=None) -> list[int]:
        a = self.e.get_allowed_tokens(list(self.hist if hist is None else hist))
        return list(getattr(a, "allowed_tokens", a))

    def fill_bitmask(self, bitmask: np.ndarray, row: int) -> None:
        r = bitmask[row]
        r[:] = 0
        ids = np.asarray([t for t in self._allowed() if t < self.vocab_size], dtype=np.int64)
        np.bitwise_or.at(r.view(np.uint32), ids >> 5, (np.uint32(1) << (ids & 31).astype(np.uint32)))

    def accept_tokens(self, request_id: str, tokens: list[int]) -> bool:
        for t in tokens:
            if int(t) not in self._allowed():
                return False
            self.hist.append(int(t))
        return True

    def validate_tokens(self, tokens: list[int]) -> list[int]:
        h, n = list(self.hist), 0
        for t in tokens:
            if int(t) not in self._allowed(h):
                break
            h.append(int(t))
            n += 1
        return list(tokens[:n])

    def rollback(self, n: int) -> None:
        if n > 0:
            del self.hist[-n:]

    def is_terminated(self) -> bool:
        a = self._allowed()
        return bool(self.hist) and (self.hist[-1] == self.eos or (len(a) == 1 and a[0] == self.eos))


class StructuredOutputRequest:
    def __init__(self, params: StructuredOutputsParams, fut: Future) -> None:
        self.params = params
        self._fut = fut
        self.grammar: Optional[Grammar] = None

    def is_ready(self) -> bool:
        if self.grammar is not None:
            return True
        if self._fut.done():
            self.grammar = self._fut.result()
            return True
        return False

    def accept_tokens(self, request_id: str, tokens: list[int]) -> bool:
        return self.grammar.accept_tokens(request_id, tokens) if self.is_ready() else True

    def validate_tokens(self, tokens: list[int]) -> list[int]:
        return self.grammar.validate_tokens(tokens) if self.is_ready() else tokens


class StructuredOutputManager:
    def __init__(self, vllm_config: VllmConfig) -> None:
        self.vllm_config = vllm_config
        self.vocab_size = vllm_config.model_config.vocab_size
        self.backend = vllm_config.structured_outputs_config.backend
        self._tokenizer = None
        self._xgr_compiler = None
        self._pool = ThreadPoolExecutor(max_workers=2, thread_name_prefix="grammar")
        self._bitmask: Optional[np.ndarray] = None

    @property
    def tokenizer(self):
        if self._tokenizer is None:
            from vllm_b200.tokenizers import init_tokenizer_from_config
            self._tokenizer = init_tokenizer_from_config(self.vllm_config.model_config)
        return self._tokenizer

    def make_request(self, request_id: str, params: StructuredOutputsParams) -> StructuredOutputRequest:
        return StructuredOutputRequest(params, self._pool.submit(self._compile, params))

    def _compile(self, p: StructuredOutputsParams) -> Grammar:
        kind = p.kind()
        tok = self.tokenizer
        eos = self.vllm_config.model_config.eos_token_id
        if self.backend == "guidance" and hasattr(tok, "backend_tokenizer"):
            return self._compile_guidance(kind, p, tok)
        if self.backend == "outlines":
            return self._compile_outlines(kind, p, tok, eos)
        if self.backend in ("lm-format-enforcer", "lm_format_enforcer"):
            return self._compile_lmfe(kind, p, tok, eos)
        use_xgr = self.backend in ("auto", "xgrammar") and hasattr(tok, "backend_tokenizer")
        if use_xgr:
            try:
                return self._compile_xgr(kind, p, tok)
            except Exception as e:
                if self.backend == "xgrammar" or p.disable_fallback:
                    raise
                logger.warning("xgrammar failed (%s); falling back to the built-in matcher", e)
        if kind == "choice":
            return ChoiceGrammar(list(p.choice), tok, self.vocab_size, eos)
        if kind == "regex":
            return RegexGrammar(p.regex, tok, self.vocab_size, eos)
        if kind in ("json", "json_object"):
            from vllm_b200.structured_output.json_regex import json_schema_to_regex
            schema = p.json if kind == "json" else {"type": "object"}
            if isinstance(schema, str):
                schema = json.loads(schema)
            return RegexGrammar(json_schema_to_regex(schema), tok, self.vocab_size, eos)
        raise ValueError(f"structured output kind {kind!r} needs the xgrammar backend")

    def _compile_xgr(self, kind: str, p: StructuredOutputsParams, tok) -> Grammar:
        import xgrammar as xgr
        if self._xgr_compiler is None:
            info = xgr.TokenizerInfo.from_huggingface(tok, vocab_size=self.vocab_size)
            self._xgr_compiler = xgr.GrammarCompiler(info, max_threads=2)
        c = self._xgr_compiler
        if kind == "json":
            s = p.json if isinstance(p.json, str) else json.dumps(p.json)
            g = c.compile_json_schema(s, any_whitespace=True)
        elif kind == "json_object":
            g = c.compile_json_schema('{"type": "object"}')
        elif kind == "regex":
            g = c.compile_regex(p.regex)
        elif kind == "grammar":
            g = c.compile_grammar(p.grammar)
        elif kind == "choice":
            alts = " | ".join(json.dumps(ch) for ch in p.choice)
            g = c.compile_grammar(f"root ::= {alts}")
        else:
            raise ValueError(kind)
        return XGrammarGrammar(xgr.GrammarMatcher(g, max_rollback_tokens=64), self.vocab_size)

    def _regex_for(self, kind: str, p: StructuredOutputsParams) -> str:
        if kind == "regex":
            return p.regex
        if kind == "choice":
            return "(" + "|".join(re.escape(c) for c in p.choice) + ")"
        if kind in ("json", "json_object"):
            import outlines_core.json_schema as ojs
            schema = p.json if kind == "json" else {"type": "object"}
            return ojs.build_regex_from_schema(schema if isinstance(schema, str) else json.dumps(schema))
        raise ValueError(f"the outlines back-end compiles regex, JSON-schema and choice constraints, not {kind!r}")

    def _compile_outlines(self, kind: str, p: StructuredOutputsParams, tok, eos) -> Grammar:
        from outlines_core import Guide, Index, Vocabulary
        if getattr(self, "_outlines_vocab", None) is None:
            by_text: dict[str, list[int]] = {}
            for i, t in enumerate(_vocab_texts(tok, self.vocab_size)):
                if t and i != eos:
                    by_text.setdefault(t, []).append(i)
            self._outlines_vocab = Vocabulary(int(eos if eos is not None else 0), by_text)
        return OutlinesGrammar(Guide(Index(self._regex_for(kind, p), self._outlines_vocab)), self.vocab_size)

    def _compile_lmfe(self, kind: str, p: StructuredOutputsParams, tok, eos) -> Grammar:
        import lmformatenforcer as lmfe
        if getattr(self, "_lmfe_data", None) is None:
            texts = _vocab_texts(tok, self.vocab_size)
            regular = [(i, t, t.startswith(" ")) for i, t in enumerate(texts) if t and i != eos]
            self._lmfe_data = lmfe.TokenEnforcerTokenizerData(regular, lambda ids: tok.decode(list(ids)), int(eos if eos is not None else 0),
                                                               False, self.vocab_size)
        if kind == "json":
            parser = lmfe.JsonSchemaParser(json.loads(p.json) if isinstance(p.json, str) else p.json)
        elif kind == "json_object":
            parser = lmfe.JsonSchemaParser(None)
        elif kind == "regex":
            parser = lmfe.RegexParser(p.regex)
        elif kind == "choice":
            parser = lmfe.UnionParser([lmfe.StringParser(c) for c in p.choice])
        else:
            raise ValueError(f"lm-format-enforcer compiles JSON-schema, regex and choice constraints, not {kind!r}")
        return LMFormatEnforcerGrammar(lmfe.TokenEnforcer(self._lmfe_data, parser), self.vocab_size, eos)

    def _compile_guidance(self, kind: str, p: StructuredOutputsParams, tok) -> Grammar:
        import llguidance
        import llguidance.hf
        if getattr(self, "_llg_tok", None) is None:
            self._llg_tok = llguidance.hf.from_tokenizer(tok, n_vocab=self.vocab_size)
        M = llguidance.LLMatcher
        if kind == "json":
            g = M.grammar_from_json_schema(p.json if isinstance(p.json, str) else json.dumps(p.json))
        elif kind == "json_object":
            g = M.grammar_from_json_schema('{"type": "object"}')
        elif kind == "regex":
            g = M.grammar_from_regex(p.regex)
        elif kind == "grammar":
            g = M.grammar_from_lark(p.grammar)
        elif kind == "choice":
            g = M.grammar_from_regex("(" + "|".join(re.escape(c) for c in p.choice) + ")")
        else:
            raise ValueError(kind)
        m = M(self._llg_tok, g)
        if m.is_error():
            raise ValueError(f"guidance grammar error: {m.get_error()}")
        return GuidanceGrammar(m, self.vocab_size)

    def grammar_bitmask(self, requests: dict, structured_ids: list[str], spec_tokens: dict[str, list[int]]
                        ) -> Optional[np.ndarray]:
        """One bitmask row per verified position: request ``rid`` with k scheduled draft tokens owns k + 1 consecutive
        rows (the mask after drafts[:j] for j = 0..k), in ``structured_ids`` order — the grammar is advanced over the
        drafts and rolled back (reference: structured_output/__init__.py ``grammar_bitmask`` with
        ``scheduled_spec_decode_tokens``).  Without speculation that is the familiar one row per request."""
        if not structured_ids:
            return None
        W = (self.vocab_size + 31) // 32
        n = sum(len(spec_tokens.get(rid) or ()) + 1 for rid in structured_ids)
        if self._bitmask is None or self._bitmask.shape[0] < n:
            self._bitmask = np.full((max(n, 8), W), -1, dtype=np.int32)
        bm = self._bitmask
        eos = self.vllm_config.model_config.eos_token_id

        def fill(g, row):
            if g.is_terminated():
                # nothing may follow a complete match: only end-of-sequence is allowed (the scheduler also finishes
                # the request as soon as its grammar terminates)
                if eos is None:
                    bm[row] = -1
                else:
                    bm[row] = 0
                    bm[row, eos >> 5] = np.int32(np.uint32(1 << (eos & 31)).view(np.int32))
            else:
                g.fill_bitmask(bm, row)
        row = 0
        for rid in structured_ids:
            req = requests.get(rid)
            so = req.structured_output_request if req is not None else None
            drafts = list(spec_tokens.get(rid) or ())
            if so is None or not so.is_ready():
                bm[row:row + len(drafts) + 1] = -1
                row += len(drafts) + 1
                continue
            g = so.grammar
            fill(g, row)
            row += 1
            adv = 0
            for d in drafts:
                if adv >= 0 and not g.is_terminated() and g.accept_tokens(rid, [d]):
                    adv += 1
                    fill(g, row)
                else:
                    # an inadmissible draft is rejected at its own position; later positions are never reached
                    if adv >= 0:
                        g.rollback(adv)
                        adv = -1
                    bm[row] = -1
                row += 1
            if adv > 0:
                g.rollback(adv)
        return bm[:n].copy()

    def shutdown(self) -> None:
        self._pool.shutdown(wait=False)
