"""Reasoning-content parsers: split chain-of-thought from the answer (reference: vllm/reasoning/*, ≈25 formats).

``deepseek_r1`` / ``qwen3`` / ``glm45`` style ``<think>…</think>`` (start tag optional, as R1 templates
pre-fill it), ``granite`` ("Here is my thought process:" / "Here is my response:"), ``mistral`` ([THINK]…[/THINK]).
"""
from __future__ import annotations

from typing import Optional


class ReasoningParser:
    start, end = "<think>", "</think>"
    start_optional = True

    def __init__(self, tokenizer=None) -> None:
        self.tokenizer = tokenizer

    def extract(self, text: str) -> tuple[Optional[str], Optional[str]]:
        """→ (reasoning_content, content)"""
        t = text
        if self.start in t:
            t = t.split(self.start, 1)[1]
        elif not self.start_optional:
            return None, text
        if self.end not in t:
            # never closed: everything so far is reasoning (R1 convention)
            return (t or None), None if self.start_optional or self.start in text else text
        r, c = t.split(self.end, 1)
        return (r.strip() or None), (c.lstrip("\n") or None)

    def extract_streaming(self, prev: str, cur: str, delta: str) -> Optional[dict]:
        if self.end in prev:
            return {"content": delta, "done": True}
        if self.end in cur:
            before, after = cur.split(self.end, 1)
            sent = len(prev.split(self.start, 1)[-1]) if self.start in prev else len(prev)
            body = before.split(self.start, 1)[-1]
            out = {"done": True}
            if len(body) > sent:
                out["reasoning_content"] = body[sent:]
            if after:
                out["content"] = after.lstrip("\n")
            return out
        d = delta.replace(self.start, "")
        # hold back a partial end tag
        for k in range(len(self.end) - 1, 0, -1):
            if cur.endswith(self.end[:k]):
                d = d[:-k] if len(d) >= k else ""
                break
        return {"reasoning_content": d} if d else None


class Qwen3ReasoningParser(ReasoningParser):
    start_optional = False

    def extract(self, text: str):
        if self.start not in text and self.end not in text:
            return None, text
        return super().extract(text if self.start in text else self.start + text)


class MistralReasoningParser(ReasoningParser):
    start, end = "[THINK]", "[/THINK]"
    start_optional = False


class GraniteReasoningParser(ReasoningParser):
    start, end = "Here is my thought process:", "Here is my response:"
    start_optional = False


_PARSERS = {
    "deepseek_r1": ReasoningParser, "deepseek_v3": ReasoningParser, "glm45": ReasoningParser,
    "qwen3": Qwen3ReasoningParser, "mistral": MistralReasoningParser, "granite": GraniteReasoningParser,
    "hunyuan_a13b": ReasoningParser, "step3": ReasoningParser, "kimi_k2": ReasoningParser,
}


def get_reasoning_parser(name: str) -> type[ReasoningParser]:
    if name not in _PARSERS:
        raise KeyError(f"unknown reasoning parser {name!r}; available: {sorted(_PARSERS)}")
    return _PARSERS[name]


def register_reasoning_parser(name: str, cls) -> None:
    _PARSERS[name] = cls
