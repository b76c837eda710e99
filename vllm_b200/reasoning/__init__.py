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
        """Incremental split (one parser instance per request): reasoning text streams until the end tag, tags
        never leak into either stream even when they are split across deltas."""
        st = self.__dict__.setdefault("_st", {"sent_r": 0, "done": False, "sent_c": 0})
        if st["done"]:
            return {"content": delta, "done": True} if delta else None
        t = cur
        if self.start in t:
            t = t.split(self.start, 1)[1]
        else:
            lead = t.lstrip()
            if lead and self.start.startswith(lead):
                return None                               # still inside a (possibly partial) start tag
            if not self.start_optional:
                st["done"] = True                         # no reasoning block at all: everything is content
                out = cur[st["sent_c"]:]
                st["sent_c"] = len(cur)
                return {"content": out, "done": True} if out else None
        if self.end in t:
            body, after = t.split(self.end, 1)
            st["done"] = True
            out: dict = {"done": True}
            if len(body) > st["sent_r"]:
                out["reasoning_content"] = body[st["sent_r"]:]
            if after.lstrip("\n"):
                out["content"] = after.lstrip("\n")
            return out
        safe = len(t)
        for k in range(len(self.end) - 1, 0, -1):        # hold back a partial end tag
            if t.endswith(self.end[:k]):
                safe = len(t) - k
                break
        if safe > st["sent_r"]:
            d = t[st["sent_r"]:safe]
            st["sent_r"] = safe
            return {"reasoning_content": d}
        return None


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


def _tag_parser(start_tag: str, end_tag: str, optional_start: bool = True) -> type[ReasoningParser]:
    return type(f"Tag_{abs(hash((start_tag, end_tag)))}", (ReasoningParser,),
                {"start": start_tag, "end": end_tag, "start_optional": optional_start})


class HunyuanReasoningParser(ReasoningParser):
    """``<think>\n…\n</think>\n<answer>\n…\n</answer>``"""

    def extract(self, text: str):
        r, c = super().extract(text)
        if c:
            c = c.replace("<answer>", "").replace("</answer>", "").strip() or None
        return r, c


class GptOssReasoningParser(ReasoningParser):
    """Harmony channels: ``<|channel|>analysis<|message|>…<|end|>`` is reasoning,
    ``<|channel|>final<|message|>…`` is the answer."""
    start, end = "<|channel|>analysis<|message|>", "<|end|>"
    start_optional = False

    def extract(self, text: str):
        r, rest = None, text
        if self.start in text:
            body = text.split(self.start, 1)[1]
            r, _, rest = body.partition(self.end)
        marker = "<|channel|>final<|message|>"
        c = rest.split(marker, 1)[1] if marker in rest else (rest if r is None else None)
        if c:
            c = c.replace("<|return|>", "").replace("<|end|>", "").strip() or None
        return (r.strip() or None) if r else None, c


class TrimmedThinkParser(ReasoningParser):
    """``<think>…</think>`` where the model tends to put a newline right before and right after the end tag
    (Step-3.5 style): those two newlines belong to neither stream."""

    def extract(self, text: str):
        r, c = super().extract(text)
        return (r.strip("\n") or None) if r else r, (c.lstrip("\n") or None) if c else c

    def extract_streaming(self, prev, cur, delta):
        out = super().extract_streaming(prev, cur, delta)
        if out and out.get("done") and out.get("reasoning_content"):
            out["reasoning_content"] = out["reasoning_content"].rstrip("\n") or None
            if out["reasoning_content"] is None:
                out.pop("reasoning_content")
        return out


class SelfChannelReasoningParser(ReasoningParser):
    """Channel-framed output where the chain of thought is addressed to the model itself:
    ``to=self<|message|>…reasoning…<|eom|>`` followed by the user-facing message
    (``<|start|>assistant<|message|>…<|eot|>``); all framing markers are dropped from both streams."""
    start, end = "to=self<|message|>", "<|eom|>"
    start_optional = False
    _MARKERS = ("<|start|>assistant", "<|message|>", "<|eom|>", "<|eot|>", "<|start|>")

    def _clean(self, c):
        if not c:
            return None
        for m in self._MARKERS:
            c = c.replace(m, "")
        return c.strip() or None

    def extract(self, text: str):
        if self.start not in text:
            return None, self._clean(text)
        body = text.split(self.start, 1)[1]
        r, _, rest = body.partition(self.end)
        return (r.strip() or None), self._clean(rest)


class Kimi3ReasoningParser(ReasoningParser):
    """Three-token markers: ``<|open|>think<|sep|> reasoning <|close|>think<|sep|>``; the opening marker may have been fed
    as the generation prefix, so output can begin inside the block."""
    start, end = "<|open|>think<|sep|>", "<|close|>think<|sep|>"
    start_optional = True


class IdentityReasoningParser(ReasoningParser):
    def extract(self, text: str):
        return None, text

    def extract_streaming(self, prev, cur, delta):
        return {"content": delta, "done": True}


_PARSERS = {
    "deepseek_r1": ReasoningParser, "deepseek_v3": ReasoningParser, "glm45": ReasoningParser,
    "qwen3": Qwen3ReasoningParser, "mistral": MistralReasoningParser, "granite": GraniteReasoningParser,
    "hunyuan_a13b": HunyuanReasoningParser, "step3": ReasoningParser, "kimi_k2": ReasoningParser,
    "ernie45": ReasoningParser, "minimax_m2": ReasoningParser, "olmo3": ReasoningParser, "nemotron_v3": ReasoningParser,
    "holo2": ReasoningParser, "exaone4": Qwen3ReasoningParser, "deepseek_v32": ReasoningParser,
    "seed_oss": _tag_parser("<seed:think>", "</seed:think>"), "kimi": _tag_parser("◁think▷", "◁/think▷"),
    "cohere_command": _tag_parser("<|START_THINKING|>", "<|END_THINKING|>", False),
    "gemma4": _tag_parser("<|think|>", "<|/think|>", False),
    "gpt_oss": GptOssReasoningParser, "openai_gptoss": GptOssReasoningParser, "gptoss": GptOssReasoningParser,
    "identity": IdentityReasoningParser,
    # further <think> dialects of the reference's registry
    "deepseek_v4": ReasoningParser, "glm47_moe": ReasoningParser, "hy_v3": ReasoningParser, "ling3": ReasoningParser,
    "inkling": ReasoningParser, "poolside_v1": ReasoningParser, "step3p5": TrimmedThinkParser,
    "minimax_m3": _tag_parser("<mm:think>", "</mm:think>"), "kimi_k3": Kimi3ReasoningParser,
    "muse_glimmer": SelfChannelReasoningParser,
}


def list_reasoning_parsers() -> list[str]:
    return sorted(_PARSERS)


def get_reasoning_parser(name: str) -> type[ReasoningParser]:
    if name not in _PARSERS:
        raise KeyError(f"unknown reasoning parser {name!r}; available: {sorted(_PARSERS)}")
    return _PARSERS[name]


def register_reasoning_parser(name: str, cls) -> None:
    _PARSERS[name] = cls
