"""Prompt input types (reference: vllm/inputs/)."""
from __future__ import annotations

from typing import Any, Optional, TypedDict, Union


class TextPrompt(TypedDict, total=False):
    prompt: str
    multi_modal_data: dict[str, Any]
    cache_salt: str


class TokensPrompt(TypedDict, total=False):
    prompt_token_ids: list[int]
    prompt: str
    multi_modal_data: dict[str, Any]
    cache_salt: str


class EmbedsPrompt(TypedDict, total=False):
    prompt_embeds: Any
    cache_salt: str


PromptType = Union[str, list[int], TextPrompt, TokensPrompt, EmbedsPrompt]


def parse_prompt(p: PromptType) -> tuple[Optional[str], Optional[list[int]], dict[str, Any]]:
    """→ (text, token_ids, extras)"""
    if isinstance(p, str):
        return p, None, {}
    if isinstance(p, list):
        return None, list(p), {}
    if isinstance(p, dict):
        extras = {k: v for k, v in p.items() if k not in ("prompt", "prompt_token_ids")}
        if "prompt_token_ids" in p:
            return p.get("prompt"), list(p["prompt_token_ids"]), extras
        if "prompt" in p:
            return p["prompt"], None, extras
        if any(k in p for k in ("prompt_embeds", "encoder_prompt_token_ids", "encoder_prompt", "multi_modal_data")):
            return None, None, extras
    raise TypeError(f"unsupported prompt type {type(p)}")
