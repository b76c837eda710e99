"""Tokenizer layer (reference: vllm/tokenizers/* — ``TokenizerLike`` protocol with
HF / mistral / custom registries).

Offline boxes have no tokenizer files, so besides the HuggingFace wrapper there
is a dependency-free ``ByteTokenizer`` (UTF-8 bytes + specials) that any model
preset can use; it is also what the tests use.
"""
from __future__ import annotations

import os
from typing import Any, Optional, Protocol, Sequence


class TokenizerLike(Protocol):
    eos_token_id: Optional[int]
    bos_token_id: Optional[int]
    vocab_size: int

    def encode(self, text: str, add_special_tokens: bool = True) -> list[int]: ...
    def decode(self, ids: Sequence[int], skip_special_tokens: bool = True) -> str: ...
    def convert_ids_to_tokens(self, ids: Sequence[int]) -> list[str]: ...
    def apply_chat_template(self, messages: list[dict], **kw) -> Any: ...


class ByteTokenizer:
    """ids 0..2 = <pad>/<bos>/<eos>; id 3+b = byte b. Ids ≥ 259 decode to U+FFFD-free
    placeholders so random-weight models with big vocabularies still detokenise."""

    PAD, BOS, EOS, OFFSET = 0, 1, 2, 3

    def __init__(self, vocab_size: int = 512, eos_token_id: Optional[int] = None,
                 bos_token_id: Optional[int] = None) -> None:
        self.vocab_size = vocab_size
        self.eos_token_id = self.EOS if eos_token_id is None else eos_token_id
        self.bos_token_id = self.BOS if bos_token_id is None else bos_token_id
        self.pad_token_id = self.PAD
        self.all_special_ids = {self.PAD, self.bos_token_id, self.eos_token_id}
        self.chat_template = None
        self.is_fast = False

    def encode(self, text: str, add_special_tokens: bool = True, **kw) -> list[int]:
        ids = [b + self.OFFSET for b in text.encode("utf-8")]
        return ([self.bos_token_id] + ids) if add_special_tokens else ids

    def __call__(self, text, add_special_tokens: bool = True, **kw):
        if isinstance(text, str):
            return {"input_ids": self.encode(text, add_special_tokens)}
        return {"input_ids": [self.encode(t, add_special_tokens) for t in text]}

    def _tok(self, i: int) -> bytes:
        if self.OFFSET <= i < self.OFFSET + 256:
            return bytes([i - self.OFFSET])
        return f"<{i}>".encode()

    def decode(self, ids: Sequence[int], skip_special_tokens: bool = True, **kw) -> str:
        out = bytearray()
        for i in ids:
            if skip_special_tokens and i in self.all_special_ids:
                continue
            out += self._tok(int(i))
        return out.decode("utf-8", errors="replace")

    def convert_ids_to_tokens(self, ids: Sequence[int], skip_special_tokens: bool = False) -> list[str]:
        return [self._tok(int(i)).decode("latin-1") for i in ids
                if not (skip_special_tokens and i in self.all_special_ids)]

    def convert_tokens_to_string(self, tokens: list[str]) -> str:
        return "".join(tokens).encode("latin-1").decode("utf-8", errors="replace")

    def get_vocab(self) -> dict[str, int]:
        return {self._tok(i).decode("latin-1"): i for i in range(min(self.vocab_size, 259))}

    def apply_chat_template(self, messages: list[dict], add_generation_prompt: bool = True,
                            tokenize: bool = False, **kw):
        text = "".join(f"<|{m['role']}|>\n{_content_text(m.get('content'))}\n" for m in messages)
        if add_generation_prompt:
            text += "<|assistant|>\n"
        return self.encode(text) if tokenize else text

    def __len__(self) -> int:
        return self.vocab_size


def _content_text(c) -> str:
    if c is None:
        return ""
    if isinstance(c, str):
        return c
    return "".join(p.get("text", "") for p in c if isinstance(p, dict))


def get_tokenizer(name: Optional[str], tokenizer_mode: str = "auto", trust_remote_code: bool = False,
                  vocab_size: int = 512, eos_token_id: Optional[int] = None, **kw):
    if tokenizer_mode == "byte" or name is None or not os.path.isdir(name):
        return ByteTokenizer(vocab_size, eos_token_id)
    import glob
    has_hf = any(os.path.exists(os.path.join(name, f)) for f in ("tokenizer.json", "tokenizer.model", "vocab.json"))
    if tokenizer_mode == "tiktoken" or (tokenizer_mode == "auto" and not has_hf
                                        and (glob.glob(os.path.join(name, "*.tiktoken")) or os.path.exists(os.path.join(name, "tiktoken.model")))):
        from vllm_b200.tokenizers.tiktoken_tok import TiktokenTokenizer
        return TiktokenTokenizer.from_dir(name)
    if tokenizer_mode == "mistral":
        from vllm_b200.tokenizers.mistral_tok import MistralTokenizer
        return MistralTokenizer.from_dir(name)
    from transformers import AutoTokenizer
    return AutoTokenizer.from_pretrained(name, trust_remote_code=trust_remote_code, **kw)


def init_tokenizer_from_config(model_config):
    if model_config.skip_tokenizer_init:
        return None
    tok = get_tokenizer(model_config.tokenizer, model_config.tokenizer_mode, model_config.trust_remote_code,
                        vocab_size=model_config.vocab_size, eos_token_id=model_config.eos_token_id)
    try:        # lets chat rendering fall back to the model family's built-in template (entrypoints/chat_templates.py)
        tok.model_type = model_config.hf_config.get("model_type")
    except Exception:
        pass
    return tok
