"""tiktoken-backed tokenizer (``tokenizer_mode="tiktoken"`` or a model directory that ships ``*.tiktoken`` /
``tiktoken.model``): rank files of Qwen-1 / Kimi / Llama-3-original style checkpoints.

Reference: vllm/tokenizers/ (registry of HF / mistral / tiktoken-style back-ends behind ``TokenizerLike``).
"""
from __future__ import annotations

import base64
import glob
import json
import os
from typing import Optional, Sequence

# the GPT-4 / Llama-3 split pattern (what rank files of this family are trained with unless the config says otherwise)
DEFAULT_PATTERN = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|"
                   r"\s+(?!\S)|\s+")


def load_ranks(path: str) -> dict[bytes, int]:
    ranks = {}
    with open(path, "rb") as f:
        for line in f:
            if line.strip():
                tok, rank = line.split()
                ranks[base64.b64decode(tok)] = int(rank)
    return ranks


class TiktokenTokenizer:
    is_fast = True

    def __init__(self, ranks: dict[bytes, int], special_tokens: Optional[dict[str, int]] = None,
                 pattern: str = DEFAULT_PATTERN, bos_token: Optional[str] = None, eos_token: Optional[str] = None,
                 chat_template: Optional[str] = None) -> None:
        import tiktoken
        self.special = dict(special_tokens or {})
        self.enc = tiktoken.Encoding("vllm_b200", pat_str=pattern, mergeable_ranks=ranks, special_tokens=self.special)
        self.vocab_size = self.enc.n_vocab
        self.bos_token, self.eos_token = bos_token, eos_token
        self.bos_token_id = self.special.get(bos_token) if bos_token else None
        self.eos_token_id = self.special.get(eos_token) if eos_token else None
        self.pad_token_id = None
        self.all_special_ids = set(self.special.values())
        self.chat_template = chat_template
        self._inv_special = {v: k for k, v in self.special.items()}

    @classmethod
    def from_dir(cls, path: str) -> "TiktokenTokenizer":
        files = glob.glob(os.path.join(path, "*.tiktoken")) + glob.glob(os.path.join(path, "tiktoken.model"))
        if not files:
            raise FileNotFoundError(f"no *.tiktoken / tiktoken.model file in {path}")
        ranks = load_ranks(files[0])
        cfg = {}
        tc = os.path.join(path, "tokenizer_config.json")
        if os.path.exists(tc):
            with open(tc) as f:
                cfg = json.load(f)
        special = {}
        for tid, d in (cfg.get("added_tokens_decoder") or {}).items():
            special[d["content"] if isinstance(d, dict) else str(d)] = int(tid)
        n = len(ranks)
        for i, t in enumerate(cfg.get("additional_special_tokens") or []):
            special.setdefault(t, n + len(special) if t not in special else special[t])
        def name(x):
            return x.get("content") if isinstance(x, dict) else x
        return cls(ranks, special, cfg.get("pat_str") or DEFAULT_PATTERN, name(cfg.get("bos_token")), name(cfg.get("eos_token")),
                   cfg.get("chat_template"))

    # ---- TokenizerLike
    def encode(self, text: str, add_special_tokens: bool = True, **kw) -> list[int]:
        ids = self.enc.encode(text, allowed_special="all")
        if add_special_tokens and self.bos_token_id is not None:
            ids = [self.bos_token_id] + ids
        return ids

    def __call__(self, text, add_special_tokens: bool = True, **kw):
        if isinstance(text, str):
            return {"input_ids": self.encode(text, add_special_tokens)}
        return {"input_ids": [self.encode(t, add_special_tokens) for t in text]}

    def decode(self, ids: Sequence[int], skip_special_tokens: bool = True, **kw) -> str:
        ids = [int(i) for i in ids if not (skip_special_tokens and int(i) in self.all_special_ids)]
        return self.enc.decode(ids, errors="replace")

    def convert_ids_to_tokens(self, ids: Sequence[int], skip_special_tokens: bool = False) -> list[str]:
        out = []
        for i in ids:
            i = int(i)
            if i in self._inv_special:
                if not skip_special_tokens:
                    out.append(self._inv_special[i])
            else:
                out.append(self.enc.decode_single_token_bytes(i).decode("latin-1"))
        return out

    def convert_tokens_to_string(self, tokens: list[str]) -> str:
        return "".join(tokens).encode("latin-1", errors="replace").decode("utf-8", errors="replace")

    def convert_tokens_to_ids(self, token):
        if isinstance(token, (list, tuple)):
            return [self.convert_tokens_to_ids(t) for t in token]
        if token in self.special:
            return self.special[token]
        return self.enc._mergeable_ranks.get(token.encode("latin-1") if isinstance(token, str) else token)

    def get_vocab(self) -> dict[str, int]:
        v = {k.decode("latin-1"): r for k, r in self.enc._mergeable_ranks.items()}
        v.update(self.special)
        return v

    def apply_chat_template(self, messages: list[dict], add_generation_prompt: bool = True, tokenize: bool = False,
                            chat_template: Optional[str] = None, **kw):
        tpl = chat_template or self.chat_template
        if tpl:
            import jinja2
            text = jinja2.Environment().from_string(tpl).render(messages=messages, add_generation_prompt=add_generation_prompt,
                                                                bos_token=self.bos_token or "", eos_token=self.eos_token or "", **kw)
        else:
            text = "".join(f"<|im_start|>{m['role']}\n{m.get('content') or ''}<|im_end|>\n" for m in messages)
            if add_generation_prompt:
                text += "<|im_start|>assistant\n"
        return self.encode(text, add_special_tokens=False) if tokenize else text

    def __len__(self) -> int:
        return self.vocab_size
