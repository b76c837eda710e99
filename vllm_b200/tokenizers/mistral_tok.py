"""``tokenizer_mode="mistral"``: mistral-common tokenizers (tekken.json / SentencePiece ``tokenizer.model.v*``) behind the
``TokenizerLike`` surface, with mistral-common's own chat-request encoder instead of a Jinja template
(reference: vllm/tokenizers/mistral.py)."""
from __future__ import annotations

import glob
import os
from typing import Optional, Sequence


class MistralTokenizer:
    is_fast = True

    def __init__(self, mt) -> None:
        self.mt = mt
        self.tok = mt.instruct_tokenizer.tokenizer
        self.vocab_size = self.tok.n_words
        self.bos_token_id, self.eos_token_id = self.tok.bos_id, self.tok.eos_id
        self.pad_token_id = getattr(self.tok, "pad_id", None)
        self.all_special_ids = {i for i in (self.bos_token_id, self.eos_token_id, self.pad_token_id) if i is not None and i >= 0}
        self.chat_template = None

    @classmethod
    def from_dir(cls, path: str) -> "MistralTokenizer":
        from mistral_common.tokens.tokenizers.mistral import MistralTokenizer as MT
        files = sorted(glob.glob(os.path.join(path, "tekken.json")) + glob.glob(os.path.join(path, "tokenizer.model.v*"))
                       + glob.glob(os.path.join(path, "tokenizer.model")))
        if not files:
            raise FileNotFoundError(f"no mistral tokenizer file (tekken.json / tokenizer.model[.v*]) in {path}")
        return cls(MT.from_file(files[0]))

    def encode(self, text: str, add_special_tokens: bool = True, **kw) -> list[int]:
        return self.tok.encode(text, bos=add_special_tokens, eos=False)

    def __call__(self, text, add_special_tokens: bool = True, **kw):
        if isinstance(text, str):
            return {"input_ids": self.encode(text, add_special_tokens)}
        return {"input_ids": [self.encode(t, add_special_tokens) for t in text]}

    def decode(self, ids: Sequence[int], skip_special_tokens: bool = True, **kw) -> str:
        ids = [int(i) for i in ids if not (skip_special_tokens and int(i) in self.all_special_ids)]
        return self.tok.decode(ids)

    def convert_ids_to_tokens(self, ids: Sequence[int], skip_special_tokens: bool = False) -> list[str]:
        return [self.tok.id_to_piece(int(i)) for i in ids if not (skip_special_tokens and int(i) in self.all_special_ids)]

    def convert_tokens_to_string(self, tokens: list[str]) -> str:
        return "".join(tokens).replace("▁", " ")

    def get_vocab(self) -> dict[str, int]:
        return {self.tok.id_to_piece(i): i for i in range(self.vocab_size)}

    def apply_chat_template(self, messages: list[dict], add_generation_prompt: bool = True, tokenize: bool = False,
                            tools: Optional[list] = None, **kw):
        from mistral_common.protocol.instruct.request import ChatCompletionRequest
        req = ChatCompletionRequest(messages=messages, tools=tools or None)
        enc = self.mt.encode_chat_completion(req)
        return enc.tokens if tokenize else enc.text

    def __len__(self) -> int:
        return self.vocab_size
