"""Speculative decoding proposers (reference: vllm/v1/spec_decode/*: n-gram
(ngram_proposer.py), suffix, medusa, draft-model, EAGLE/EAGLE3, MTP)."""
from __future__ import annotations

from typing import Optional

import numpy as np


class NgramDrafter:
    """Prompt-lookup decoding: find the longest suffix n-gram (min_n..max_n) that occurred
    earlier in the sequence and propose the k tokens that followed it."""

    def __init__(self, k: int, max_n: int, min_n: int) -> None:
        self.k, self.max_n, self.min_n = k, max_n, min_n

    def propose_one(self, toks: np.ndarray) -> list[int]:
        L = len(toks)
        for n in range(min(self.max_n, L - 1), self.min_n - 1, -1):
            suffix = toks[L - n:]
            # windows[i] == toks[i:i+n]; search the right-most earlier occurrence
            win = np.lib.stride_tricks.sliding_window_view(toks[:L - 1], n)
            hits = np.nonzero((win == suffix).all(axis=1))[0]
            hits = hits[hits + n < L]
            if len(hits):
                s = int(hits[-1]) + n
                return toks[s:s + self.k].tolist()
        return []

    def propose(self, runner, rows, sampled: list[list[int]]) -> list[list[int]]:
        out = []
        for j, r in enumerate(rows):
            if not sampled[j]:
                out.append([])
                continue
            n = int(runner.num_tokens_cpu[r])
            room = runner.max_model_len - n - 1
            d = self.propose_one(runner.token_ids_cpu[r, :n]) if room > 0 else []
            out.append(d[:max(room, 0)])
        return out


def make_drafter(vllm_config, runner) -> Optional[object]:
    sc = vllm_config.speculative_config
    if sc is None:
        return None
    if sc.method == "ngram":
        return NgramDrafter(sc.num_speculative_tokens, sc.prompt_lookup_max, sc.prompt_lookup_min)
    if sc.method == "medusa":
        from vllm_b200.spec_decode.medusa import MedusaDrafter
        return MedusaDrafter(vllm_config, runner)
    if sc.method in ("eagle", "eagle3", "mtp"):
        from vllm_b200.spec_decode.eagle import EagleDrafter
        return EagleDrafter(vllm_config, runner)
    if sc.method == "draft_model":
        from vllm_b200.spec_decode.draft_model import DraftModelDrafter
        return DraftModelDrafter(vllm_config, runner)
    raise ValueError(f"unknown speculative method {sc.method!r}")
