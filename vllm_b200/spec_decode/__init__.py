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


class SuffixDrafter:
    """Suffix decoding (reference: v1/spec_decode/suffix_decoding.py): continuation statistics over the request's own
    tokens AND over everything this server generated before (bounded global table), so repetitive agentic / templated
    workloads get long accepted drafts.  A draft follows the most frequent successor chain of the longest suffix
    (``min_n``..``max_n`` tokens) seen so far."""

    def __init__(self, k: int, max_n: int, min_n: int, global_entries: int = 1 << 18) -> None:
        from collections import Counter, OrderedDict
        self.k, self.max_n, self.min_n = k, max(max_n, 2), max(min_n, 1)
        self.Counter = Counter
        self.local: dict[str, dict] = {}            # req id → {n-gram: Counter(next)}
        self.indexed: dict[str, int] = {}           # req id → tokens already indexed
        self.glob: "OrderedDict[tuple, Counter]" = OrderedDict()
        self.global_entries = global_entries

    def _index(self, table: dict, toks: np.ndarray, start: int) -> None:
        L = len(toks)
        for i in range(max(start, 1), L):
            nxt = int(toks[i])
            for n in range(self.min_n, min(self.max_n, i) + 1):
                key = tuple(int(x) for x in toks[i - n:i])
                c = table.get(key)
                if c is None:
                    c = table[key] = self.Counter()
                c[nxt] += 1

    def _follow(self, tables, ctx: list[int], k: int) -> list[int]:
        out: list[int] = []
        seq = list(ctx)
        for _ in range(k):
            nxt = None
            for n in range(min(self.max_n, len(seq)), self.min_n - 1, -1):
                key = tuple(seq[-n:])
                for t in tables:
                    c = t.get(key)
                    if c:
                        nxt = c.most_common(1)[0][0]
                        break
                if nxt is not None:
                    break
            if nxt is None:
                break
            out.append(nxt)
            seq.append(nxt)
        return out

    def propose(self, runner, rows, sampled: list[list[int]]) -> list[list[int]]:
        live = set()
        out = []
        for j, r in enumerate(rows):
            rid = runner.req_ids[r]
            live.add(rid)
            n = int(runner.num_tokens_cpu[r])
            toks = runner.token_ids_cpu[r, :n]
            tab = self.local.setdefault(rid, {})
            self._index(tab, toks, self.indexed.get(rid, 0))
            self.indexed[rid] = n
            room = runner.max_model_len - n - 1
            if not sampled[j] or room <= 0:
                out.append([])
                continue
            out.append(self._follow((tab, self.glob), toks[-self.max_n:].tolist(), min(self.k, room)))
        # requests that left the batch donate their statistics to the global table
        for rid in [x for x in self.local if x not in live and x not in runner.req_index]:
            for key, c in self.local.pop(rid).items():
                g = self.glob.get(key)
                if g is None:
                    self.glob[key] = c
                else:
                    g.update(c)
                    self.glob.move_to_end(key)
            self.indexed.pop(rid, None)
            while len(self.glob) > self.global_entries:
                self.glob.popitem(last=False)
        return out


# Draft / prediction-head architectures as they appear in checkpoints' config.json → proposer that serves them
# (reference: vllm/model_executor/models/registry.py ``_SPECULATIVE_DECODING_MODELS``, config/speculative.py method inference)
DRAFT_ARCHS: dict[str, str] = {
    **{a: "eagle" for a in ("EagleLlamaForCausalLM", "EagleLlama4ForCausalLM", "EagleMistralForCausalLM", "EagleMiniCPMForCausalLM",
                            "EagleCohereForCausalLM", "EagleMistralLarge3ForCausalLM", "EagleDeepSeekMTPModel")},
    **{a: "eagle3" for a in ("Eagle3LlamaForCausalLM", "LlamaForCausalLMEagle3", "Eagle3Qwen3ForCausalLM", "Eagle3Qwen2_5vlForCausalLM",
                             "Eagle3Qwen3vlForCausalLM", "Eagle3MiniMaxM2ForCausalLM", "Eagle3DeepseekV2ForCausalLM",
                             "Eagle3DeepseekV3ForCausalLM", "PeagleLlamaForCausalLM", "PeagleQwen3ForCausalLM", "PEagleDraftModel")},
    **{a: "mtp" for a in ("DeepSeekMTPModel", "DeepseekV32MTPModel", "Glm4MoeMTPModel", "Glm4MoeLiteMTPModel", "GlmOcrMTPModel",
                          "MiMoMTPModel", "MiMoV2MTPModel", "ErnieMTPModel", "Qwen3NextMTP", "Qwen3_5MTP", "Qwen3_5MoeMTP",
                          "ExaoneMoeMTP", "Exaone4_5_MTP", "NemotronHMTPModel", "LongCatFlashMTPModel", "OpenPanguMTPModel",
                          "Step3p5MTP", "BailingMoeV25MTPModel", "BailingMoeV3MTPModel", "HYV3MTPModel", "KimiK3MTPModel",
                          "Gemma4MTPModel", "InklingMTPModel", "Dots3NoteMTPModel", "MiniMaxM3MTP", "MiMoV2OmniMTPModel",
                          "InternS2MobiusMTP", "DeepSeekV4MTPModel")},
    "MedusaModel": "medusa",
}


def infer_method(draft: Optional[str]) -> str:
    """``method="auto"``: no draft checkpoint → n-gram lookup; a draft whose config names a known head architecture → that
    head's proposer; any other causal LM → an independent draft model."""
    if not draft:
        return "ngram"
    import json
    import os
    cfg = None
    p = os.path.join(draft, "config.json")
    if os.path.isfile(p):
        with open(p) as f:
            cfg = json.load(f)
    else:
        from vllm_b200.config import get_preset_hf_config
        cfg = get_preset_hf_config(draft)
    archs = (cfg or {}).get("architectures") or []
    for a in archs:
        if a in DRAFT_ARCHS:
            return DRAFT_ARCHS[a]
    low = " ".join(archs).lower() + " " + os.path.basename(str(draft)).lower()
    if "eagle3" in low:
        return "eagle3"
    if "eagle" in low:
        return "eagle"
    if "medusa" in low:
        return "medusa"
    return "draft_model"


def make_drafter(vllm_config, runner) -> Optional[object]:
    sc = vllm_config.speculative_config
    if sc is None:
        return None
    if sc.method == "ngram":
        return NgramDrafter(sc.num_speculative_tokens, sc.prompt_lookup_max, sc.prompt_lookup_min)
    if sc.method == "suffix":
        return SuffixDrafter(sc.num_speculative_tokens, sc.prompt_lookup_max, sc.prompt_lookup_min)
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
