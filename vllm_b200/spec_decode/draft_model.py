"""Draft-model speculative proposer.

Behavioural reference: vllm/v1/spec_decode/draft_model.py (a small model of the same vocabulary proposes k
tokens per step, the target verifies them in one forward pass through the rejection sampler).

Design: the drafter owns a *private* KV arena (``max_num_seqs`` rows of ``window + k`` slots) and never
touches the engine's paged cache: each proposal round prefills the last ``window`` tokens of every row that
needs drafts in ONE batched forward pass and then runs k-1 single-token decode passes against the arena.  No
scheduler / block-manager coupling, no cache invalidation on rejection, and the cost per round is bounded by
the window, not by the sequence length.
"""
from __future__ import annotations

import copy
from typing import Optional

import numpy as np
import torch

from vllm_b200.config import ModelConfig, VllmConfig
from vllm_b200.forward_context import AttentionMetadata, ForwardContext, set_forward_context
from vllm_b200.logger import init_logger
from vllm_b200.utils import cdiv

logger = init_logger(__name__)


class DraftModelDrafter:
    def __init__(self, vllm_config: VllmConfig, runner, window: int = 256) -> None:
        from vllm_b200.layers import attention as attn_mod
        from vllm_b200.model_loader import get_model
        sc = vllm_config.speculative_config
        if not sc.model:
            raise ValueError(f"speculative method {sc.method!r} needs speculative_config['model']")
        self.k = sc.num_speculative_tokens
        self.device = runner.device
        cfg = copy.copy(vllm_config)
        tmc = vllm_config.model_config
        cfg.model_config = ModelConfig(model=sc.model, dtype=tmc.dtype, seed=tmc.seed,
                                       max_model_len=tmc.max_model_len)
        cfg.speculative_config = None
        before = set(attn_mod.attention_layers())
        self.model = get_model(cfg, self.device, prefix="draft.")
        # the draft's attention layers must not be sized into the engine's KV cache: move them to our own table
        reg = attn_mod.attention_layers()
        self.layers = {n: reg.pop(n) for n in list(reg) if n not in before}
        if not self.layers:
            raise RuntimeError("draft model registered no attention layers (name clash with the target?)")
        dmc = cfg.model_config
        self.vocab = dmc.hf_config["vocab_size"]
        if self.vocab != tmc.hf_config["vocab_size"]:
            raise ValueError("draft and target models must share a vocabulary")
        self.bs = vllm_config.cache_config.block_size
        self.window = min(window, runner.max_model_len)
        self.blocks_per_row = cdiv(self.window + self.k, self.bs)
        R = runner.max_num_reqs
        self.kv = {}
        dt = dmc.torch_dtype
        for name, layer in self.layers.items():
            shape = (R * self.blocks_per_row, 2, layer.num_kv_heads, self.bs, layer.head_size)
            self.kv[name] = torch.zeros(shape, dtype=dt, device=self.device)
        self.block_table = torch.arange(R * self.blocks_per_row, dtype=torch.int32,
                                        device=self.device).view(R, self.blocks_per_row)
        self.layer_to_group = {n: 0 for n in self.layers}
        logger.info("draft model %s: %d layers, private KV arena %d rows x %d slots", sc.model, len(self.layers), R,
                    self.blocks_per_row * self.bs)

    def _run(self, ids: torch.Tensor, pos: torch.Tensor, md: AttentionMetadata, last_idx: torch.Tensor) -> torch.Tensor:
        ctx = ForwardContext(attn_metadata={0: md}, kv_caches=self.kv, layer_to_group=self.layer_to_group,
                             num_tokens=ids.numel())
        with set_forward_context(ctx):
            hidden = self.model(ids, pos)
        return self.model.compute_logits(hidden[last_idx]).float().argmax(-1)

    @torch.inference_mode()
    def propose(self, runner, rows, sampled: list[list[int]]) -> list[list[int]]:
        out: list[list[int]] = [[] for _ in rows]
        act = []
        for j, r in enumerate(rows):
            n = int(runner.num_tokens_cpu[r])
            if sampled[j] and runner.max_model_len - n - 1 > 0:
                act.append((j, int(r), n))
        if not act:
            return out
        dev, bs, W = self.device, self.bs, self.window
        B = len(act)
        lens = np.array([min(n, W) for _, _, n in act], dtype=np.int64)
        ids = np.concatenate([runner.token_ids_cpu[r, n - l:n] for (_, r, n), l in zip(act, lens)]).astype(np.int64)
        pos = np.concatenate([np.arange(n - l, n) for (_, _, n), l in zip(act, lens)]).astype(np.int64)
        arena_rows = np.arange(B, dtype=np.int64)            # arena row b ↔ active request b
        base = arena_rows * self.blocks_per_row * bs
        slots = np.concatenate([b + np.arange(l) for b, l in zip(base, lens)])
        qsl = np.zeros(B + 1, dtype=np.int32)
        qsl[1:] = np.cumsum(lens)
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)  # noqa: E731
        md = AttentionMetadata(num_reqs=B, num_actual_tokens=int(qsl[-1]), max_query_len=int(lens.max()),
                               max_seq_len=int(lens.max()), query_start_loc=t(qsl, torch.int32),
                               seq_lens=t(lens, torch.int32), block_table=self.block_table[:B],
                               slot_mapping=t(slots, torch.int64), query_start_loc_cpu=torch.from_numpy(qsl),
                               seq_lens_cpu=torch.from_numpy(lens.astype(np.int32)))
        last = t(qsl[1:] - 1, torch.int64)
        nxt = self._run(t(ids, torch.int64), t(pos, torch.int64), md, last)
        drafts = [nxt]
        room = np.array([runner.max_model_len - n - 1 for _, _, n in act])
        one = torch.arange(B + 1, dtype=torch.int32, device=dev)
        for step in range(1, self.k):
            cur = lens + step                      # tokens in the arena row after this step's write
            p = np.array([n + step - 1 for _, _, n in act], dtype=np.int64)
            sl = base + cur - 1
            md = AttentionMetadata(num_reqs=B, num_actual_tokens=B, max_query_len=1, max_seq_len=int(cur.max()),
                                   query_start_loc=one, seq_lens=t(cur, torch.int32), block_table=self.block_table[:B],
                                   slot_mapping=t(sl, torch.int64), query_start_loc_cpu=torch.arange(B + 1, dtype=torch.int32),
                                   seq_lens_cpu=torch.from_numpy(cur.astype(np.int32)), decode_only=True)
            nxt = self._run(nxt, t(p, torch.int64), md, torch.arange(B, device=dev))
            drafts.append(nxt)
        mat = torch.stack(drafts, dim=1).cpu().numpy()        # [B, k]
        for b, (j, _, _) in enumerate(act):
            out[j] = mat[b, :max(min(self.k, int(room[b])), 0)].tolist()
        return out
