"""Medusa speculative heads (reference: vllm/v1/spec_decode/medusa.py, model_executor/models/medusa.py).

k extra heads read the target model's last hidden state of the position that produced the newest token; head i
is ``lm_head_i(h + SiLU(W_i h + b_i))`` and guesses the token i+1 positions ahead.  One tiny batched GEMM pair per
head and no KV cache: the cheapest proposer when a trained head checkpoint exists.  Verification (and therefore
output quality) is the target model's job — the rejection sampler keeps the result lossless whatever the heads say.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn as nn


class MedusaHeads(nn.Module):
    def __init__(self, hidden: int, vocab: int, k: int, num_layers: int = 1, dtype=None) -> None:
        super().__init__()
        self.k, self.num_layers = k, num_layers
        self.blocks = nn.ModuleList([nn.ModuleList([nn.Linear(hidden, hidden, dtype=dtype) for _ in range(num_layers)])
                                     for _ in range(k)])
        self.heads = nn.ModuleList([nn.Linear(hidden, vocab, bias=False, dtype=dtype) for _ in range(k)])

    def forward(self, h: torch.Tensor) -> torch.Tensor:
        """h [B, H] → draft token ids [B, k] (greedy)."""
        out = []
        for blocks, head in zip(self.blocks, self.heads):
            x = h
            for lin in blocks:
                x = x + torch.nn.functional.silu(lin(x))
            out.append(head(x).float().argmax(-1))
        return torch.stack(out, dim=1)

    def load_weights(self, weights) -> int:
        """Checkpoint names: ``medusa_heads.{i}.{j}.linear.{weight,bias}`` (res blocks), ``medusa_heads.{i}.{L}.weight``."""
        n = 0
        params = dict(self.named_parameters())
        for name, w in weights:
            parts = name.split(".")
            if len(parts) < 4 or parts[0] != "medusa_heads":
                continue
            i, j = int(parts[1]), int(parts[2])
            if i >= self.k:
                continue
            key = f"heads.{i}.weight" if j == self.num_layers else f"blocks.{i}.{j}.{parts[-1]}"
            if key in params:
                params[key].data.copy_(w)
                n += 1
        return n


class MedusaDrafter:
    needs_hidden = True

    def __init__(self, vllm_config, runner) -> None:
        sc = vllm_config.speculative_config
        mc = vllm_config.model_config
        self.k = sc.num_speculative_tokens
        self.heads = MedusaHeads(mc.hidden_size, mc.vocab_size, self.k, dtype=mc.torch_dtype).to(runner.device)
        with torch.no_grad():
            for p in self.heads.parameters():          # dummy init keeps the proposer runnable without a checkpoint
                p.uniform_(-0.02, 0.02)
        if sc.model and os.path.isdir(sc.model):
            from vllm_b200.model_loader import iter_weights
            self.heads.load_weights(iter_weights(sc.model, "auto"))
        self.heads.eval()

    @torch.inference_mode()
    def propose(self, runner, rows, sampled: list[list[int]]) -> list[list[int]]:
        out: list[list[int]] = [[] for _ in rows]
        act = [(j, int(r)) for j, r in enumerate(rows) if sampled[j]]
        if not act or runner.spec_hidden is None:
            return out
        idx = torch.tensor([r for _, r in act], dtype=torch.int64, device=runner.device)
        drafts = self.heads(runner.spec_hidden[idx]).cpu().tolist()
        for (j, r), d in zip(act, drafts):
            room = runner.max_model_len - int(runner.num_tokens_cpu[r]) - 1
            out[j] = d[:max(min(self.k, room), 0)]
        return out
