"""Pooling adapters: any registered decoder becomes an embedding model (``<X>Model`` checkpoints, ``--runner pooling``)
or a sequence classifier / reward model (``<X>ForSequenceClassification``: last-token state → ``score`` head).

Reference parity: vllm/model_executor/models/adapters.py (``as_embedding_model``, ``as_seq_cls_model``,
``as_reward_model``) — there implemented as class wrappers too; the pooling itself lives in ``layers/pooler.py``.
"""
from __future__ import annotations

from typing import Iterable, Type

import torch
import torch.nn as nn


def as_embedding_model(base: Type[nn.Module]) -> Type[nn.Module]:
    class EmbeddingModel(base):                                   # type: ignore[misc, valid-type]
        default_pooling = "LAST"
        is_pooling_model = True

        def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
            def fix():
                for name, w in weights:
                    if name.startswith(("lm_head.", "score.")):
                        continue
                    # bare ``<X>Model`` checkpoints have no "model." prefix
                    yield (name if name.startswith("model.") else "model." + name), w
            return super().load_weights(fix())
    EmbeddingModel.__name__ = base.__name__.replace("ForCausalLM", "") + "EmbeddingModel"
    return EmbeddingModel


def as_seq_cls_model(base: Type[nn.Module]) -> Type[nn.Module]:
    class SeqClsModel(base):                                       # type: ignore[misc, valid-type]
        default_pooling = "LAST"
        is_pooling_model = True

        def __init__(self, vllm_config, prefix: str = "") -> None:
            super().__init__(vllm_config, prefix)
            cfg = vllm_config.model_config.hf_config
            n = int(cfg.get("num_labels") or len(cfg.get("id2label") or {}) or 1)
            self.score = nn.Linear(vllm_config.model_config.hidden_size, n, bias=bool(cfg.get("score_bias", False)),
                                   dtype=vllm_config.model_config.torch_dtype)
            self.num_labels = n

        def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
            loaded: set[str] = set()

            def rest():
                for name, w in weights:
                    if name.startswith("score."):
                        getattr(self.score, name.split(".", 1)[1]).data.copy_(w)
                        loaded.add(name)
                    elif not name.startswith("lm_head."):
                        yield name, w
            out = super().load_weights(rest())
            return out | loaded
    SeqClsModel.__name__ = base.__name__.replace("ForCausalLM", "") + "ForSequenceClassification"
    return SeqClsModel


def as_reward_model(base: Type[nn.Module], n_out: int = 1, mlp: bool = True, head_name: str = "score") -> Type[nn.Module]:
    """Per-token reward heads: ``score`` = Linear(H, H) → ReLU → Linear(H, n) (Qwen2.5-Math RM / PRM) or a single bias-free
    ``v_head`` (InternLM2 reward); every position is returned (``LLM.encode`` with pooling type ALL), the caller picks the
    step positions (reference: models/qwen2_rm.py, internlm2.py ``InternLM2ForRewardModel``)."""
    class RewardModel(base):                                       # type: ignore[misc, valid-type]
        default_pooling = "ALL"
        is_pooling_model = True

        def __init__(self, vllm_config, prefix: str = "") -> None:
            super().__init__(vllm_config, prefix)
            cfg, dt = vllm_config.model_config.hf_config, vllm_config.model_config.torch_dtype
            H = vllm_config.model_config.hidden_size
            self.reward_head = (nn.Sequential(nn.Linear(H, H, dtype=dt), nn.ReLU(), nn.Linear(H, n_out, dtype=dt)) if mlp
                                else nn.Linear(H, n_out, bias=False, dtype=dt))

        def score(self, hidden: torch.Tensor) -> torch.Tensor:
            return self.reward_head(hidden)

        def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
            loaded: set[str] = set()
            own = dict(self.reward_head.named_parameters())

            def rest():
                for name, w in weights:
                    if name.startswith(head_name + "."):
                        own[name[len(head_name) + 1:]].data.copy_(w)
                        loaded.add(name)
                    elif not name.startswith("lm_head."):
                        yield name, w
            return super().load_weights(rest()) | loaded
    RewardModel.__name__ = base.__name__.replace("ForCausalLM", "") + "RewardModel"
    return RewardModel
