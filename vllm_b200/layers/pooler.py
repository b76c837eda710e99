"""Pooling heads for embedding / classification models (reference: vllm/model_executor/layers/pooler.py)."""
from __future__ import annotations

import torch


def pool_hidden(hidden: torch.Tensor, params, model=None) -> torch.Tensor:
    """hidden: [n_tokens, H] of ONE request."""
    kind = (params.pooling_type or getattr(model, "default_pooling", "LAST")).upper()
    if params.task in ("token_classify", "token_embed"):
        kind = "ALL"
    h = hidden.float()
    if kind == "LAST":
        out = h[-1]
    elif kind == "CLS":
        out = h[0]
    elif kind == "MEAN":
        out = h.mean(dim=0)
    elif kind == "ALL":
        out = h
    else:
        raise ValueError(f"unknown pooling type {kind}")
    if params.task in ("embed", "token_embed") and model is not None and hasattr(model, "project_embedding"):
        out = model.project_embedding(out.to(hidden.dtype), per_token=(kind == "ALL")).float()   # ColBERT / sparse heads
    if params.task in ("classify", "token_classify") and model is not None and hasattr(model, "score"):
        out = model.score(out.to(hidden.dtype)).float()
        if not getattr(model, "raw_scores", False):
            out = out.softmax(-1)
    if params.dimensions:
        out = out[..., :params.dimensions]
    if params.normalize and params.task in ("embed", "token_embed"):
        out = torch.nn.functional.normalize(out, p=2, dim=-1)
    return out
