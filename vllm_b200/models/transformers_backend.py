"""Transformers back-end: run ANY HuggingFace decoder-only architecture that is not implemented natively.

Reference: vllm/model_executor/models/transformers/ (the "Transformers backend" fallback behind
``registry.py``): the HF modeling code provides embeddings / projections / MLPs / norms / rotary embedding, while
attention itself is routed — through ``transformers.AttentionInterface`` — to this engine's paged-KV attention
layers, so continuous batching, chunked prefill, prefix caching and CUDA kernels all apply.  Slower than a native
model (no fused QKV / fused collectives, tensor_parallel_size must be 1) but it makes the long tail of ~250
architectures usable on day one.
"""
from __future__ import annotations

import os
from typing import Iterable

import torch
import torch.nn as nn

from vllm_b200.config import VllmConfig
from vllm_b200.layers.attention import Attention
from vllm_b200.logger import init_logger
from vllm_b200.parallel.state import get_tp_group

logger = init_logger(__name__)
_IMPL = "vllm_b200"
_registered = False


def _vb_attention(module, query, key, value, attention_mask, scaling=None, **kwargs):
    """HF attention-interface hook: query [1, H, T, D] (rotary already applied) → paged attention → [1, T, H, D]."""
    layer: Attention = module._vb_attn
    _, H, T, D = query.shape
    q = query[0].transpose(0, 1).reshape(T, H * D)
    k = key[0].transpose(0, 1).reshape(T, -1)
    v = value[0].transpose(0, 1).reshape(T, -1)
    out = layer(q.contiguous(), k.contiguous(), v.contiguous())
    return out.view(1, T, H, D), None


def _ensure_registered() -> None:
    global _registered
    if _registered:
        return
    from transformers import AttentionInterface
    AttentionInterface.register(_IMPL, _vb_attention)
    try:
        from transformers import AttentionMaskInterface
        AttentionMaskInterface.register(_IMPL, lambda *a, **k: None)      # causality is enforced by our kernels
    except Exception:  # pragma: no cover - older transformers
        pass
    _registered = True


class TransformersForCausalLM(nn.Module):
    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        import transformers
        if get_tp_group().world_size != 1:
            raise NotImplementedError("the Transformers back-end runs with tensor_parallel_size=1")
        _ensure_registered()
        mc = vllm_config.model_config
        self.cfg = mc.hf_config
        if os.path.isdir(mc.model):
            hf_cfg = transformers.AutoConfig.from_pretrained(mc.model, trust_remote_code=mc.trust_remote_code)
        else:
            kw = {k: v for k, v in self.cfg.items() if k not in ("architectures", "model_type", "torch_dtype")}
            hf_cfg = transformers.AutoConfig.for_model(self.cfg["model_type"], **kw)
        self.hf = transformers.AutoModelForCausalLM.from_config(hf_cfg, attn_implementation=_IMPL,
                                                                torch_dtype=mc.torch_dtype).eval()
        for p in self.hf.parameters():
            p.requires_grad_(False)
        tc = hf_cfg.get_text_config() if hasattr(hf_cfg, "get_text_config") else hf_cfg
        nh = tc.num_attention_heads
        nkv = getattr(tc, "num_key_value_heads", None) or nh
        n = 0
        for name, mod in self.hf.named_modules():
            if not (hasattr(mod, "layer_idx") and type(mod).__name__.endswith("Attention")):
                continue
            hd = getattr(mod, "head_dim", None) or tc.hidden_size // nh
            scale = float(getattr(mod, "scaling", None) or hd ** -0.5)
            mod._vb_attn = Attention(nh, hd, scale, nkv, sliding_window=getattr(mod, "sliding_window", None),
                                     logits_soft_cap=float(getattr(mod, "attn_logit_softcapping", None) or 0.0),
                                     prefix=f"{prefix}hf.{name}.attn")
            n += 1
        if n == 0:
            raise RuntimeError(f"{type(self.hf).__name__}: no attention modules found for the Transformers back-end")
        logger.info("Transformers back-end: %s with %d attention layers routed to paged attention",
                    type(self.hf).__name__, n)

    def embed_input_ids(self, ids: torch.Tensor) -> torch.Tensor:
        return self.hf.get_input_embeddings()(ids)

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        kw = dict(position_ids=positions[None], use_cache=False)
        if inputs_embeds is not None:
            out = self.hf.base_model(inputs_embeds=inputs_embeds[None], **kw)
        else:
            out = self.hf.base_model(input_ids=input_ids[None], **kw)
        return out.last_hidden_state[0]

    def compute_logits(self, hidden: torch.Tensor) -> torch.Tensor:
        return self.hf.get_output_embeddings()(hidden)

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        params = dict(self.hf.named_parameters())
        params.update(dict(self.hf.named_buffers()))
        loaded: set[str] = set()
        for name, w in weights:
            p = params.get(name)
            if p is None or p.shape != w.shape:
                continue
            p.data.copy_(w)
            loaded.add("hf." + name)
        self.hf.tie_weights()
        return loaded | {"hf." + n for n, _ in self.hf.named_parameters()}     # tied / derived tensors count as loaded
