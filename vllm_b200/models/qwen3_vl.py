"""Qwen3-VL: dynamic-resolution ViT whose learned position table is bilinearly resampled to each image's patch grid
(plus 2-D rotary inside attention), a patch merger, and DeepStack — features tapped at a few ViT depths go through their
own mergers and are ADDED to the decoder's residual stream after its first layers; the decoder is Qwen3 (per-head q/k
norms) with interleaved M-RoPE (frequency j rotates with axis j mod 3).

Reference parity: vllm/model_executor/models/qwen3_vl.py, HF ``modeling_qwen3_vl``.  The tower returns
``[tokens, hidden · (1 + levels)]``: the worker splices the first ``hidden`` columns over the placeholders and hands the
rest to the decoder through the forward context (``models/llama.py``: DeepStack), the same packing the reference uses.
"""
from __future__ import annotations

from typing import Iterable

import torch
import torch.nn as nn
import torch.nn.functional as F

from vllm_b200.config import VllmConfig
from vllm_b200.models.llama import Qwen3ForCausalLM
from vllm_b200.models.qwen2_vl import _rot_half


class _Block(nn.Module):
    def __init__(self, D, nh, I, act, dtype) -> None:
        super().__init__()
        self.nh = nh
        self.norm1, self.norm2 = nn.LayerNorm(D, eps=1e-6, dtype=dtype), nn.LayerNorm(D, eps=1e-6, dtype=dtype)
        self.qkv, self.proj = nn.Linear(D, 3 * D, dtype=dtype), nn.Linear(D, D, dtype=dtype)
        self.fc1, self.fc2 = nn.Linear(D, I, dtype=dtype), nn.Linear(I, D, dtype=dtype)
        self.act = (lambda x: F.gelu(x, approximate="tanh")) if act == "gelu_pytorch_tanh" else F.gelu

    def forward(self, x, cos, sin, frames):
        N, D = x.shape
        q, k, v = self.qkv(self.norm1(x)).view(N, 3, self.nh, D // self.nh).unbind(1)
        q = (q.float() * cos + _rot_half(q.float()) * sin).to(x.dtype)
        k = (k.float() * cos + _rot_half(k.float()) * sin).to(x.dtype)
        q, k, v = (t.reshape(frames, N // frames, self.nh, -1).transpose(1, 2) for t in (q, k, v))
        x = x + self.proj(F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(N, D))
        return x + self.fc2(self.act(self.fc1(self.norm2(x))))


class _Merger(nn.Module):
    def __init__(self, D, m, out, post_shuffle, dtype) -> None:
        super().__init__()
        self.w, self.post = D * m * m, post_shuffle
        self.norm = nn.LayerNorm(self.w if post_shuffle else D, eps=1e-6, dtype=dtype)
        self.linear_fc1, self.linear_fc2 = nn.Linear(self.w, self.w, dtype=dtype), nn.Linear(self.w, out, dtype=dtype)

    def forward(self, x):
        x = self.norm(x.view(-1, self.w)) if self.post else self.norm(x).view(-1, self.w)
        return self.linear_fc2(F.gelu(self.linear_fc1(x)))


class Qwen3VisionTransformer(nn.Module):
    def __init__(self, vc: dict, dtype) -> None:
        super().__init__()
        D, self.patch, self.merge = vc["hidden_size"], vc.get("patch_size", 16), vc.get("spatial_merge_size", 2)
        self.tp, self.nh = vc.get("temporal_patch_size", 2), vc["num_heads"]
        self.patch_embed = nn.Linear(vc.get("in_channels", 3) * self.tp * self.patch ** 2, D, bias=True, dtype=dtype)
        self.pos_embed = nn.Embedding(vc["num_position_embeddings"], D, dtype=dtype)
        self.side = int(vc["num_position_embeddings"] ** 0.5)
        self.blocks = nn.ModuleList([_Block(D, self.nh, vc["intermediate_size"], vc.get("hidden_act", "gelu_pytorch_tanh"), dtype)
                                     for _ in range(vc["depth"])])
        out = vc["out_hidden_size"]
        self.merger = _Merger(D, self.merge, out, False, dtype)
        self.taps = list(vc.get("deepstack_visual_indexes") or [])
        self.deepstack_merger_list = nn.ModuleList([_Merger(D, self.merge, out, True, dtype) for _ in self.taps])
        hd = D // self.nh
        self.register_buffer("inv_freq", 1.0 / (10000.0 ** (torch.arange(0, hd // 2, 2, dtype=torch.float32) / (hd // 2))),
                             persistent=False)

    def _block_order(self, t, h, w, x):
        """[h·w, …] row-major → merge-block order (the order the patches arrive in), repeated for t temporal patches."""
        m = self.merge
        x = x.view(h // m, m, w // m, m, -1).permute(0, 2, 1, 3, 4).reshape(h * w, -1)
        return x.repeat(t, 1)

    def _pos(self, grid):
        t, h, w = grid
        g, dev = self.side, self.pos_embed.weight.device
        hi, wi = torch.linspace(0, g - 1, h), torch.linspace(0, g - 1, w)
        h0, w0 = hi.int(), wi.int()
        h1, w1 = (h0 + 1).clip(max=g - 1), (w0 + 1).clip(max=g - 1)
        dh, dw = hi - h0, wi - w0
        tab = self.pos_embed.weight
        def at(a, b):
            return tab[(a[:, None] * g + b[None, :]).flatten().long().to(dev)]
        wts = [((1 - dh)[:, None] * (1 - dw)[None]), ((1 - dh)[:, None] * dw[None]), (dh[:, None] * (1 - dw)[None]), (dh[:, None] * dw[None])]
        vals = [at(h0, w0), at(h0, w1), at(h1, w0), at(h1, w1)]
        pe = sum(v * wt.flatten()[:, None].to(dev, tab.dtype) for v, wt in zip(vals, wts))
        return self._block_order(t, h, w, pe)

    def _rope(self, grid):
        t, h, w = grid
        dev = self.inv_freq.device
        hp = self._block_order(1, h, w, torch.arange(h).unsqueeze(1).expand(-1, w).reshape(-1, 1).float())[:, 0].long()
        wp = self._block_order(1, h, w, torch.arange(w).unsqueeze(0).expand(h, -1).reshape(-1, 1).float())[:, 0].long()
        ids = torch.stack([hp, wp], dim=-1).repeat(t, 1).to(dev)
        table = torch.outer(torch.arange(max(h, w), device=dev, dtype=torch.float32), self.inv_freq)
        f = table[ids].flatten(1)
        emb = torch.cat((f, f), dim=-1)
        return emb.cos()[:, None, :], emb.sin()[:, None, :]

    def forward(self, patches: torch.Tensor, grid) -> torch.Tensor:
        x = self.patch_embed(patches.to(self.patch_embed.weight.dtype))
        x = x + self._pos(grid).to(x.dtype)
        cos, sin = self._rope(grid)
        deep = []
        for i, b in enumerate(self.blocks):
            x = b(x, cos, sin, grid[0])
            if i in self.taps:
                deep.append(self.deepstack_merger_list[self.taps.index(i)](x))
        return torch.cat([self.merger(x)] + deep, dim=-1)


class _Qwen3VLMixin:
    uses_mrope = True

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        mc = vllm_config.model_config
        cfg = mc.hf_config
        rp = cfg.get("rope_parameters") or cfg.get("rope_scaling") or {}
        sec = rp.get("mrope_section") or [24, 20, 20]
        cfg["rope_scaling"] = {"rope_type": "mrope", "mrope_section": list(sec), "mrope_interleaved": bool(rp.get("mrope_interleaved", True))}
        cfg.setdefault("rope_theta", rp.get("rope_theta", 5000000.0))
        super().__init__(vllm_config, prefix)
        self.visual = Qwen3VisionTransformer(cfg["vision_config"], mc.torch_dtype)

    @torch.inference_mode()
    def embed_multimodal_item(self, feature: dict) -> torch.Tensor:
        dev = self.visual.patch_embed.weight.device
        return self.visual(torch.as_tensor(feature["data"]).to(dev), tuple(feature["grid"]))

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        vparams = dict(self.visual.named_parameters())
        loaded: set[str] = set()
        rest = []
        ren = [("patch_embed.proj.", "patch_embed."), (".attn.qkv.", ".qkv."), (".attn.proj.", ".proj."),
               (".mlp.linear_fc1.", ".fc1."), (".mlp.linear_fc2.", ".fc2.")]
        for name, w in weights:
            if name.startswith(("model.visual.", "visual.")):
                n = name.split("visual.", 1)[1]
                for a, b in ren:
                    n = n.replace(a, b)
                if n in vparams:
                    p = vparams[n]
                    p.data.copy_(w.reshape(p.shape))
                    loaded.add("visual." + n)
                continue
            if name.startswith("model.language_model."):
                name = "model." + name[len("model.language_model."):]
            rest.append((name, w))
        if isinstance(self, _VisualOnly):
            return loaded
        return loaded | super().load_weights(rest)


class Qwen3VLForConditionalGeneration(_Qwen3VLMixin, Qwen3ForCausalLM):
    pass


def _moe_base():
    from vllm_b200.models.mixtral import Qwen3MoeForCausalLM
    return Qwen3MoeForCausalLM


class Qwen3VLMoeForConditionalGeneration(_Qwen3VLMixin, _moe_base()):
    """The same tower and DeepStack path in front of the Qwen3-MoE decoder (reference: models/qwen3_vl_moe.py)."""


def _q35_base():
    from vllm_b200.models.qwen3_next import Qwen3_5ForCausalLM
    return Qwen3_5ForCausalLM


class Qwen3_5ForConditionalGeneration(_Qwen3VLMixin, _q35_base()):
    """Qwen3.5 vision-language (dense and MoE): the Qwen3-VL tower without DeepStack in front of the gated-delta-net hybrid
    decoder; the full-attention layers rotate with interleaved M-RoPE (reference: models/qwen3_5.py)."""

    def load_weights(self, weights):
        vis, rest = [], []
        for n, w in weights:
            (vis if n.startswith(("model.visual.", "visual.")) else rest).append((n, w))
        loaded = _Qwen3VLMixin.load_weights(_VisualOnly(self), vis)
        return loaded | _q35_base().load_weights(self, rest)


class _VisualOnly:
    """Runs the mixin's visual-weight loop against a model without touching its decoder loader."""

    def __init__(self, model) -> None:
        self.visual = model.visual

    def __getattr__(self, name):
        raise AttributeError(name)


class Qwen3_5MoeForConditionalGeneration(Qwen3_5ForConditionalGeneration):
    pass
