"""Qwen2-VL: dynamic-resolution ViT (2-D rotary, 2x2 patch merger) + Qwen2 decoder with multimodal rotary positions.

Reference parity: vllm/model_executor/models/qwen2_vl.py (``Qwen2VisionTransformer``, ``Qwen2VLForConditionalGeneration``),
layers/rotary_embedding/mrope.py.
"""
from __future__ import annotations

from typing import Iterable

import torch
import torch.nn as nn
import torch.nn.functional as F

from vllm_b200.config import VllmConfig
from vllm_b200.models.llama import Qwen2ForCausalLM


def _rot_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


class _VisionBlock(nn.Module):
    def __init__(self, D: int, nh: int, mlp: int, dtype) -> None:
        super().__init__()
        self.nh = nh
        self.norm1, self.norm2 = nn.LayerNorm(D, eps=1e-6, dtype=dtype), nn.LayerNorm(D, eps=1e-6, dtype=dtype)
        self.qkv, self.proj = nn.Linear(D, 3 * D, dtype=dtype), nn.Linear(D, D, dtype=dtype)
        self.fc1, self.fc2 = nn.Linear(D, mlp, dtype=dtype), nn.Linear(mlp, D, dtype=dtype)

    def forward(self, x, cos, sin, frames: int = 1):               # x [N, D] patches of ONE image / video
        N, D = x.shape
        q, k, v = self.qkv(self.norm1(x)).view(N, 3, self.nh, D // self.nh).unbind(1)
        q = (q.float() * cos + _rot_half(q.float()) * sin).to(x.dtype)
        k = (k.float() * cos + _rot_half(k.float()) * sin).to(x.dtype)
        # every temporal patch attends within itself: [frames, heads, N/frames, hd]
        q, k, v = (t.reshape(frames, N // frames, self.nh, -1).transpose(1, 2) for t in (q, k, v))
        a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2)
        x = x + self.proj(a.reshape(N, D))
        h = self.fc1(self.norm2(x))
        return x + self.fc2(h * torch.sigmoid(1.702 * h))          # quick_gelu


class Qwen2VisionTransformer(nn.Module):
    def __init__(self, vc: dict, dtype) -> None:
        super().__init__()
        D, self.patch, self.merge = vc["embed_dim"], vc.get("patch_size", 14), vc.get("spatial_merge_size", 2)
        self.tp = vc.get("temporal_patch_size", 2)
        self.nh = vc["num_heads"]
        self.patch_embed = nn.Linear(vc.get("in_channels", 3) * self.tp * self.patch ** 2, D, bias=False, dtype=dtype)
        self.blocks = nn.ModuleList([_VisionBlock(D, self.nh, int(D * vc.get("mlp_ratio", 4)), dtype) for _ in range(vc["depth"])])
        self.ln_q = nn.LayerNorm(D, eps=1e-6, dtype=dtype)
        m = D * self.merge ** 2
        self.merger_fc1, self.merger_fc2 = nn.Linear(m, m, dtype=dtype), nn.Linear(m, vc["hidden_size"], dtype=dtype)
        hd = D // self.nh
        self.register_buffer("inv_freq", 1.0 / (10000.0 ** (torch.arange(0, hd // 2, 2, dtype=torch.float32) / (hd // 2))),
                             persistent=False)

    def _rope(self, grid):
        t, h, w = grid
        m = self.merge
        hp = torch.arange(h).unsqueeze(1).expand(-1, w).reshape(h // m, m, w // m, m).permute(0, 2, 1, 3).flatten()
        wp = torch.arange(w).unsqueeze(0).expand(h, -1).reshape(h // m, m, w // m, m).permute(0, 2, 1, 3).flatten()
        ids = torch.stack([hp, wp], dim=-1).repeat(t, 1).to(self.inv_freq.device)
        table = torch.outer(torch.arange(max(h, w), device=self.inv_freq.device, dtype=torch.float32), self.inv_freq)
        f = table[ids].flatten(1)                                       # [N, hd/2]
        emb = torch.cat((f, f), dim=-1)
        return emb.cos()[:, None, :], emb.sin()[:, None, :]

    def forward(self, patches: torch.Tensor, grid) -> torch.Tensor:
        """patches [N, C*tp*p*p] of one image (merge-block order) → [N / merge^2, hidden] tokens for the decoder."""
        x = self.patch_embed(patches.to(self.patch_embed.weight.dtype))
        cos, sin = self._rope(grid)
        for b in self.blocks:
            x = b(x, cos, sin, grid[0])
        x = self.ln_q(x).view(-1, x.shape[-1] * self.merge ** 2)
        return self.merger_fc2(F.gelu(self.merger_fc1(x)))


class Qwen2VLForConditionalGeneration(Qwen2ForCausalLM):
    uses_mrope = True

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        mc = vllm_config.model_config
        cfg = mc.hf_config
        tc = cfg.get("text_config") or {}
        for k, v in tc.items():                                # the decoder reads the flat config
            cfg.setdefault(k, v)
        rp = cfg.get("rope_parameters") or tc.get("rope_parameters") or {}
        sec = (cfg.get("rope_scaling") or {}).get("mrope_section") or rp.get("mrope_section") or [16, 24, 24]
        cfg["rope_scaling"] = {"rope_type": "mrope", "mrope_section": list(sec)}
        cfg.setdefault("rope_theta", rp.get("rope_theta", 1000000.0))
        super().__init__(vllm_config, prefix)
        if type(self) is Qwen2VLForConditionalGeneration:
            self.visual = Qwen2VisionTransformer(cfg["vision_config"], mc.torch_dtype)

    @torch.inference_mode()
    def embed_multimodal_item(self, feature: dict) -> torch.Tensor:
        dev = self.visual.patch_embed.weight.device
        return self.visual(torch.as_tensor(feature["data"]).to(dev), tuple(feature["grid"]))

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        vparams = dict(self.visual.named_parameters())
        loaded: set[str] = set()
        rest = []
        ren = [("patch_embed.proj.", "patch_embed."), ("merger.ln_q.", "ln_q."), ("merger.mlp.0.", "merger_fc1."),
               ("merger.mlp.2.", "merger_fc2."), (".attn.qkv.", ".qkv."), (".attn.proj.", ".proj."),
               (".mlp.fc1.", ".fc1."), (".mlp.fc2.", ".fc2.")]
        for name, w in weights:
            if name.startswith("model.visual.") or name.startswith("visual."):
                n = name.split("visual.", 1)[1]
                for a, b in ren:
                    n = n.replace(a, b)
                if n in vparams:
                    p = vparams[n]
                    p.data.copy_(w.reshape(p.shape))               # conv3d kernel → one matrix row per output channel
                    loaded.add("visual." + n)
                continue
            if name.startswith("model.language_model."):
                name = "model." + name[len("model.language_model."):]
            rest.append((name, w))
        return loaded | super().load_weights(rest)


# ------------------------------------------------------------------------------------------------ Qwen2.5-VL
class _VisionBlock25(nn.Module):
    """RMSNorm + biased SwiGLU MLP; attention over whole image or within windows (segments of the reordered sequence)."""

    def __init__(self, D: int, nh: int, inter: int, dtype) -> None:
        super().__init__()
        from vllm_b200.layers.common import RMSNorm
        self.nh = nh
        self.norm1, self.norm2 = RMSNorm(D, 1e-6, dtype), RMSNorm(D, 1e-6, dtype)
        self.qkv, self.proj = nn.Linear(D, 3 * D, dtype=dtype), nn.Linear(D, D, dtype=dtype)
        self.gate_proj, self.up_proj = nn.Linear(D, inter, dtype=dtype), nn.Linear(D, inter, dtype=dtype)
        self.down_proj = nn.Linear(inter, D, dtype=dtype)

    def forward(self, x, cos, sin, bounds):
        N, D = x.shape
        q, k, v = self.qkv(self.norm1(x)).view(N, 3, self.nh, D // self.nh).unbind(1)
        q = (q.float() * cos + _rot_half(q.float()) * sin).to(x.dtype)
        k = (k.float() * cos + _rot_half(k.float()) * sin).to(x.dtype)
        out = torch.empty(N, self.nh, D // self.nh, dtype=x.dtype, device=x.device)
        for s, e in zip(bounds[:-1], bounds[1:]):
            out[s:e] = F.scaled_dot_product_attention(q[s:e].transpose(0, 1), k[s:e].transpose(0, 1),
                                                      v[s:e].transpose(0, 1)).transpose(0, 1)
        x = x + self.proj(out.reshape(N, D))
        h = self.norm2(x)
        return x + self.down_proj(F.silu(self.gate_proj(h)) * self.up_proj(h))


class Qwen2_5_VisionTransformer(nn.Module):
    def __init__(self, vc: dict, dtype) -> None:
        super().__init__()
        from vllm_b200.layers.common import RMSNorm
        D, self.patch, self.merge = vc["hidden_size"], vc.get("patch_size", 14), vc.get("spatial_merge_size", 2)
        self.tp, self.nh = vc.get("temporal_patch_size", 2), vc["num_heads"]
        self.window = vc.get("window_size", 112) // self.merge // self.patch        # window edge in merged tokens
        self.full_blocks = set(vc.get("fullatt_block_indexes") or [])
        self.patch_embed = nn.Linear(vc.get("in_channels", 3) * self.tp * self.patch ** 2, D, bias=False, dtype=dtype)
        self.blocks = nn.ModuleList([_VisionBlock25(D, self.nh, vc["intermediate_size"], dtype) for _ in range(vc["depth"])])
        self.ln_q = RMSNorm(D, 1e-6, dtype)
        m = D * self.merge ** 2
        self.merger_fc1, self.merger_fc2 = nn.Linear(m, m, dtype=dtype), nn.Linear(m, vc["out_hidden_size"], dtype=dtype)
        hd = D // self.nh
        self.register_buffer("inv_freq", 1.0 / (10000.0 ** (torch.arange(0, hd // 2, 2, dtype=torch.float32) / (hd // 2))),
                             persistent=False)

    _rope = Qwen2VisionTransformer._rope

    def _windows(self, grid):
        """Order of the merged tokens grouped by window, and the patch-level segment bounds of the windows."""
        t, h, w = grid
        lh, lw, ws = h // self.merge, w // self.merge, self.window
        idx = torch.arange(t * lh * lw).reshape(t, lh, lw)
        ph, pw = ws - lh % ws, ws - lw % ws
        pad = F.pad(idx, (0, pw, 0, ph), value=-100)
        nh_, nw_ = (lh + ph) // ws, (lw + pw) // ws
        pad = pad.reshape(t, nh_, ws, nw_, ws).permute(0, 1, 3, 2, 4).reshape(t, nh_ * nw_, ws, ws)
        lens = (pad != -100).sum([2, 3]).reshape(-1)
        order = pad.reshape(-1)
        order = order[order != -100]
        unit = self.merge ** 2
        bounds = [0] + (lens.cumsum(0) * unit).tolist()
        bounds = [b for i, b in enumerate(bounds) if i == 0 or b != bounds[i - 1]]          # drop empty windows
        return order, bounds

    def forward(self, patches: torch.Tensor, grid) -> torch.Tensor:
        x = self.patch_embed(patches.to(self.patch_embed.weight.dtype))
        cos, sin = self._rope(grid)
        N, unit = x.shape[0], self.merge ** 2
        order, wbounds = self._windows(grid)
        order = order.to(x.device)
        fbounds = [f * (N // grid[0]) for f in range(grid[0] + 1)]          # "full" attention is per temporal patch
        # the patch sequence is already in merge-block order: reorder whole 2x2 units by window
        x = x.view(N // unit, unit, -1)[order].reshape(N, -1)
        cos = cos.view(N // unit, unit, *cos.shape[1:])[order].reshape(N, *cos.shape[1:])
        sin = sin.view(N // unit, unit, *sin.shape[1:])[order].reshape(N, *sin.shape[1:])
        for i, b in enumerate(self.blocks):
            x = b(x, cos, sin, fbounds if i in self.full_blocks else wbounds)
        x = self.ln_q(x).view(-1, x.shape[-1] * unit)
        x = self.merger_fc2(F.gelu(self.merger_fc1(x)))
        return x[torch.argsort(order)]


class Qwen2_5_VLForConditionalGeneration(Qwen2VLForConditionalGeneration):
    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__(vllm_config, prefix)
        mc = vllm_config.model_config
        self.visual = Qwen2_5_VisionTransformer(mc.hf_config["vision_config"], mc.torch_dtype)

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        def ren():
            for name, w in weights:
                if "visual." in name:
                    name = (name.replace(".mlp.gate_proj.", ".gate_proj.").replace(".mlp.up_proj.", ".up_proj.")
                                .replace(".mlp.down_proj.", ".down_proj."))
                yield name, w
        return super().load_weights(ren())
