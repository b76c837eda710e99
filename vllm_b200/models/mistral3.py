"""Mistral-Small-3.x / Pixtral-style vision-language: Pixtral ViT (bias-free patch conv, RMSNorm, 2-D rotary positions over
the patch grid, SwiGLU blocks) → RMSNorm → learned 2×2 patch merger → 2-layer GELU projector → Mistral decoder.

Reference parity: vllm/model_executor/models/mistral3.py, pixtral.py (``PixtralHFVisionModel``), HF ``modeling_pixtral`` /
``modeling_mistral3``.  One image = one ``image_size``-square tile here (the tower itself is resolution-agnostic: the
rotary table is indexed by (row, column)); the prompt carries ``[IMG]`` × columns + ``[IMG_BREAK]`` per row and ``[IMG_END]``
after the last row, which ``PixtralLayoutProcessor`` lays out — only the ``[IMG]`` positions receive tower outputs.
"""
from __future__ import annotations

from typing import Any, Iterable

import torch
import torch.nn as nn
import torch.nn.functional as F

from vllm_b200.config import VllmConfig
from vllm_b200.layers.common import RMSNorm


def _rot_half(x):
    a, b = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-b, a), dim=-1)


class PixtralLayer(nn.Module):
    def __init__(self, vc: dict, dtype) -> None:
        super().__init__()
        H, self.nh = vc["hidden_size"], vc["num_attention_heads"]
        self.hd = vc.get("head_dim") or H // self.nh
        I = vc["intermediate_size"]
        self.q_proj, self.k_proj, self.v_proj = (nn.Linear(H, self.nh * self.hd, bias=False, dtype=dtype) for _ in range(3))
        self.o_proj = nn.Linear(self.nh * self.hd, H, bias=False, dtype=dtype)
        self.gate_proj, self.up_proj = nn.Linear(H, I, bias=False, dtype=dtype), nn.Linear(H, I, bias=False, dtype=dtype)
        self.down_proj = nn.Linear(I, H, bias=False, dtype=dtype)
        self.attention_norm, self.ffn_norm = RMSNorm(H, 1e-5, dtype), RMSNorm(H, 1e-5, dtype)

    def forward(self, x, cos, sin):                          # x [N, H]; cos / sin [N, hd]
        N = x.shape[0]
        h = self.attention_norm(x)
        q, k, v = (p(h).view(N, self.nh, self.hd).transpose(0, 1) for p in (self.q_proj, self.k_proj, self.v_proj))
        q = q * cos + _rot_half(q) * sin
        k = k * cos + _rot_half(k) * sin
        a = F.scaled_dot_product_attention(q, k, v).transpose(0, 1).reshape(N, -1)
        x = x + self.o_proj(a)
        h = self.ffn_norm(x)
        return x + self.down_proj(F.silu(self.gate_proj(h)) * self.up_proj(h))


class PixtralTower(nn.Module):
    def __init__(self, vc: dict, dtype) -> None:
        super().__init__()
        H, self.P = vc["hidden_size"], vc["patch_size"]
        self.patch_conv = nn.Conv2d(vc.get("num_channels", 3), H, self.P, self.P, bias=False, dtype=dtype)
        self.ln_pre = RMSNorm(H, 1e-5, dtype)
        self.layers = nn.ModuleList([PixtralLayer(vc, dtype) for _ in range(vc["num_hidden_layers"])])
        hd = vc.get("head_dim") or H // vc["num_attention_heads"]
        side = vc["image_size"] // self.P
        base = (vc.get("rope_parameters") or {}).get("rope_theta", vc.get("rope_theta", 10000.0))
        f = 1.0 / (base ** (torch.arange(0, hd, 2).float() / hd))
        fh, fw = torch.outer(torch.arange(side).float(), f[::2]), torch.outer(torch.arange(side).float(), f[1::2])
        table = torch.cat([fh[:, None, :].expand(side, side, -1), fw[None, :, :].expand(side, side, -1)], dim=-1)   # [side, side, hd/2]
        self.register_buffer("freqs", torch.cat((table, table), dim=-1), persistent=False)

    def forward(self, pixels: torch.Tensor) -> torch.Tensor:
        """pixels [3, Hpx, Wpx] → [rows · cols, H] (row-major)."""
        x = self.patch_conv(pixels.to(self.patch_conv.weight.dtype).unsqueeze(0))[0]          # [H, rows, cols]
        rows, cols = x.shape[-2:]
        x = self.ln_pre(x.flatten(1).t().contiguous())
        fr = self.freqs[:rows, :cols].reshape(rows * cols, -1).to(x.device)
        cos, sin = fr.cos().to(x.dtype), fr.sin().to(x.dtype)
        for layer in self.layers:
            x = layer(x, cos, sin)
        return x, rows, cols


class Mistral3ForConditionalGeneration(nn.Module):
    supports_multimodal = True

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        from vllm_b200.models.registry import text_decoder_cls
        mc = vllm_config.model_config
        cfg, dtype = mc.hf_config, mc.torch_dtype
        self.cfg = cfg
        vc = cfg["vision_config"]
        self.image_token = cfg.get("image_token_index", cfg.get("image_token_id"))
        self.merge = int(cfg.get("spatial_merge_size", 2))
        self.vision_tower = PixtralTower(vc, dtype)
        Hv, Ht = vc["hidden_size"], cfg["hidden_size"]
        bias = bool(cfg.get("multimodal_projector_bias", False))
        self.norm = RMSNorm(Hv, cfg.get("rms_norm_eps", 1e-5), dtype)
        self.merging_layer = nn.Linear(Hv * self.merge ** 2, Hv, bias=False, dtype=dtype)
        self.linear_1 = nn.Linear(Hv, Ht, bias=bias, dtype=dtype)
        self.linear_2 = nn.Linear(Ht, Ht, bias=bias, dtype=dtype)
        self.language_model = text_decoder_cls(cfg, "MistralForCausalLM")(vllm_config, prefix=f"{prefix}language_model.")

    @torch.inference_mode()
    def embed_multimodal(self, pixels: torch.Tensor) -> torch.Tensor:
        outs = []
        for px in pixels:
            x, rows, cols = self.vision_tower(px)
            x = self.norm(x)
            grid = x.view(rows, cols, -1).permute(2, 0, 1).unsqueeze(0)
            x = F.unfold(grid, kernel_size=self.merge, stride=self.merge)[0].t()             # [(rows/m)(cols/m), Hv·m²] channel-major
            outs.append(self.linear_2(F.gelu(self.linear_1(self.merging_layer(x)))))
        return torch.stack(outs)

    def embed_input_ids(self, input_ids):
        return self.language_model.embed_input_ids(input_ids)

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        return self.language_model(input_ids, positions, intermediate_tensors, inputs_embeds)

    def compute_logits(self, hidden):
        return self.language_model.compute_logits(hidden)

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        own = dict(self.named_parameters())
        text, loaded = [], set()
        for name, w in weights:
            n = name[len("model."):] if name.startswith("model.") else name
            if n.startswith("language_model."):
                rest = n[len("language_model."):]
                rest = rest[len("model."):] if rest.startswith("model.") else rest
                text.append((rest if rest.startswith("lm_head") else "model." + rest, w))
            elif n.startswith("lm_head."):
                text.append((n, w))
            else:
                t = (n.replace("vision_tower.transformer.layers.", "vision_tower.layers.").replace(".attention.", ".")
                     .replace(".feed_forward.", ".").replace("multi_modal_projector.patch_merger.", "").replace("multi_modal_projector.", ""))
                if t in own:
                    own[t].data.copy_(w)
                    loaded.add(t)
        loaded |= {"language_model." + x for x in self.language_model.load_weights(text)}
        return loaded


class PixtralLayoutProcessor:
    """One placeholder → ``rows × ([IMG] × cols + [IMG_BREAK])`` with the final break replaced by ``[IMG_END]``; every row of
    ``[IMG]`` tokens is a feature record pointing at its slice of the same tower output."""

    def __init__(self, model_config) -> None:
        from collections import OrderedDict
        cfg = model_config.hf_config
        vc = cfg["vision_config"]
        self.image_size, self.patch = vc.get("image_size", 1024), vc.get("patch_size", 14)
        self.image_token = cfg.get("image_token_index", cfg.get("image_token_id"))
        self.break_token, self.end_token = cfg.get("image_break_token_id", 12), cfg.get("image_end_token_id", 13)
        self.rows_layout = bool(cfg.get("image_row_breaks", True)) and self.break_token < cfg["vocab_size"] and self.end_token < cfg["vocab_size"]
        self.side = self.image_size // self.patch // int(cfg.get("spatial_merge_size", 2))
        self.cache: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        self.cache_items = 256

    def process(self, ids: list[int], mm_data: dict[str, Any]) -> tuple[list[int], list[dict[str, Any]]]:
        from vllm_b200.multimodal import hash_item, preprocess_image
        images = mm_data.get("image")
        if images is None:
            return ids, []
        if not isinstance(images, (list, tuple)):
            images = [images]
        slots = [i for i, t in enumerate(ids) if t == self.image_token]
        if len(slots) != len(images):
            raise ValueError(f"{len(images)} images but {len(slots)} image placeholders in the prompt")
        out, feats, prev = [], [], 0
        for slot, img in zip(slots, images):
            px = preprocess_image(img, self.image_size)
            key = hash_item(px.numpy())
            out.extend(ids[prev:slot])
            if not self.rows_layout:
                feats.append({"modality": "image", "hash": key, "offset": len(out), "length": self.side ** 2, "data": px})
                out.extend([self.image_token] * self.side ** 2)
            else:
                for r in range(self.side):
                    feats.append({"modality": "image", "hash": key, "offset": len(out), "length": self.side, "data": px,
                                  "enc_offset": r * self.side})
                    out.extend([self.image_token] * self.side)
                    out.append(self.end_token if r == self.side - 1 else self.break_token)
            prev = slot + 1
        out.extend(ids[prev:])
        return out, feats
