"""InternVL (HF-native ``InternVLForConditionalGeneration``; InternVL3 / 3.5): InternViT tower → drop [CLS] → pixel shuffle
(2×2 patches per token) → LayerNorm + 2-layer MLP projector → Qwen2 / Qwen3 / Llama / GPT-OSS decoder.

InternViT is a BEiT-style pre-norm ViT: class token, learned absolute positions, per-channel residual scales
(``lambda_1`` / ``lambda_2``), optional RMSNorm on the whole q / k projections, no final norm when mean pooling is on.
Reference parity: vllm/model_executor/models/interns1.py / internvl.py, intern_vit.py, HF ``modeling_internvl``.
One image = one tile of ``image_size``; the dynamic-tiling of the HF processor yields extra tiles = extra images.
"""
from __future__ import annotations

from typing import Iterable

import torch
import torch.nn as nn
import torch.nn.functional as F

from vllm_b200.config import VllmConfig
from vllm_b200.models.idefics3 import pixel_shuffle


def _norm(kind: str, H: int, eps: float, dtype) -> nn.Module:
    if kind == "rms_norm":
        from vllm_b200.layers.common import RMSNorm
        return RMSNorm(H, eps, dtype)
    return nn.LayerNorm(H, eps, dtype=dtype)


class InternViTLayer(nn.Module):
    def __init__(self, vc: dict, dtype) -> None:
        super().__init__()
        H, self.nh = vc["hidden_size"], vc["num_attention_heads"]
        eps, kind, bias = vc.get("layer_norm_eps", 1e-6), vc.get("norm_type", "layer_norm"), bool(vc.get("attention_bias", False))
        self.q_proj, self.k_proj, self.v_proj = (nn.Linear(H, H, bias=bias, dtype=dtype) for _ in range(3))
        self.projection_layer = nn.Linear(H, H, dtype=dtype)
        self.q_norm = _norm("rms_norm", H, eps, dtype) if vc.get("use_qk_norm") else None
        self.k_norm = _norm("rms_norm", H, eps, dtype) if vc.get("use_qk_norm") else None
        self.fc1, self.fc2 = nn.Linear(H, vc["intermediate_size"], dtype=dtype), nn.Linear(vc["intermediate_size"], H, dtype=dtype)
        self.layernorm_before, self.layernorm_after = _norm(kind, H, eps, dtype), _norm(kind, H, eps, dtype)
        self.lambda_1 = nn.Parameter(torch.ones(H, dtype=dtype), requires_grad=False)
        self.lambda_2 = nn.Parameter(torch.ones(H, dtype=dtype), requires_grad=False)
        self.act = F.gelu if vc.get("hidden_act", "gelu") == "gelu" else (lambda x: F.gelu(x, approximate="tanh"))

    def _n(self, norm, x):
        return norm(x.reshape(-1, x.shape[-1])).view(x.shape) if norm is not None else x

    def forward(self, x):
        B, N, H = x.shape
        h = self._n(self.layernorm_before, x)
        q, k, v = self._n(self.q_norm, self.q_proj(h)), self._n(self.k_norm, self.k_proj(h)), self.v_proj(h)
        q, k, v = (t.view(B, N, self.nh, H // self.nh).transpose(1, 2) for t in (q, k, v))
        a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, H)
        x = x + self.lambda_1 * self.projection_layer(a)
        return x + self.lambda_2 * self.fc2(self.act(self.fc1(self._n(self.layernorm_after, x))))


class InternViT(nn.Module):
    def __init__(self, vc: dict, dtype) -> None:
        super().__init__()
        H = vc["hidden_size"]
        P = vc["patch_size"][0] if isinstance(vc["patch_size"], (list, tuple)) else vc["patch_size"]
        S = vc["image_size"][0] if isinstance(vc["image_size"], (list, tuple)) else vc["image_size"]
        self.patch = nn.Conv2d(vc.get("num_channels", 3), H, P, P, dtype=dtype)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, H, dtype=dtype), requires_grad=False)
        self.position_embeddings = (nn.Parameter(torch.zeros(1, (S // P) ** 2 + 1, H, dtype=dtype), requires_grad=False)
                                    if vc.get("use_absolute_position_embeddings", True) else None)
        self.layers = nn.ModuleList([InternViTLayer(vc, dtype) for _ in range(vc["num_hidden_layers"])])
        self.final = None if vc.get("use_mean_pooling", True) else nn.LayerNorm(H, vc.get("layer_norm_eps", 1e-6), dtype=dtype)

    def forward(self, pixels):
        x = self.patch(pixels.to(self.patch.weight.dtype)).flatten(2).transpose(1, 2)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1)
        if self.position_embeddings is not None:
            x = x + self.position_embeddings[:, : x.shape[1]]
        for layer in self.layers:
            x = layer(x)
        return self.final(x) if self.final is not None else x


class InternVLForConditionalGeneration(nn.Module):
    supports_multimodal = True

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        from vllm_b200.models.registry import text_decoder_cls
        mc = vllm_config.model_config
        cfg, dtype = mc.hf_config, mc.torch_dtype
        self.cfg = cfg
        self.image_token = cfg.get("image_token_id", cfg.get("image_token_index"))
        vc = cfg["vision_config"]
        self.ratio = float(cfg.get("downsample_ratio", 0.5))
        self.select = cfg.get("vision_feature_select_strategy", "default")
        self.vision_tower = InternViT(vc, dtype)
        Hin = vc["hidden_size"] * int(1 / self.ratio) ** 2
        self.layer_norm = nn.LayerNorm(Hin, dtype=dtype)
        self.linear_1 = nn.Linear(Hin, cfg["hidden_size"], dtype=dtype)
        self.linear_2 = nn.Linear(cfg["hidden_size"], cfg["hidden_size"], dtype=dtype)
        if cfg.get("text_model_type") == "qwen2":
            cfg.setdefault("attention_bias", True)
        self.language_model = text_decoder_cls(cfg, "Qwen2ForCausalLM")(vllm_config, prefix=f"{prefix}language_model.")

    @torch.inference_mode()
    def embed_multimodal(self, pixels: torch.Tensor) -> torch.Tensor:
        x = self.vision_tower(pixels)
        if self.select == "default":
            x = x[:, 1:]
        x = pixel_shuffle(x, int(1 / self.ratio))
        return self.linear_2(F.gelu(self.linear_1(self.layer_norm(x))))

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
            elif n.startswith("vision_tower."):
                t = (n.replace("embeddings.patch_embeddings.projection.", "patch.").replace("embeddings.", "")
                     .replace("encoder.layer.", "layers.").replace(".attention.", ".").replace(".mlp.", ".")
                     .replace("vision_tower.layernorm.", "vision_tower.final."))
                if t in own:
                    own[t].data.copy_(w)
                    loaded.add(t)
            elif n.startswith("multi_modal_projector."):
                t = n[len("multi_modal_projector."):]
                if t in own:
                    own[t].data.copy_(w)
                    loaded.add(t)
        loaded |= {"language_model." + x for x in self.language_model.load_weights(text)}
        return loaded
