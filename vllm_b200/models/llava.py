"""LLaVA-1.5 style vision-language model: CLIP ViT tower → 2-layer MLP projector → Llama decoder.

Reference parity: vllm/model_executor/models/llava.py (``LlavaForConditionalGeneration``,
``_process_image_input``, placeholder merge), clip.py (``CLIPVisionModel``), interfaces.py
``SupportsMultiModal`` (``embed_multimodal`` / ``embed_input_ids``).

The tower is replicated on every TP rank (it is ~0.3 B parameters next to a 7-70 B decoder and runs once per
image); its attention is a dense bidirectional SDPA over a few hundred patches.
"""
from __future__ import annotations

import copy
from typing import Iterable, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from vllm_b200.config import VllmConfig
from vllm_b200.models.llama import LlamaForCausalLM


def _act(name: str):
    return {"quick_gelu": lambda x: x * torch.sigmoid(1.702 * x), "gelu": F.gelu,
            "gelu_pytorch_tanh": lambda x: F.gelu(x, approximate="tanh")}[name]


class CLIPLayer(nn.Module):
    def __init__(self, H: int, I: int, nh: int, eps: float, act: str, dtype) -> None:
        super().__init__()
        self.nh = nh
        self.layer_norm1 = nn.LayerNorm(H, eps, dtype=dtype)
        self.q_proj = nn.Linear(H, H, dtype=dtype)
        self.k_proj = nn.Linear(H, H, dtype=dtype)
        self.v_proj = nn.Linear(H, H, dtype=dtype)
        self.out_proj = nn.Linear(H, H, dtype=dtype)
        self.layer_norm2 = nn.LayerNorm(H, eps, dtype=dtype)
        self.fc1 = nn.Linear(H, I, dtype=dtype)
        self.fc2 = nn.Linear(I, H, dtype=dtype)
        self.act = _act(act)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, N, H = x.shape
        h = self.layer_norm1(x)
        q, k, v = (p(h).view(B, N, self.nh, H // self.nh).transpose(1, 2) for p in (self.q_proj, self.k_proj, self.v_proj))
        a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, H)
        x = x + self.out_proj(a)
        return x + self.fc2(self.act(self.fc1(self.layer_norm2(x))))


class CLIPVisionTower(nn.Module):
    def __init__(self, vc: dict, dtype, feature_layer: int) -> None:
        super().__init__()
        H, P, S = vc["hidden_size"], vc["patch_size"], vc["image_size"]
        self.patch_embedding = nn.Conv2d(3, H, P, P, bias=False, dtype=dtype)
        self.class_embedding = nn.Parameter(torch.zeros(H, dtype=dtype))
        self.position_embedding = nn.Embedding((S // P) ** 2 + 1, H, dtype=dtype)
        eps = vc.get("layer_norm_eps", 1e-5)
        self.pre_layrnorm = nn.LayerNorm(H, eps, dtype=dtype)
        n = vc["num_hidden_layers"]
        # hidden_states[feature_layer]: index 0 is the embedding output → run this many layers
        keep = n + 1 + feature_layer if feature_layer < 0 else feature_layer
        self.layers = nn.ModuleList([CLIPLayer(H, vc["intermediate_size"], vc["num_attention_heads"], eps,
                                               vc.get("hidden_act", "quick_gelu"), dtype) for _ in range(keep)])

    def forward(self, pixels: torch.Tensor) -> torch.Tensor:
        x = self.patch_embedding(pixels.to(self.patch_embedding.weight.dtype)).flatten(2).transpose(1, 2)
        cls = self.class_embedding.expand(x.shape[0], 1, -1)
        x = torch.cat([cls, x], dim=1) + self.position_embedding.weight[None]
        x = self.pre_layrnorm(x)
        for layer in self.layers:
            x = layer(x)
        return x


class LlavaForConditionalGeneration(nn.Module):
    supports_multimodal = True

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        mc = vllm_config.model_config
        cfg = mc.hf_config
        self.cfg = cfg
        dtype = mc.torch_dtype
        self.image_token = cfg.get("image_token_index", cfg.get("image_token_id"))
        self.vision_tower = self._build_tower(cfg, dtype)
        self.select = cfg.get("vision_feature_select_strategy", "default")
        Hv, Ht = cfg["vision_config"]["hidden_size"], cfg["hidden_size"]
        pb = bool(cfg.get("multimodal_projector_bias", True))
        self.linear_1 = nn.Linear(Hv, Ht, bias=pb, dtype=dtype)
        self.linear_2 = nn.Linear(Ht, Ht, bias=pb, dtype=dtype)
        self.proj_act = _act(cfg.get("projector_hidden_act", "gelu"))
        # the text stack reads the flattened text fields of the same config
        from vllm_b200.models.registry import text_decoder_cls
        if cfg.get("text_model_type") == "qwen2":
            cfg.setdefault("attention_bias", True)
        self.language_model = text_decoder_cls(cfg, "LlamaForCausalLM")(vllm_config, prefix=f"{prefix}language_model.")

    def _build_tower(self, cfg, dtype) -> nn.Module:
        if cfg["vision_config"].get("model_type") == "pixtral":
            # pixtral-12b in the HF layout: Pixtral tower (2-D rotary, no class token) behind the LLaVA projector
            from vllm_b200.models.mistral3 import PixtralTower
            tower = PixtralTower(cfg["vision_config"], dtype)
            inner = tower.forward
            tower.forward = lambda pixels: torch.stack([inner(px)[0] for px in pixels])
            self._tower_rename = lambda n: (n.replace("transformer.layers.", "layers.").replace(".attention.", ".")
                                            .replace(".feed_forward.", "."))
            cfg["vision_feature_select_strategy"] = "full"
            return tower
        return CLIPVisionTower(cfg["vision_config"], dtype, cfg.get("vision_feature_layer", -2))

    _tower_rename = None           # subclasses with another tower family: checkpoint name → tower parameter name

    # ---------------------------------------------------------------- multimodal
    @torch.inference_mode()
    def embed_multimodal(self, pixels: torch.Tensor) -> torch.Tensor:
        """pixels [n, 3, S, S] → [n, tokens_per_image, H_text]"""
        feats = self.vision_tower(pixels)
        if self.select == "default":
            feats = feats[:, 1:]
        return self.linear_2(self.proj_act(self.linear_1(feats)))

    def embed_input_ids(self, input_ids: torch.Tensor) -> torch.Tensor:
        return self.language_model.embed_input_ids(input_ids)

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        return self.language_model(input_ids, positions, intermediate_tensors, inputs_embeds)

    def compute_logits(self, hidden: torch.Tensor) -> torch.Tensor:
        return self.language_model.compute_logits(hidden)

    # ---------------------------------------------------------------- weights
    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        own = dict(self.named_parameters())
        text: list[tuple[str, torch.Tensor]] = []
        loaded: set[str] = set()
        for name, w in weights:
            n = name[len("model."):] if name.startswith("model.") else name
            if n.startswith("language_model."):
                rest = n[len("language_model."):]
                rest = rest[len("model."):] if rest.startswith("model.") else rest
                text.append((rest if rest.startswith("lm_head") else "model." + rest, w))
            elif n.startswith("lm_head."):
                text.append((n, w))
            elif n.startswith("vision_tower."):
                if self._tower_rename is not None:
                    t = "vision_tower." + self._tower_rename(n[len("vision_tower."):].removeprefix("vision_model."))
                else:
                    t = (n.replace("vision_tower.vision_model.", "vision_tower.").replace("embeddings.", "")
                         .replace("encoder.layers.", "layers.").replace("self_attn.", "").replace("mlp.", ""))
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
