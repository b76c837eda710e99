"""Llama-4 vision-language (Scout / Maverick): ViT over unfolded patches (bias-free linear patch projection, class token
APPENDED, learned positions, pre / post LayerNorm) with 2-D rotary angles — half of each head's frequency pairs follow the
patch column, half the row, the class token is not rotated — → drop the class token → pixel shuffle (ratio 0.5: 2×2 patches
per token) → 2-layer GELU MLP → one linear projector → the Llama-4 MoE decoder.

Reference parity: vllm/model_executor/models/mllama4.py, HF ``modeling_llama4`` (``Llama4VisionModel``).
One image = one ``image_size`` tile here (the HF processor's tiling yields extra tiles = extra images + separator tokens).
"""
from __future__ import annotations

from typing import Iterable

import torch
import torch.nn as nn
import torch.nn.functional as F

from vllm_b200.config import VllmConfig
from vllm_b200.models.idefics3 import pixel_shuffle
from vllm_b200.models.llama4 import Llama4ForCausalLM


class _Layer(nn.Module):
    def __init__(self, H, I, nh, dtype) -> None:
        super().__init__()
        self.nh = nh
        self.q_proj, self.k_proj, self.v_proj, self.o_proj = (nn.Linear(H, H, dtype=dtype) for _ in range(4))
        self.fc1, self.fc2 = nn.Linear(H, I, dtype=dtype), nn.Linear(I, H, dtype=dtype)
        self.input_layernorm, self.post_attention_layernorm = nn.LayerNorm(H, dtype=dtype), nn.LayerNorm(H, dtype=dtype)

    @staticmethod
    def _rot(x, cos, sin):                                      # x [B, N, nh, hd]; adjacent pairs rotate together
        a, b = x.float()[..., 0::2], x.float()[..., 1::2]
        return torch.stack((a * cos - b * sin, a * sin + b * cos), dim=-1).flatten(-2).to(x.dtype)

    def forward(self, x, cos, sin):
        B, N, H = x.shape
        h = self.input_layernorm(x)
        q, k, v = (p(h).view(B, N, self.nh, H // self.nh) for p in (self.q_proj, self.k_proj, self.v_proj))
        q, k = self._rot(q, cos, sin), self._rot(k, cos, sin)
        a = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2).reshape(B, N, H)
        x = x + self.o_proj(a)
        return x + self.fc2(F.gelu(self.fc1(self.post_attention_layernorm(x))))


class Llama4VisionTower(nn.Module):
    def __init__(self, vc: dict, dtype) -> None:
        super().__init__()
        H, self.P, S = vc["hidden_size"], vc["patch_size"], vc["image_size"]
        side = S // self.P
        self.patch_linear = nn.Linear(vc.get("num_channels", 3) * self.P ** 2, H, bias=False, dtype=dtype)
        self.class_embedding = nn.Parameter(torch.zeros(H, dtype=dtype), requires_grad=False)
        self.positional_embedding_vlm = nn.Parameter(torch.zeros(side ** 2 + 1, H, dtype=dtype), requires_grad=False)
        self.layernorm_pre, self.layernorm_post = nn.LayerNorm(H, dtype=dtype), nn.LayerNorm(H, dtype=dtype)
        nh = vc["num_attention_heads"]
        self.layers = nn.ModuleList([_Layer(H, vc["intermediate_size"], nh, dtype) for _ in range(vc["num_hidden_layers"])])
        self.ratio = float(vc.get("pixel_shuffle_ratio", 0.5))
        self.adapter_fc1 = nn.Linear(vc["intermediate_size"], vc["projector_input_dim"], bias=False, dtype=dtype)
        self.adapter_fc2 = nn.Linear(vc["projector_output_dim"], vc["projector_output_dim"], bias=False, dtype=dtype)
        hd = H // nh
        fd = hd // 2
        theta = (vc.get("rope_parameters") or {}).get("rope_theta", vc.get("rope_theta", 10000.0))
        rf = 1.0 / (theta ** (torch.arange(0, fd, 2)[: fd // 2].float() / fd))                 # hd / 4 frequencies
        idx = torch.arange(side ** 2)
        ang = torch.cat([((idx % side) + 1)[:, None] * rf[None], ((idx // side) + 1)[:, None] * rf[None]], dim=-1)   # [N, hd/2]
        ang = torch.cat([ang, torch.zeros(1, ang.shape[1])], dim=0)                              # class token: no rotation
        self.register_buffer("cos", ang.cos()[None, :, None, :], persistent=False)
        self.register_buffer("sin", ang.sin()[None, :, None, :], persistent=False)

    def forward(self, pixels: torch.Tensor) -> torch.Tensor:
        x = F.unfold(pixels.to(self.patch_linear.weight.dtype), self.P, stride=self.P).transpose(1, 2)
        x = self.patch_linear(x)
        x = torch.cat([x, self.class_embedding.expand(x.shape[0], 1, -1)], dim=1) + self.positional_embedding_vlm
        x = self.layernorm_pre(x)
        for layer in self.layers:
            x = layer(x, self.cos, self.sin)
        x = self.layernorm_post(x)[:, :-1]
        x = pixel_shuffle(x, int(round(1 / self.ratio)))
        return F.gelu(self.adapter_fc2(F.gelu(self.adapter_fc1(x))))


class Llama4ForConditionalGeneration(nn.Module):
    supports_multimodal = True

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        mc = vllm_config.model_config
        cfg, dtype = mc.hf_config, mc.torch_dtype
        self.cfg = cfg
        self.image_token = cfg.get("image_token_index", cfg.get("image_token_id"))
        vc = cfg["vision_config"]
        self.vision_model = Llama4VisionTower(vc, dtype)
        self.projector = nn.Linear(vc["vision_output_dim"], cfg["hidden_size"], bias=False, dtype=dtype)
        self.language_model = Llama4ForCausalLM(vllm_config, prefix=f"{prefix}language_model.")

    @torch.inference_mode()
    def embed_multimodal(self, pixels: torch.Tensor) -> torch.Tensor:
        return self.projector(self.vision_model(pixels))

    def embed_input_ids(self, input_ids):
        return self.language_model.embed_input_ids(input_ids)

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        return self.language_model(input_ids, positions, intermediate_tensors, inputs_embeds)

    def compute_logits(self, hidden):
        return self.language_model.compute_logits(hidden)

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        own = dict(self.named_parameters())
        text, loaded = [], set()
        ren = [("vision_model.patch_embedding.linear.", "vision_model.patch_linear."), ("vision_model.model.layers.", "vision_model.layers."),
               (".self_attn.", "."), (".mlp.fc", ".fc"), ("vision_model.vision_adapter.mlp.fc1.", "vision_model.adapter_fc1."),
               ("vision_model.vision_adapter.mlp.fc2.", "vision_model.adapter_fc2."), ("multi_modal_projector.linear_1.", "projector.")]
        for name, w in weights:
            n = name[len("model."):] if name.startswith("model.") else name
            if n.startswith("language_model.") or n.startswith("lm_head."):
                text.append((n, w))
                continue
            t = n
            if t.startswith("vision_model.vision_adapter"):
                t = t.replace("vision_model.vision_adapter.mlp.fc1.", "vision_model.adapter_fc1.").replace(
                    "vision_model.vision_adapter.mlp.fc2.", "vision_model.adapter_fc2.")
            else:
                for a, b in ren:
                    t = t.replace(a, b)
            if t in own:
                own[t].data.copy_(w)
                loaded.add(t)
        loaded |= {"language_model." + x for x in self.language_model.load_weights(text)}
        return loaded
