"""Idefics3 / SmolVLM: SigLIP vision tower → pixel-shuffle connector (s×s neighbouring patches concatenated channel-wise,
then one bias-free linear) → Llama decoder.

Reference parity: vllm/model_executor/models/idefics3.py, smolvlm.py, siglip.py (``SiglipVisionModel``), HF
``modeling_smolvlm`` (``SmolVLMConnector.pixel_shuffle``).  One image → one ``image_size``-square tile (the front-end
resizes; sub-image splitting of the HF processor is a preprocessing choice, each extra tile is just another image here).
The tower is replicated on every TP rank, like the CLIP tower of LLaVA.
"""
from __future__ import annotations

from typing import Iterable

import torch
import torch.nn as nn

from vllm_b200.config import VllmConfig
from vllm_b200.models.llama import LlamaForCausalLM
from vllm_b200.models.llava import CLIPLayer


class SiglipVisionTower(nn.Module):
    """SigLIP ViT: biased patch convolution, learned positions for the patch grid (no class token), pre-norm encoder
    layers with tanh-GELU MLPs, final LayerNorm."""

    def __init__(self, vc: dict, dtype) -> None:
        super().__init__()
        H, P, S = vc["hidden_size"], vc["patch_size"], vc["image_size"]
        self.patch_embedding = nn.Conv2d(vc.get("num_channels", 3), H, P, P, bias=True, dtype=dtype)
        self.position_embedding = nn.Embedding((S // P) ** 2, H, dtype=dtype)
        eps = vc.get("layer_norm_eps", 1e-6)
        self.layers = nn.ModuleList([CLIPLayer(H, vc["intermediate_size"], vc["num_attention_heads"], eps,
                                               vc.get("hidden_act", "gelu_pytorch_tanh"), dtype)
                                     for _ in range(vc["num_hidden_layers"])])
        self.post_layernorm = nn.LayerNorm(H, eps, dtype=dtype)

    def forward(self, pixels: torch.Tensor, final_norm: bool = True) -> torch.Tensor:
        x = self.patch_embedding(pixels.to(self.patch_embedding.weight.dtype)).flatten(2).transpose(1, 2)
        x = x + self.position_embedding.weight[None, : x.shape[1]]
        for layer in self.layers:
            x = layer(x)
        return self.post_layernorm(x) if final_norm else x

    @staticmethod
    def rename(n: str) -> str:
        return (n.replace("embeddings.", "").replace("encoder.layers.", "layers.").replace("self_attn.", "").replace("mlp.", ""))


def pixel_shuffle(x: torch.Tensor, s: int) -> torch.Tensor:
    """[B, N, C] on a √N × √N grid → [B, N / s², C · s²]: each s×s block of patches becomes one token."""
    B, N, C = x.shape
    h = w = int(N ** 0.5)
    x = x.reshape(B, h, w // s, C * s).permute(0, 2, 1, 3)
    x = x.reshape(B, w // s, h // s, C * s * s).permute(0, 2, 1, 3)
    return x.reshape(B, N // (s * s), C * s * s)


class SiglipVLM(nn.Module):
    """SigLIP tower + a family-specific projector + a registered text decoder; subclasses fill in ``_build_projector`` /
    ``_project`` and the checkpoint prefixes."""
    supports_multimodal = True
    text_cls = LlamaForCausalLM
    text_prefixes = ("text_model.", "language_model.")
    vision_prefixes = ("vision_model.", "vision_tower.vision_model.")
    projector_prefixes = ("connector.", "multi_modal_projector.")

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        mc = vllm_config.model_config
        cfg, dtype = mc.hf_config, mc.torch_dtype
        self.cfg = cfg
        self.image_token = cfg.get("image_token_id", cfg.get("image_token_index"))
        self.vision_model = SiglipVisionTower(cfg["vision_config"], dtype)
        self._build_projector(cfg, dtype)
        self.language_model = self.text_cls(vllm_config, prefix=f"{prefix}language_model.")

    def _build_projector(self, cfg, dtype) -> None:
        raise NotImplementedError

    def _project(self, feats: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def _projector_name(self, n: str) -> str:
        return n

    @torch.inference_mode()
    def embed_multimodal(self, pixels: torch.Tensor) -> torch.Tensor:
        return self._project(self.vision_model(pixels))

    def embed_input_ids(self, input_ids: torch.Tensor) -> torch.Tensor:
        return self.language_model.embed_input_ids(input_ids)

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        return self.language_model(input_ids, positions, intermediate_tensors, inputs_embeds)

    def compute_logits(self, hidden: torch.Tensor) -> torch.Tensor:
        return self.language_model.compute_logits(hidden)

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        own = dict(self.named_parameters())
        text: list[tuple[str, torch.Tensor]] = []
        loaded: set[str] = set()
        for name, w in weights:
            n = name[len("model."):] if name.startswith("model.") else name
            tp = next((p for p in self.text_prefixes if n.startswith(p)), None)
            vp = next((p for p in self.vision_prefixes if n.startswith(p)), None)
            pp = next((p for p in self.projector_prefixes if n.startswith(p)), None)
            if tp is not None:
                rest = n[len(tp):]
                rest = rest[len("model."):] if rest.startswith("model.") else rest
                text.append((rest if rest.startswith("lm_head") else "model." + rest, w))
            elif n.startswith("lm_head."):
                text.append((n, w))
            elif vp is not None:
                t = "vision_model." + SiglipVisionTower.rename(n[len(vp):])
                if t in own:
                    own[t].data.copy_(w)
                    loaded.add(t)
            elif pp is not None:
                t = self._projector_name(n[len(pp):])
                if t in own:
                    own[t].data.copy_(w)
                    loaded.add(t)
        loaded |= {"language_model." + x for x in self.language_model.load_weights(text)}
        return loaded


class Idefics3ForConditionalGeneration(SiglipVLM):
    def _build_projector(self, cfg, dtype) -> None:
        self.scale = int(cfg.get("scale_factor", 2))
        self.proj = nn.Linear(cfg["vision_config"]["hidden_size"] * self.scale ** 2, cfg["hidden_size"], bias=False, dtype=dtype)

    def _project(self, feats):
        return self.proj(pixel_shuffle(feats, self.scale))

    def _projector_name(self, n):
        return n.replace("modality_projection.proj.", "proj.")


class SmolVLMForConditionalGeneration(Idefics3ForConditionalGeneration):
    pass


class Gemma3ForConditionalGeneration(SiglipVLM):
    """Gemma-3 vision-language: SigLIP tower → average-pool the patch grid down to ``mm_tokens_per_image`` → (1+w) RMSNorm →
    one projection matrix → Gemma-3 decoder (reference: models/gemma3_mm.py).  Image positions attend causally here; the
    bidirectional image-block mask of the HF implementation (``token_type_ids``) is not applied."""

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        from vllm_b200.models.llama_variants import Gemma3ForCausalLM
        self.text_cls = Gemma3ForCausalLM
        super().__init__(vllm_config, prefix)

    def _build_projector(self, cfg, dtype) -> None:
        from vllm_b200.layers.common import RMSNorm
        vc = cfg["vision_config"]
        self.mm_input_projection_weight = nn.Parameter(torch.zeros(vc["hidden_size"], cfg["hidden_size"], dtype=dtype), requires_grad=False)
        self.mm_soft_emb_norm = RMSNorm(vc["hidden_size"], vc.get("layer_norm_eps", 1e-6), dtype, True)
        self.side = vc["image_size"] // vc["patch_size"]
        self.kernel = self.side // int(int(cfg.get("mm_tokens_per_image", 256)) ** 0.5)

    def _project(self, feats):
        B, _, C = feats.shape
        x = feats.transpose(1, 2).reshape(B, C, self.side, self.side)
        x = torch.nn.functional.avg_pool2d(x, self.kernel, self.kernel).flatten(2).transpose(1, 2)
        x = self.mm_soft_emb_norm(x.reshape(-1, C)).view(B, -1, C)
        return torch.matmul(x, self.mm_input_projection_weight)



class Cohere2VisionForConditionalGeneration(SiglipVLM):
    """Command-A Vision: SigLIP tower → pixel shuffle (f×f patches per token) → SwiGLU projector → Cohere-2 decoder
    (reference: models/cohere2_vision.py)."""
    final_norm = True
    pre_ln = False

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        from vllm_b200.models.llama_variants import Cohere2ForCausalLM
        self.text_cls = Cohere2ForCausalLM
        super().__init__(vllm_config, prefix)

    def _build_projector(self, cfg, dtype) -> None:
        self.f = int(cfg.get("downsample_factor", 2))
        Hin = cfg["vision_config"]["hidden_size"] * self.f ** 2
        mid = int(cfg.get("alignment_intermediate_size") or cfg["hidden_size"])
        if self.pre_ln:
            self.layernorm = nn.LayerNorm(Hin, cfg.get("adapter_layer_norm_eps", 1e-6), dtype=dtype)
        self.linear_1 = nn.Linear(Hin, mid, dtype=dtype)
        self.linear_2 = nn.Linear(mid // 2, cfg["hidden_size"], dtype=dtype)

    @torch.inference_mode()
    def embed_multimodal(self, pixels: torch.Tensor) -> torch.Tensor:
        return self._project(self.vision_model(pixels, final_norm=self.final_norm))

    def _project(self, feats):
        x = pixel_shuffle(feats, self.f)
        if self.pre_ln:
            x = self.layernorm(x)
        x, gate = self.linear_1(x).chunk(2, dim=-1)
        return self.linear_2(torch.nn.functional.silu(gate) * x)


class AyaVisionForConditionalGeneration(Cohere2VisionForConditionalGeneration):
    """Aya Vision: the same connector with a LayerNorm in front, fed from the last encoder layer BEFORE the tower's final
    LayerNorm (``vision_feature_layer = -1``) (reference: models/aya_vision.py)."""
    final_norm = False
    pre_ln = True


class PaliGemmaForConditionalGeneration(SiglipVLM):
    """PaliGemma / PaliGemma-2: SigLIP tower → one linear projection → Gemma (or Gemma-2) decoder run as a PREFIX-LM: image
    and prompt tokens attend to each other in both directions, generated tokens causally.  The attention layers are built
    non-causal; a decode step has one query at the end of its sequence, for which that is the causal result, and the
    engine keeps the prompt in a single prefill pass (no chunked prefill, no prefix reuse — ``ModelConfig.prefix_lm``).
    The ``1/sqrt(hidden)`` HF applies to the projected image features only cancels Gemma's embedding scale, which here
    never touches spliced features (reference: models/paligemma.py)."""

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        from vllm_b200.models.registry import text_decoder_cls
        cfg = vllm_config.model_config.hf_config
        cfg["is_causal"] = False
        cfg.setdefault("tie_word_embeddings", True)
        cfg["model_type"] = cfg.get("text_model_type") or "gemma"          # the decoder keys its Gemma-isms on this
        self.text_cls = text_decoder_cls(cfg, "GemmaForCausalLM")
        super().__init__(vllm_config, prefix)

    def _build_projector(self, cfg, dtype) -> None:
        self.linear = nn.Linear(cfg["vision_config"]["hidden_size"], cfg["hidden_size"], dtype=dtype)

    def _project(self, feats):
        return self.linear(feats)
