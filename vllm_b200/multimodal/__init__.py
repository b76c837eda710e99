"""Multimodal inputs (reference: vllm/multimodal/* — ``MULTIMODAL_REGISTRY``, processors, hashing, processing
cache; vllm/v1/core/encoder_cache_manager.py).

Flow: the front-end ``MultiModalProcessor`` turns raw items (images) into model-ready tensors, expands each
placeholder token of the prompt into as many positions as the encoder produces, and attaches one feature
record per item ``{modality, hash, offset, length, data}`` to the request.  The content hash makes the item
part of the prefix-cache key (``core/block_pool._extra_keys``) and is the key of the worker-side encoder-output
cache, so the same image in many requests is encoded once.  The worker runs the vision tower when the item's
positions are first scheduled and splices the projected embeddings into ``inputs_embeds``.
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict
from typing import Any, Optional

import numpy as np
import torch

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def hash_item(arr: np.ndarray) -> str:
    h = hashlib.blake2b(digest_size=16)
    h.update(str(arr.shape).encode())
    h.update(str(arr.dtype).encode())
    h.update(np.ascontiguousarray(arr).tobytes())
    return h.hexdigest()


def preprocess_image(img: Any, size: int, mean=CLIP_MEAN, std=CLIP_STD) -> torch.Tensor:
    """PIL image / HWC uint8 array / CHW float tensor → normalised float32 [3, size, size]."""
    if isinstance(img, torch.Tensor) and img.dim() == 3 and img.shape[0] == 3 and img.is_floating_point():
        t = img.float()
        if t.shape[1] != size or t.shape[2] != size:
            t = torch.nn.functional.interpolate(t[None], size=(size, size), mode="bicubic", align_corners=False)[0]
        return t                                    # caller supplied normalised pixels
    if hasattr(img, "convert"):                     # PIL
        img = np.array(img.convert("RGB"))
    a = torch.as_tensor(np.array(img))
    if a.dim() == 3 and a.shape[-1] == 3:
        a = a.permute(2, 0, 1)
    t = a.float() / (255.0 if a.dtype == torch.uint8 else 1.0)
    t = torch.nn.functional.interpolate(t[None], size=(size, size), mode="bicubic", align_corners=False)[0].clamp(0, 1)
    m = torch.tensor(mean)[:, None, None]
    s = torch.tensor(std)[:, None, None]
    return (t - m) / s


class MultiModalProcessor:
    """Image processor for CLIP-ViT style towers (LLaVA family)."""

    def __init__(self, model_config) -> None:
        cfg = model_config.hf_config
        vc = cfg["vision_config"]
        self.image_size = vc.get("image_size", 336)
        self.patch = vc.get("patch_size", 14)
        if isinstance(self.image_size, (list, tuple)):
            self.image_size = self.image_size[0]
        if isinstance(self.patch, (list, tuple)):
            self.patch = self.patch[0]
        self.image_token = cfg.get("image_token_index", cfg.get("image_token_id"))
        n = (self.image_size // self.patch) ** 2
        self.tokens_per_image = n if cfg.get("vision_feature_select_strategy", "default") == "default" else n + 1
        if vc.get("pixel_shuffle_ratio"):                   # Llama-4: ratio 0.5 → 2 x 2 patches per token
            self.tokens_per_image = int(n * float(vc["pixel_shuffle_ratio"]) ** 2)
        elif cfg.get("downsample_factor"):                    # Aya Vision / Command-A Vision
            self.tokens_per_image = n // int(cfg["downsample_factor"]) ** 2
        elif cfg.get("scale_factor"):                         # pixel-shuffle connectors (Idefics3 / SmolVLM): s x s patches per token
            self.tokens_per_image = n // int(cfg["scale_factor"]) ** 2
        elif cfg.get("downsample_ratio"):                   # InternVL: ratio 0.5 → 4 patches per token
            self.tokens_per_image = int(n * float(cfg["downsample_ratio"]) ** 2)
        elif cfg.get("mm_tokens_per_image"):                # pooled projector with a fixed output length (Gemma-3)
            self.tokens_per_image = int(cfg["mm_tokens_per_image"])
        self.cache: "OrderedDict[str, torch.Tensor]" = OrderedDict()   # processing cache (hash → pixels)
        self.cache_items = 256

    def process(self, ids: list[int], mm_data: dict[str, Any]) -> tuple[list[int], list[dict[str, Any]]]:
        images = mm_data.get("image")
        if images is None:
            return ids, []
        if not isinstance(images, (list, tuple)):
            images = [images]
        slots = [i for i, t in enumerate(ids) if t == self.image_token]
        if len(slots) != len(images):
            raise ValueError(f"{len(images)} images but {len(slots)} image placeholders in the prompt")
        out: list[int] = []
        feats = []
        prev = 0
        for slot, img in zip(slots, images):
            px = preprocess_image(img, self.image_size)
            key = hash_item(px.numpy())
            self.cache[key] = px
            while len(self.cache) > self.cache_items:
                self.cache.popitem(last=False)
            out.extend(ids[prev:slot])
            feats.append({"modality": "image", "hash": key, "offset": len(out), "length": self.tokens_per_image,
                          "data": px})
            out.extend([self.image_token] * self.tokens_per_image)
            prev = slot + 1
        out.extend(ids[prev:])
        return out, feats


class EncoderCache:
    """Worker-side cache of encoder outputs keyed by item hash (LRU, bounded by embedding rows)."""

    def __init__(self, max_rows: int = 1 << 16) -> None:
        self.max_rows = max_rows
        self.rows = 0
        self.od: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        self.hits = self.misses = 0

    def get(self, key: str) -> Optional[torch.Tensor]:
        t = self.od.get(key)
        if t is None:
            self.misses += 1
            return None
        self.od.move_to_end(key)
        self.hits += 1
        return t

    def put(self, key: str, t: torch.Tensor) -> None:
        self.od[key] = t
        self.rows += t.shape[0]
        while self.rows > self.max_rows and len(self.od) > 1:
            _, old = self.od.popitem(last=False)
            self.rows -= old.shape[0]


class _Registry:
    def __init__(self) -> None:
        self.processors: dict[str, type] = {"LlavaForConditionalGeneration": MultiModalProcessor}

    def register_processor(self, arch: str, cls: type) -> None:
        self.processors[arch] = cls

    def create_processor(self, model_config):
        if (model_config.architecture in ("Qwen2VLForConditionalGeneration", "Qwen2_5_VLForConditionalGeneration",
                                          "Qwen3VLForConditionalGeneration", "Qwen3VLMoeForConditionalGeneration",
                                          "Qwen3_5ForConditionalGeneration", "Qwen3_5MoeForConditionalGeneration")
                and model_config.architecture not in self.processors):
            from vllm_b200.multimodal.qwen_vl import Qwen2VLProcessor
            self.processors[model_config.architecture] = Qwen2VLProcessor
        if model_config.architecture in ("Mistral3ForConditionalGeneration",) and model_config.architecture not in self.processors:
            from vllm_b200.models.mistral3 import PixtralLayoutProcessor
            self.processors[model_config.architecture] = PixtralLayoutProcessor
        if (model_config.architecture == "LlavaForConditionalGeneration"
                and (model_config.hf_config.get("vision_config") or {}).get("model_type") == "pixtral"):
            from vllm_b200.models.mistral3 import PixtralLayoutProcessor
            model_config.hf_config.setdefault("spatial_merge_size", 1)
            return PixtralLayoutProcessor(model_config)
        if model_config.architecture == "LlavaNextForConditionalGeneration" and model_config.architecture not in self.processors:
            from vllm_b200.models.llava_next import LlavaNextProcessor
            self.processors[model_config.architecture] = LlavaNextProcessor
        if model_config.architecture == "LlavaNextVideoForConditionalGeneration" and model_config.architecture not in self.processors:
            from vllm_b200.models.llava_next import LlavaNextVideoProcessor
            self.processors[model_config.architecture] = LlavaNextVideoProcessor
        if model_config.architecture == "LlavaOnevisionForConditionalGeneration" and model_config.architecture not in self.processors:
            from vllm_b200.models.llava_next import LlavaOnevisionProcessor
            self.processors[model_config.architecture] = LlavaOnevisionProcessor
        if model_config.hf_config.get("audio_config") is not None and model_config.architecture not in self.processors:
            from vllm_b200.models.audio_lm import AudioLMProcessor
            self.processors[model_config.architecture] = AudioLMProcessor
        return self.processors.get(model_config.architecture, MultiModalProcessor)(model_config)

    def supports_multimodal(self, model_config) -> bool:
        c = model_config.hf_config
        return c.get("vision_config") is not None or (c.get("audio_config") is not None and not c.get("is_encoder_decoder"))


MULTIMODAL_REGISTRY = _Registry()
