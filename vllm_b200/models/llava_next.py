"""LLaVA-NeXT (LLaVA-1.6): any-resolution input — the image is fitted into the best of a few tile grids, cut into
``image_size`` tiles, every tile and a down-scaled copy of the whole image go through the CLIP tower + projector; the
tile features are laid back out as one big feature map, the padding the fit introduced is cropped away, a learned
``image_newline`` vector closes every feature row, and [whole-image tokens | rows] replace the placeholder.

Reference parity: vllm/model_executor/models/llava_next.py, HF ``modeling_llava_next`` (``pack_image_features``,
``unpad_image``, ``select_best_resolution``).
"""
from __future__ import annotations

import math
from typing import Any

import torch
import torch.nn as nn
import torch.nn.functional as F

from vllm_b200.config import VllmConfig
from vllm_b200.models.llava import LlavaForConditionalGeneration


def select_best_resolution(size: tuple[int, int], candidates: list) -> tuple[int, int]:
    """(height, width) of the candidate canvas that keeps the most image pixels, ties broken by least padding."""
    oh, ow = size
    best, best_eff, best_waste = None, 0, float("inf")
    for h, w in candidates:
        s = min(w / ow, h / oh)
        eff = min(int(ow * s) * int(oh * s), ow * oh)
        waste = w * h - eff
        if eff > best_eff or (eff == best_eff and waste < best_waste):
            best, best_eff, best_waste = (h, w), eff, waste
    return best


def unpadded_grid(cur_h: int, cur_w: int, orig: tuple[int, int]) -> tuple[int, int, int, int]:
    """Rows / columns of the feature map that hold image content: (row0, row1, col0, col1)."""
    oh, ow = orig
    if ow / oh > cur_w / cur_h:
        new_h = int(round(oh * (cur_w / ow), 7))
        pad = (cur_h - new_h) // 2
        return pad, cur_h - pad, 0, cur_w
    new_w = int(round(ow * (cur_h / oh), 7))
    pad = (cur_w - new_w) // 2
    return 0, cur_h, pad, cur_w - pad


class LlavaNextForConditionalGeneration(LlavaForConditionalGeneration):
    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__(vllm_config, prefix)
        cfg = self.cfg
        self.image_newline = nn.Parameter(torch.zeros(cfg["hidden_size"], dtype=vllm_config.model_config.torch_dtype), requires_grad=False)
        vc = cfg["vision_config"]
        self.side = vc["image_size"] // vc["patch_size"]
        self.tile = vc["image_size"]
        self.pinpoints = [tuple(p) for p in cfg.get("image_grid_pinpoints") or [(self.tile, self.tile)]]

    @torch.inference_mode()
    def embed_multimodal_item(self, f: dict) -> torch.Tensor:
        tiles = torch.as_tensor(f["data"]).to(self.image_newline.device)
        feats = self.embed_multimodal(tiles)                              # [T, side², H]
        if feats.shape[0] == 1:
            return torch.cat([feats[0], self.image_newline[None].to(feats.dtype)], dim=0)
        base, rest = feats[0], feats[1:]
        bh, bw = select_best_resolution(tuple(f["image_size"]), self.pinpoints)
        gh, gw = bh // self.tile, bw // self.tile
        s = self.side
        fmap = rest.view(gh, gw, s, s, -1).permute(4, 0, 2, 1, 3).reshape(-1, gh * s, gw * s)     # [H, rows, cols]
        r0, r1, c0, c1 = unpadded_grid(gh * s, gw * s, tuple(f["image_size"]))
        fmap = self._limit(fmap[:, r0:r1, c0:c1])
        nl = self.image_newline.to(fmap.dtype)[:, None, None].expand(-1, fmap.shape[1], 1)
        rows = torch.cat([fmap, nl], dim=-1).flatten(1, 2).t()
        return torch.cat([base, rows], dim=0)

    def _limit(self, fmap: torch.Tensor) -> torch.Tensor:
        return fmap

    def load_weights(self, weights):
        rest = []
        loaded = set()
        for n, w in weights:
            if n.endswith("image_newline"):
                self.image_newline.data.copy_(w)
                loaded.add("image_newline")
            else:
                rest.append((n, w))
        return super().load_weights(rest) | loaded


class LlavaNextProcessor:
    """Image → fitted canvas → tiles (+ the whole image at tile size first); ``{"tiles": [T,3,S,S], "image_size": (H, W)}``
    is accepted as-is (what an external HF processor returns as ``pixel_values`` / ``image_sizes``)."""

    def __init__(self, model_config) -> None:
        cfg = model_config.hf_config
        vc = cfg["vision_config"]
        self.tile, self.side = vc["image_size"], vc["image_size"] // vc["patch_size"]
        self.image_token = cfg.get("image_token_index", cfg.get("image_token_id"))
        self.pinpoints = [tuple(p) for p in cfg.get("image_grid_pinpoints") or [(self.tile, self.tile)]]
        self.default = cfg.get("vision_feature_select_strategy", "default") == "default"

    has_cls = True
    max_tiles = 0                       # > 0: feature maps larger than this many tiles' worth of patches are down-sampled

    def num_tokens(self, size: tuple[int, int], n_tiles: int) -> int:
        base = self.side ** 2 + (1 if self.has_cls and not self.default else 0)
        if n_tiles == 1:
            return base + 1
        bh, bw = select_best_resolution(size, self.pinpoints)
        r0, r1, c0, c1 = unpadded_grid(bh // self.tile * self.side, bw // self.tile * self.side, size)
        r, c = r1 - r0, c1 - c0
        if self.max_tiles:
            ratio = math.sqrt(r * c / (self.max_tiles * self.side ** 2))
            if ratio > 1.1:
                r, c = int(r // ratio), int(c // ratio)
        return base + r * (c + 1)

    def _tiles(self, img: Any) -> tuple[torch.Tensor, tuple[int, int]]:
        from vllm_b200.multimodal import CLIP_MEAN, CLIP_STD
        import numpy as np
        if isinstance(img, dict):
            return torch.as_tensor(img["tiles"]).float(), tuple(int(v) for v in img["image_size"])
        if hasattr(img, "convert"):
            img = np.array(img.convert("RGB"))
        t = torch.as_tensor(np.asarray(img)) if not isinstance(img, torch.Tensor) else img
        normalised = t.is_floating_point() and t.dim() == 3 and t.shape[0] == 3
        if t.dim() == 3 and t.shape[-1] == 3 and not normalised:
            t = t.permute(2, 0, 1)
        t = t.float() / (255.0 if not normalised and t.max() > 1.5 else 1.0)
        oh, ow = t.shape[1:]
        bh, bw = select_best_resolution((oh, ow), self.pinpoints)
        s = min(bw / ow, bh / oh)
        nh, nw = min(math.ceil(oh * s), bh), min(math.ceil(ow * s), bw)
        fit = F.interpolate(t[None], size=(nh, nw), mode="bicubic", align_corners=False)[0]
        canvas = torch.zeros(3, bh, bw)
        y0, x0 = (bh - nh) // 2, (bw - nw) // 2
        canvas[:, y0:y0 + nh, x0:x0 + nw] = fit
        tiles = [F.interpolate(t[None], size=(self.tile, self.tile), mode="bicubic", align_corners=False)[0]]
        for y in range(0, bh, self.tile):
            for x in range(0, bw, self.tile):
                tiles.append(canvas[:, y:y + self.tile, x:x + self.tile])
        px = torch.stack(tiles)
        if not normalised:
            px = (px.clamp(0, 1) - torch.tensor(CLIP_MEAN)[None, :, None, None]) / torch.tensor(CLIP_STD)[None, :, None, None]
        return px, (oh, ow)

    def process(self, ids: list[int], mm_data: dict[str, Any]) -> tuple[list[int], list[dict[str, Any]]]:
        from vllm_b200.multimodal import hash_item
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
            px, size = self._tiles(img)
            n = self.num_tokens(size, px.shape[0])
            out.extend(ids[prev:slot])
            feats.append({"modality": "image", "hash": hash_item(px.numpy()), "offset": len(out), "length": n, "data": px,
                          "image_size": size})
            out.extend([self.image_token] * n)
            prev = slot + 1
        out.extend(ids[prev:])
        return out, feats


class LlavaOnevisionForConditionalGeneration(LlavaNextForConditionalGeneration):
    """LLaVA-OneVision: the any-resolution scheme over a SigLIP tower (no class token, features taken before the tower's
    final norm), biased projector, Qwen2 decoder, and a cap on visual tokens — feature maps worth more than
    ``anyres_max_N`` tiles are bilinearly shrunk before the newline column is added (reference: models/llava_onevision.py)."""

    def _build_tower(self, cfg, dtype):
        from vllm_b200.models.idefics3 import SiglipVisionTower
        tower = SiglipVisionTower(cfg["vision_config"], dtype)
        inner = tower.forward
        tower.forward = lambda pixels: inner(pixels, final_norm=False)      # vision_feature_layer = -1: last encoder layer
        return tower

    @staticmethod
    def _tower_rename(n: str) -> str:
        from vllm_b200.models.idefics3 import SiglipVisionTower
        return SiglipVisionTower.rename(n)

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        cfg = vllm_config.model_config.hf_config
        cfg["vision_feature_select_strategy"] = "full"
        super().__init__(vllm_config, prefix)
        ar = str(cfg.get("vision_aspect_ratio", "anyres_max_9"))
        self.max_tiles = int(ar.rsplit("_", 1)[1]) if ar.startswith("anyres_max_") else 0

    def _limit(self, fmap):
        if not self.max_tiles:
            return fmap
        _, r, c = fmap.shape
        ratio = math.sqrt(r * c / (self.max_tiles * self.side ** 2))
        if ratio > 1.1:
            fmap = F.interpolate(fmap[None], [int(r // ratio), int(c // ratio)], mode="bilinear")[0]
        return fmap


class LlavaOnevisionProcessor(LlavaNextProcessor):
    has_cls = False

    def __init__(self, model_config) -> None:
        super().__init__(model_config)
        ar = str(model_config.hf_config.get("vision_aspect_ratio", "anyres_max_9"))
        self.max_tiles = int(ar.rsplit("_", 1)[1]) if ar.startswith("anyres_max_") else 0


class LlavaNextVideoForConditionalGeneration(LlavaNextForConditionalGeneration):
    """LLaVA-NeXT-Video: images as in LLaVA-NeXT; a video is a stack of frames, each through the tower, spatially pooled
    (stride-s average / max) and projected, the frames' tokens concatenated over the ``<video>`` placeholder
    (reference: models/llava_next_video.py)."""

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__(vllm_config, prefix)
        cfg = self.cfg
        self.video_token = cfg.get("video_token_index", cfg.get("video_token_id"))
        self.pool_stride = int(cfg.get("spatial_pool_stride", 2))
        self.pool_mode = cfg.get("spatial_pool_mode", "average")
        if self.pool_mode not in ("average", "max"):
            raise NotImplementedError(f"spatial_pool_mode={self.pool_mode!r}")

    @torch.inference_mode()
    def embed_multimodal_item(self, f: dict) -> torch.Tensor:
        if f.get("modality") != "video":
            return super().embed_multimodal_item(f)
        frames = torch.as_tensor(f["data"]).to(self.image_newline.device)
        x = self.vision_tower(frames)
        if self.select == "default":
            x = x[:, 1:]
        F_, N, C = x.shape
        s = int(N ** 0.5)
        g = x.view(F_, s, s, C).permute(0, 3, 1, 2)
        g = (F.avg_pool2d if self.pool_mode == "average" else F.max_pool2d)(g, self.pool_stride, self.pool_stride)
        x = g.flatten(2).transpose(1, 2)
        return self.linear_2(self.proj_act(self.linear_1(x))).reshape(-1, self.linear_2.out_features)

    def load_weights(self, weights):
        return super().load_weights((n, w) for n, w in weights if "vision_resampler" not in n)


class LlavaNextVideoProcessor(LlavaNextProcessor):
    def __init__(self, model_config) -> None:
        super().__init__(model_config)
        cfg = model_config.hf_config
        self.video_token = cfg.get("video_token_index", cfg.get("video_token_id"))
        self.stride = int(cfg.get("spatial_pool_stride", 2))

    def process(self, ids, mm_data):
        from vllm_b200.multimodal import hash_item, preprocess_image
        import numpy as np
        ids, feats = super().process(ids, mm_data)
        vids = mm_data.get("video")
        if vids is None:
            return ids, feats
        if isinstance(vids, torch.Tensor) and vids.dim() == 4 or isinstance(vids, np.ndarray) and vids.ndim == 4:
            vids = [vids]
        elif isinstance(vids, (list, tuple)) and vids and not (isinstance(vids[0], (torch.Tensor, np.ndarray)) and np.asarray(vids[0]).ndim == 4):
            vids = [vids]
        slots = [i for i, t in enumerate(ids) if t == self.video_token]
        if len(slots) != len(vids):
            raise ValueError(f"{len(vids)} videos but {len(slots)} video placeholders in the prompt")
        per = (self.side // self.stride) ** 2
        out, shift, prev = [], 0, 0
        new_feats = []
        for slot, v in zip(slots, vids):
            frames = torch.stack([preprocess_image(fr, self.tile) for fr in v])
            out.extend(ids[prev:slot])
            n = frames.shape[0] * per
            new_feats.append({"modality": "video", "hash": hash_item(frames.numpy()), "offset": len(out), "length": n, "data": frames})
            out.extend([self.video_token] * n)
            prev = slot + 1
        out.extend(ids[prev:])
        # image records sit at offsets of the pre-expansion prompt: shift those that come after an expanded video slot
        for f in feats:
            moved = sum(nf["length"] - 1 for nf, sl in zip(new_feats, slots) if sl < f["offset"])
            f["offset"] += moved
        return out, sorted(feats + new_feats, key=lambda r: r["offset"])
