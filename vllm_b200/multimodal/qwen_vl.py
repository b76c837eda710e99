"""Qwen2-VL front-end: dynamic-resolution image patching and 3-D (temporal, height, width) rotary position ids.

Behavioural reference: vllm/model_executor/models/qwen2_vl.py (``Qwen2VLMultiModalProcessor``), the HF
``Qwen2VLImageProcessor`` it delegates to (smart resize to multiples of patch*merge, CLIP normalisation, merge-block
patch order) and ``get_rope_index`` / vllm MRotaryEmbedding.get_input_positions (M-RoPE ids + decode delta).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Any

import numpy as np
import torch

from vllm_b200.multimodal import CLIP_MEAN, CLIP_STD, hash_item


def smart_resize(h: int, w: int, factor: int = 28, min_pixels: int = 56 * 56, max_pixels: int = 14 * 14 * 4 * 1280) -> tuple[int, int]:
    if max(h, w) / min(h, w) > 200:
        raise ValueError("absolute aspect ratio must be smaller than 200")
    hb, wb = max(factor, round(h / factor) * factor), max(factor, round(w / factor) * factor)
    if hb * wb > max_pixels:
        beta = math.sqrt((h * w) / max_pixels)
        hb, wb = max(factor, math.floor(h / beta / factor) * factor), max(factor, math.floor(w / beta / factor) * factor)
    elif hb * wb < min_pixels:
        beta = math.sqrt(min_pixels / (h * w))
        hb, wb = math.ceil(h * beta / factor) * factor, math.ceil(w * beta / factor) * factor
    return hb, wb


def image_to_patches(img: Any, patch: int = 14, merge: int = 2, temporal: int = 2, min_pixels: int = 56 * 56,
                     max_pixels: int = 14 * 14 * 4 * 1280) -> tuple[torch.Tensor, tuple[int, int, int]]:
    """→ (patches [grid_h*grid_w, 3*temporal*patch*patch] in merge-block order, (1, grid_h, grid_w))."""
    if hasattr(img, "convert"):                                # PIL
        img = img.convert("RGB")
        w, h = img.size
        hb, wb = smart_resize(h, w, patch * merge, min_pixels, max_pixels)
        if (hb, wb) != (h, w):
            from PIL import Image
            img = img.resize((wb, hb), resample=Image.BICUBIC)
        arr = torch.from_numpy(np.asarray(img).copy()).permute(2, 0, 1).float()
    else:
        arr = torch.as_tensor(np.asarray(img))
        if arr.shape[-1] == 3:
            arr = arr.permute(2, 0, 1)
        arr = arr.float()
        h, w = arr.shape[1:]
        hb, wb = smart_resize(h, w, patch * merge, min_pixels, max_pixels)
        if (hb, wb) != (h, w):
            arr = torch.nn.functional.interpolate(arr[None], size=(hb, wb), mode="bicubic", align_corners=False)[0].clamp(0, 255)
    x = (arr / 255.0 - torch.tensor(CLIP_MEAN)[:, None, None]) / torch.tensor(CLIP_STD)[:, None, None]
    gh, gw = hb // patch, wb // patch
    x = x[None].repeat(temporal, 1, 1, 1)                                         # a still image fills the temporal patch
    x = x.reshape(1, temporal, 3, gh // merge, merge, patch, gw // merge, merge, patch)
    x = x.permute(0, 3, 6, 4, 7, 2, 1, 5, 8)
    return x.reshape(gh * gw, 3 * temporal * patch * patch).contiguous(), (1, gh, gw)


def mrope_positions(ids: list[int], grids: list[tuple[int, int, int]], image_token: int, merge: int = 2
                    ) -> tuple[torch.Tensor, int]:
    """[3, len(ids)] position ids (text: all three rows equal and sequential; image placeholders: (t, h, w) grid
    coordinates offset by the running position) and the delta that maps a later sequence index to its position."""
    pos = torch.zeros(3, len(ids), dtype=torch.long)
    i, st, g = 0, 0, 0
    n = len(ids)
    while i < n:
        if ids[i] == image_token and g < len(grids):
            t, h, w = grids[g]
            lh, lw = h // merge, w // merge
            L = t * lh * lw
            tt = torch.arange(t).view(-1, 1, 1).expand(t, lh, lw).reshape(-1)
            hh = torch.arange(lh).view(1, -1, 1).expand(t, lh, lw).reshape(-1)
            ww = torch.arange(lw).view(1, 1, -1).expand(t, lh, lw).reshape(-1)
            pos[:, i:i + L] = torch.stack([tt, hh, ww]) + st
            st += max(t, lh, lw)
            i += L
            g += 1
        else:
            pos[:, i] = st
            st += 1
            i += 1
    return pos, int(pos.max().item()) + 1 - n if n else 0


class Qwen2VLProcessor:
    """Prompt ids with one ``<|image_pad|>`` per image → expanded placeholders + feature records (+ M-RoPE ids)."""

    def __init__(self, model_config) -> None:
        cfg = model_config.hf_config
        vc = cfg["vision_config"]
        self.patch, self.merge = vc.get("patch_size", 14), vc.get("spatial_merge_size", 2)
        self.temporal = vc.get("temporal_patch_size", 2)
        self.image_token = cfg.get("image_token_id", 151655)
        self.min_pixels = cfg.get("min_pixels", 56 * 56)
        self.max_pixels = cfg.get("max_pixels", 14 * 14 * 4 * 1280)
        self.cache: "OrderedDict[str, Any]" = OrderedDict()

    def process(self, ids: list[int], mm_data: dict[str, Any]) -> tuple[list[int], list[dict[str, Any]]]:
        images = mm_data.get("image")
        if images is None:
            return ids, []
        if not isinstance(images, (list, tuple)):
            images = [images]
        slots = [i for i, t in enumerate(ids) if t == self.image_token]
        if len(slots) != len(images):
            raise ValueError(f"{len(images)} images but {len(slots)} image placeholders in the prompt")
        out, feats, grids, prev = [], [], [], 0
        for slot, img in zip(slots, images):
            px, grid = image_to_patches(img, self.patch, self.merge, self.temporal, self.min_pixels, self.max_pixels)
            n = grid[0] * grid[1] * grid[2] // (self.merge ** 2)
            out.extend(ids[prev:slot])
            feats.append({"modality": "image", "hash": hash_item(px.numpy()), "offset": len(out), "length": n,
                          "data": px, "grid": grid})
            out.extend([self.image_token] * n)
            grids.append(grid)
            prev = slot + 1
        out.extend(ids[prev:])
        pos, delta = mrope_positions(out, grids, self.image_token, self.merge)
        feats.append({"modality": "mrope", "hash": "", "offset": 0, "length": 0, "data": pos, "delta": delta})
        return out, feats
