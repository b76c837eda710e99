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
                     max_pixels: int = 14 * 14 * 4 * 1280, mean=CLIP_MEAN, std=CLIP_STD) -> tuple[torch.Tensor, tuple[int, int, int]]:
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
    x = (arr / 255.0 - torch.tensor(mean)[:, None, None]) / torch.tensor(std)[:, None, None]
    gh, gw = hb // patch, wb // patch
    x = x[None].repeat(temporal, 1, 1, 1)                                         # a still image fills the temporal patch
    x = x.reshape(1, temporal, 3, gh // merge, merge, patch, gw // merge, merge, patch)
    x = x.permute(0, 3, 6, 4, 7, 2, 1, 5, 8)
    return x.reshape(gh * gw, 3 * temporal * patch * patch).contiguous(), (1, gh, gw)


def video_to_patches(frames: Any, patch: int = 14, merge: int = 2, temporal: int = 2, min_pixels: int = 56 * 56,
                     max_pixels: int = 14 * 14 * 4 * 1280) -> tuple[torch.Tensor, tuple[int, int, int]]:
    """Frames [T, H, W, 3] uint8 (array or list of PIL images / arrays) → (patches [grid_t*grid_h*grid_w,
    3*temporal*patch*patch], (grid_t, grid_h, grid_w)).  ``temporal`` consecutive frames form one temporal patch (the
    last frame is repeated when T is not a multiple); all frames share one smart-resized resolution."""
    if isinstance(frames, (list, tuple)):
        frames = np.stack([np.asarray(f.convert("RGB")) if hasattr(f, "convert") else np.asarray(f) for f in frames])
    arr = torch.as_tensor(np.asarray(frames))
    if arr.dim() != 4:
        raise ValueError("a video is [frames, height, width, 3]")
    if arr.shape[-1] == 3:
        arr = arr.permute(0, 3, 1, 2)
    arr = arr.float()
    T, _, h, w = arr.shape
    hb, wb = smart_resize(h, w, patch * merge, min_pixels, max_pixels)
    if (hb, wb) != (h, w):
        arr = torch.nn.functional.interpolate(arr, size=(hb, wb), mode="bicubic", align_corners=False).clamp(0, 255)
    if T % temporal:
        arr = torch.cat([arr, arr[-1:].expand(temporal - T % temporal, -1, -1, -1)])
    x = (arr / 255.0 - torch.tensor(CLIP_MEAN)[None, :, None, None]) / torch.tensor(CLIP_STD)[None, :, None, None]
    gt, gh, gw = x.shape[0] // temporal, hb // patch, wb // patch
    x = x.reshape(gt, temporal, 3, gh // merge, merge, patch, gw // merge, merge, patch)
    x = x.permute(0, 3, 6, 4, 7, 2, 1, 5, 8)
    return x.reshape(gt * gh * gw, 3 * temporal * patch * patch).contiguous(), (gt, gh, gw)


def mrope_positions(ids: list[int], grids: list[tuple[int, int, int]], image_token: int, merge: int = 2,
                    video_token: int = -1, t_scales: Any = None) -> tuple[torch.Tensor, int]:
    """[3, len(ids)] position ids (text: all three rows equal and sequential; image / video placeholders: (t, h, w)
    grid coordinates offset by the running position) and the delta that maps a later sequence index to its position.
    ``grids`` lists the items in order of appearance; ``t_scales[i]`` stretches item i's temporal index
    (Qwen2.5-VL: seconds per temporal patch x tokens_per_second)."""
    pos = torch.zeros(3, len(ids), dtype=torch.long)
    i, st, g = 0, 0, 0
    n = len(ids)
    while i < n:
        if ids[i] in (image_token, video_token) and g < len(grids):
            t, h, w = grids[g]
            lh, lw = h // merge, w // merge
            L = t * lh * lw
            tt = torch.arange(t).view(-1, 1, 1).expand(t, lh, lw).reshape(-1)
            if t_scales is not None and t_scales[g] != 1:
                tt = (tt.float() * float(t_scales[g])).long()
            hh = torch.arange(lh).view(1, -1, 1).expand(t, lh, lw).reshape(-1)
            ww = torch.arange(lw).view(1, 1, -1).expand(t, lh, lw).reshape(-1)
            pos[:, i:i + L] = torch.stack([tt, hh, ww]) + st
            st = int(pos[:, i:i + L].max().item()) + 1
            i += L
            g += 1
        else:
            pos[:, i] = st
            st += 1
            i += 1
    return pos, int(pos.max().item()) + 1 - n if n else 0


class Qwen2VLProcessor:
    """Prompt ids with one ``<|image_pad|>`` per image / ``<|video_pad|>`` per video → expanded placeholders + feature records (+ M-RoPE ids)."""

    def __init__(self, model_config) -> None:
        cfg = model_config.hf_config
        vc = cfg["vision_config"]
        self.patch, self.merge = vc.get("patch_size", 14), vc.get("spatial_merge_size", 2)
        self.temporal = vc.get("temporal_patch_size", 2)
        self.image_token = cfg.get("image_token_id", 151655)
        self.video_token = cfg.get("video_token_id", 151656)
        self.tokens_per_second = vc.get("tokens_per_second", 0)           # Qwen2.5-VL only
        self.min_pixels = cfg.get("min_pixels", 56 * 56)
        self.max_pixels = cfg.get("max_pixels", 14 * 14 * 4 * 1280)
        q3 = str(cfg.get("model_type", "")).startswith(("qwen3_vl", "qwen3_5"))
        self.mean, self.std = ((0.5,) * 3, (0.5,) * 3) if q3 else (CLIP_MEAN, CLIP_STD)
        if q3:
            self.min_pixels, self.max_pixels = cfg.get("min_pixels", 65536), cfg.get("max_pixels", 16 * 16 * 4 * 4096)
        self.cache: "OrderedDict[str, Any]" = OrderedDict()

    def process(self, ids: list[int], mm_data: dict[str, Any]) -> tuple[list[int], list[dict[str, Any]]]:
        def listify(v):
            if v is None:
                return []
            return list(v) if isinstance(v, (list, tuple)) else [v]
        images = listify(mm_data.get("image"))
        videos = mm_data.get("video")
        # one video is an array [T,H,W,3] or a list of frames; several videos are a list of those
        if videos is None:
            videos = []
        elif isinstance(videos, np.ndarray) or torch.is_tensor(videos):
            videos = [videos] if np.asarray(videos).ndim == 4 else list(videos)
        elif isinstance(videos, (list, tuple)) and videos and (hasattr(videos[0], "convert") or np.asarray(videos[0]).ndim == 3):
            videos = [videos]
        if not images and not videos:
            return ids, []
        fps = mm_data.get("fps") or mm_data.get("video_fps")
        n_is = sum(t == self.image_token for t in ids)
        n_vs = sum(t == self.video_token for t in ids)
        if n_is != len(images) or n_vs != len(videos):
            raise ValueError(f"{len(images)} images / {len(videos)} videos but {n_is} / {n_vs} placeholders in the prompt")
        out, feats, grids, scales, prev = [], [], [], [], 0
        ii = vi = 0
        for slot, tok in enumerate(ids):
            if tok == self.image_token:
                px, grid = image_to_patches(images[ii], self.patch, self.merge, self.temporal, self.min_pixels, self.max_pixels,
                                            self.mean, self.std)
                ii += 1
                modality, scale = "image", 1.0
            elif tok == self.video_token:
                px, grid = video_to_patches(videos[vi], self.patch, self.merge, self.temporal, self.min_pixels, self.max_pixels)
                f = fps[vi] if isinstance(fps, (list, tuple)) else fps
                # Qwen2.5-VL: temporal ids advance by (seconds per temporal patch) x tokens_per_second
                scale = (self.temporal / float(f) if f else 1.0) * self.tokens_per_second if self.tokens_per_second else 1.0
                vi += 1
                modality = "video"
            else:
                continue
            n = grid[0] * grid[1] * grid[2] // (self.merge ** 2)
            out.extend(ids[prev:slot])
            feats.append({"modality": modality, "hash": hash_item(px.numpy()), "offset": len(out), "length": n,
                          "data": px, "grid": grid})
            out.extend([tok] * n)
            grids.append(grid)
            scales.append(scale)
            prev = slot + 1
        out.extend(ids[prev:])
        pos, delta = mrope_positions(out, grids, self.image_token, self.merge, self.video_token, scales)
        feats.append({"modality": "mrope", "hash": "", "offset": 0, "length": 0, "data": pos, "delta": delta})
        return out, feats
