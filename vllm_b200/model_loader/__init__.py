"""Model construction + weight loading.

Reference: vllm/model_executor/model_loader/ (default safetensors/pt loader,
``dummy`` loader = random uniform ±1e-3 with a per-parameter seed
(weight_utils.py:1289-1349), sharded-state loader).
"""
from __future__ import annotations

import glob
import json
import os
import time
import zlib
from typing import Iterable, Iterator, Optional

import torch
import torch.nn as nn

from vllm_b200.config import VllmConfig, set_current_vllm_config
from vllm_b200.logger import init_logger
from vllm_b200.models.registry import resolve_model_cls

logger = init_logger(__name__)


def initialize_dummy_weights(model: nn.Module, seed: int = 1234, low: float = -1e-3, high: float = 1e-3) -> None:
    for name, p in model.state_dict().items():
        if not torch.is_floating_point(p):
            continue
        if name.endswith("cos_sin_cache"):
            continue
        g = torch.Generator(device=p.device)
        g.manual_seed(seed + (zlib.crc32(name.encode()) & 0xFFFF))
        if "norm" in name and name.endswith("weight") and p.dim() == 1:
            # keep norms near identity so activations stay well-scaled through many layers
            base = 0.0 if p.abs().max() == 0 else 1.0
            p.copy_(base + torch.empty(p.shape, dtype=torch.float32, device=p.device).uniform_(-0.05, 0.05, generator=g))
            continue
        if "scale" in name:
            p.fill_(1.0)
            continue
        if p.element_size() < 2:  # fp8: sample in fp16 then cast
            tmp = torch.empty(p.shape, dtype=torch.float16, device=p.device).uniform_(-1.0, 1.0, generator=g)
            p.copy_(tmp.to(p.dtype))
        else:
            p.uniform_(low, high, generator=g)


def _iter_safetensors(path: str) -> Iterator[tuple[str, torch.Tensor]]:
    from safetensors import safe_open
    files = sorted(glob.glob(os.path.join(path, "*.safetensors")))
    idx = os.path.join(path, "model.safetensors.index.json")
    if os.path.exists(idx):
        with open(idx) as f:
            files = sorted({os.path.join(path, v) for v in json.load(f)["weight_map"].values()})
    for fn in files:
        with safe_open(fn, framework="pt", device="cpu") as f:
            for k in f.keys():
                yield k, f.get_tensor(k)


def _iter_pt(path: str) -> Iterator[tuple[str, torch.Tensor]]:
    for fn in sorted(glob.glob(os.path.join(path, "*.bin")) + glob.glob(os.path.join(path, "*.pt"))):
        sd = torch.load(fn, map_location="cpu", weights_only=True)
        yield from sd.items()
        del sd


def _iter_fastsafetensors(path: str, extra: Optional[dict]) -> Iterator[tuple[str, torch.Tensor]]:
    """``--load-format fastsafetensors``: the fastsafetensors library reads whole files into a device-side buffer with many
    threads (GPUDirect Storage when the box has it: ``{"gds": true}`` in ``--model-loader-extra-config``) and hands out
    tensor views; the buffer of a file is released after its tensors were consumed, so each yielded tensor is copied into
    the parameter by the caller before the next file is opened (reference: model_loader/weight_utils.py
    ``fastsafetensors_weights_iterator``)."""
    from fastsafetensors import SafeTensorsFileLoader, SingleGroup
    extra = extra or {}
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    files = sorted(glob.glob(os.path.join(path, "*.safetensors")))
    per = max(1, int(extra.get("files_per_batch", 1)))
    for i in range(0, len(files), per):
        ld = SafeTensorsFileLoader(SingleGroup(), dev, nogds=not bool(extra.get("gds", False)), debug_log=False,
                                   max_threads=int(extra.get("concurrency", 16)))
        try:
            ld.add_filenames({0: files[i:i + per]})
            fb = ld.copy_files_to_device()
            try:
                for k in list(fb.key_to_rank_lidx.keys()):
                    yield k, fb.get_tensor(k)
            finally:
                fb.close()
        finally:
            ld.close()


def iter_weights(path: str, load_format: str = "auto", extra: Optional[dict] = None) -> Iterator[tuple[str, torch.Tensor]]:
    has_st = bool(glob.glob(os.path.join(path, "*.safetensors")))
    if load_format == "fastsafetensors":
        return _iter_fastsafetensors(path, extra)
    if load_format in ("streamer", "runai_streamer", "runai_streamer_sharded"):
        from vllm_b200.model_loader.streamer import stream_safetensors
        extra = extra or {}
        return stream_safetensors(path, int(extra.get("concurrency", os.environ.get("VLLM_B200_STREAMER_THREADS", 16))),
                                  extra.get("memory_limit"), pin_memory=torch.cuda.is_available())
    if load_format == "safetensors" or (load_format == "auto" and has_st):
        return _iter_safetensors(path)
    return _iter_pt(path)


def get_model(vllm_config: VllmConfig, device: torch.device,
              weights: Optional[Iterable[tuple[str, torch.Tensor]]] = None, prefix: str = "") -> nn.Module:
    mc, lc = vllm_config.model_config, vllm_config.load_config
    cls = resolve_model_cls(mc.architecture, getattr(mc, "model_impl", "auto"))
    t0 = time.perf_counter()
    old = torch.get_default_dtype()
    torch.set_default_dtype(mc.torch_dtype)
    try:
        with set_current_vllm_config(vllm_config), torch.device(device):
            model = cls(vllm_config, prefix=prefix) if prefix else cls(vllm_config)
    finally:
        torch.set_default_dtype(old)
    fmt = lc.load_format
    hq = mc.hf_config.get("quantization_config") or {}
    ct_weight_only = ((hq.get("quant_method") or "").lower() in ("compressed-tensors", "compressed_tensors")
                      and not next(iter((hq.get("config_groups") or {"_": {}}).values())).get("input_activations"))
    if (ct_weight_only or (hq.get("quant_method") or "").lower() in ("awq", "gptq", "nvfp4", "modelopt_fp4", "bitsandbytes")
            or (hq.get("quant_method") == "modelopt" and "FP4" in str(hq.get("quant_algo", "")).upper())):
        # weight-only int4 checkpoints are expanded to the model dtype while streaming (model_loader/weight_only.py)
        from vllm_b200.model_loader.weight_only import dequantize_stream
        _wrap = lambda it: dequantize_stream(it, hq, mc.torch_dtype)  # noqa: E731
    else:
        _wrap = lambda it: it  # noqa: E731
    with torch.no_grad():
        if weights is not None:
            model.load_weights(_wrap(weights))
        elif fmt in ("auto", "gguf") and mc.model.endswith(".gguf") and os.path.isfile(mc.model):
            from vllm_b200.model_loader.gguf import iter_gguf_weights
            model.load_weights(iter_gguf_weights(mc.model, mc.torch_dtype))
        elif fmt == "dummy" or (fmt == "auto" and not os.path.isdir(mc.model)):
            rng_ = float((getattr(lc, "model_loader_extra_config", None) or {}).get("dummy_range", 1e-3))
            initialize_dummy_weights(model, lc.dummy_seed, -rng_, rng_)
        elif fmt == "sharded_state":
            from vllm_b200.model_loader.sharded_state import load_sharded_state
            load_sharded_state(model, mc.model)
        else:
            loaded = model.load_weights(_wrap(iter_weights(mc.model, fmt, getattr(lc, 'model_loader_extra_config', None))))
            missing = {n for n, _ in model.named_parameters()} - set(loaded)
            missing = {m for m in missing if "lm_head" not in m}
            if missing:
                logger.warning("%d parameters were not found in the checkpoint, e.g. %s", len(missing),
                               sorted(missing)[:4])
        for m in model.modules():
            f = getattr(m, "process_weights_after_loading", None)
            if f is not None and m is not model:
                f()
    model.eval()
    n_bytes = sum(p.numel() * p.element_size() for p in model.parameters())
    logger.info("loaded %s (%s, %.2f GiB) in %.1fs", mc.architecture, fmt, n_bytes / 2**30,
                time.perf_counter() - t0)
    return model
