"""Usage statistics (reference: vllm/usage/usage_lib.py — ``UsageMessage.report_usage``, opt-out through
``VLLM_NO_USAGE_STATS`` / ``DO_NOT_TRACK`` / ``~/.config/vllm/do_not_track``, ``set_runtime_usage_data``).

The reference posts one record at start-up (platform, GPU, model architecture, dtype, parallel sizes, context) and a
heartbeat every ten minutes to a stats server.  This build keeps the record and the opt-out rules but is local-first:
the record is appended to ``$VLLM_B200_CONFIG_ROOT/usage_stats.json`` (one JSON object per line) and is only POSTed
when ``VLLM_B200_USAGE_STATS_SERVER`` names a server, so an air-gapped node never opens a socket for it.
"""
from __future__ import annotations

import json
import os
import platform
import threading
import time
import uuid
from typing import Any, Optional

_RUNTIME: dict[str, Any] = {}
_CONFIG_ROOT = os.path.expanduser(os.environ.get("VLLM_B200_CONFIG_ROOT", "~/.config/vllm_b200"))


def set_runtime_usage_data(key: str, value: Any) -> None:
    """Facts learnt after start-up (e.g. which API front was used) ride along with the next heartbeat."""
    _RUNTIME[key] = value


def is_usage_stats_enabled() -> bool:
    if os.environ.get("VLLM_NO_USAGE_STATS", "0") == "1" or os.environ.get("VLLM_B200_NO_USAGE_STATS", "0") == "1":
        return False
    if os.environ.get("DO_NOT_TRACK", "0") == "1":
        return False
    root = os.path.expanduser(os.environ.get("VLLM_B200_CONFIG_ROOT", _CONFIG_ROOT))
    return not os.path.exists(os.path.join(root, "do_not_track"))


def _gpu_facts() -> dict[str, Any]:
    import torch
    if not torch.cuda.is_available():
        return {"gpu_count": 0, "gpu_type": None, "gpu_memory_per_device": None}
    p = torch.cuda.get_device_properties(0)
    return {"gpu_count": torch.cuda.device_count(), "gpu_type": p.name, "gpu_memory_per_device": p.total_memory,
            "sm": f"{p.major}{p.minor}", "cuda_runtime": torch.version.cuda}


class UsageMessage:
    def __init__(self) -> None:
        self.uuid = str(uuid.uuid4())
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None

    @property
    def path(self) -> str:                                     # resolved at use so the root can be redirected late
        return os.path.join(os.path.expanduser(os.environ.get("VLLM_B200_CONFIG_ROOT", _CONFIG_ROOT)), "usage_stats.json")

    def collect(self, vllm_config, usage_context: str, extra: Optional[dict[str, Any]] = None) -> dict[str, Any]:
        import torch
        mc, pc, cc = vllm_config.model_config, vllm_config.parallel_config, vllm_config.cache_config
        rec = {"uuid": self.uuid, "log_time": int(time.time() * 1e9), "source": os.environ.get("VLLM_USAGE_SOURCE", "production"),
               "context": usage_context, "provider": _cloud_provider(), "architecture": platform.machine(),
               "platform": platform.platform(), "num_cpu": os.cpu_count(), "torch_version": torch.__version__,
               "model_architecture": mc.architecture, "dtype": str(mc.torch_dtype), "max_model_len": mc.max_model_len,
               "tensor_parallel_size": pc.tensor_parallel_size, "pipeline_parallel_size": pc.pipeline_parallel_size,
               "data_parallel_size": pc.data_parallel_size, "enable_expert_parallel": pc.enable_expert_parallel,
               "block_size": cc.block_size, "kv_cache_dtype": cc.cache_dtype,
               "enable_prefix_caching": cc.enable_prefix_caching,
               "quantization": getattr(mc, "quantization", None), **_gpu_facts(), **(extra or {})}
        return rec

    def _write(self, rec: dict[str, Any]) -> None:
        try:
            os.makedirs(os.path.dirname(self.path), exist_ok=True)
            with open(self.path, "a") as f:
                f.write(json.dumps(rec, default=str) + "\n")
        except OSError:
            return
        server = os.environ.get("VLLM_B200_USAGE_STATS_SERVER")
        if server:
            try:
                import requests
                requests.post(server, json=rec, timeout=2)
            except Exception:                                  # noqa: BLE001 - telemetry must never break serving
                pass

    def report_usage(self, vllm_config, usage_context: str = "ENGINE_CONTEXT", extra: Optional[dict[str, Any]] = None,
                     heartbeat_s: float = 600.0) -> bool:
        """Start-up record now, heartbeats on a daemon thread.  Returns False (and does nothing) when opted out."""
        if not is_usage_stats_enabled():
            return False
        rec = self.collect(vllm_config, usage_context, extra)

        def run() -> None:
            self._write(rec)
            while not self._stop.wait(heartbeat_s):
                self._write({"uuid": self.uuid, "log_time": int(time.time() * 1e9), **_RUNTIME})
        self._thread = threading.Thread(target=run, daemon=True, name="usage-stats")
        self._thread.start()
        return True

    def stop(self) -> None:
        self._stop.set()


def _cloud_provider() -> str:
    for path, names in (("/sys/class/dmi/id/product_version", {"amazon": "AWS"}),
                        ("/sys/class/dmi/id/bios_vendor", {"google": "GCP"}),
                        ("/sys/class/dmi/id/chassis_asset_tag", {"7783-7084-3265-9085-8269-3286-77": "AZURE"}),
                        ("/sys/class/dmi/id/sys_vendor", {"oraclecloud": "OCI", "nebius": "NEBIUS"})):
        try:
            with open(path) as f:
                v = f.read().lower()
            for k, name in names.items():
                if k in v:
                    return name
        except OSError:
            continue
    return "UNKNOWN"


usage_message = UsageMessage()
