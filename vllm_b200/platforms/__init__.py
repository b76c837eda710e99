"""Platform layer (reference: vllm/platforms/* — ``CudaPlatform`` NVML-based, fully-connected NVLink check
``cuda.py:791``, ``current_platform``).  This framework targets exactly one accelerator platform (B200 /
sm_100a); the CPU platform exists for the reference-op test path."""
from __future__ import annotations

import os
import platform as _pyplatform
import subprocess
from functools import lru_cache
from typing import Optional


class Platform:
    device_type = "cpu"
    device_name = "cpu"
    dispatch_key = "CPU"
    dist_backend = "gloo"

    def is_cuda(self) -> bool:
        return False

    def get_device_name(self, device_id: int = 0) -> str:
        return _pyplatform.processor() or "cpu"

    def get_device_total_memory(self, device_id: int = 0) -> int:
        import psutil
        return psutil.virtual_memory().total

    def device_count(self) -> int:
        return 1

    def is_fully_connected(self, device_ids: list[int]) -> bool:
        return True

    def supports_fp8(self) -> bool:
        return False


class CpuPlatform(Platform):
    pass


class CudaPlatform(Platform):
    device_type = "cuda"
    device_name = "cuda"
    dispatch_key = "CUDA"
    dist_backend = "nccl"

    def is_cuda(self) -> bool:
        return True

    @staticmethod
    def _nvml():
        import pynvml
        pynvml.nvmlInit()
        return pynvml

    def device_count(self) -> int:
        import torch
        return torch.cuda.device_count()

    @lru_cache(maxsize=8)
    def get_device_capability(self, device_id: int = 0) -> tuple[int, int]:
        import torch
        return torch.cuda.get_device_capability(device_id)

    def is_blackwell(self, device_id: int = 0) -> bool:
        return self.get_device_capability(device_id)[0] == 10

    def get_device_name(self, device_id: int = 0) -> str:
        import torch
        return torch.cuda.get_device_name(device_id)

    def get_device_uuid(self, device_id: int = 0) -> str:
        nv = self._nvml()
        return nv.nvmlDeviceGetUUID(nv.nvmlDeviceGetHandleByIndex(self._physical(device_id)))

    def get_device_total_memory(self, device_id: int = 0) -> int:
        import torch
        return torch.cuda.get_device_properties(device_id).total_memory

    @staticmethod
    def _physical(device_id: int) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x for x in vis.split(",") if x.strip()]
            if device_id < len(ids) and ids[device_id].strip().isdigit():
                return int(ids[device_id])
        return device_id

    def is_fully_connected(self, device_ids: list[int]) -> bool:
        """True if every pair of the given GPUs is linked by NVLink (NVSwitch domains report NVLink P2P status
        OK for all pairs); the symmetric-heap collectives are enabled only then."""
        try:
            nv = self._nvml()
            hs = [nv.nvmlDeviceGetHandleByIndex(self._physical(i)) for i in device_ids]
            for i, a in enumerate(hs):
                for b in hs[i + 1:]:
                    st = nv.nvmlDeviceGetP2PStatus(a, b, nv.NVML_P2P_CAPS_INDEX_NVLINK)
                    if st != nv.NVML_P2P_STATUS_OK:
                        return False
            return True
        except Exception:
            return False

    def supports_fp8(self) -> bool:
        return self.get_device_capability()[0] >= 9

    def clocks(self, device_id: int = 0) -> dict:
        nv = self._nvml()
        h = nv.nvmlDeviceGetHandleByIndex(self._physical(device_id))
        return {"sm_mhz": nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
                "sm_max_mhz": nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM),
                "throttle_reasons": nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)}


_current: Optional[Platform] = None


def resolve_platform() -> Platform:
    from vllm_b200.plugins import load_platform_plugin
    plug = load_platform_plugin()
    if plug is not None:
        return plug
    try:
        import torch
        if torch.cuda.is_available():
            return CudaPlatform()
    except Exception:
        pass
    return CpuPlatform()


def __getattr__(name: str):
    global _current
    if name == "current_platform":
        if _current is None:
            _current = resolve_platform()
        return _current
    raise AttributeError(name)


def collect_env() -> str:
    """Environment report for bug reports (reference: vllm/collect_env.py)."""
    import torch
    from vllm_b200.version import __version__
    lines = [f"vllm_b200 {__version__}", f"python {_pyplatform.python_version()}  {_pyplatform.platform()}",
             f"torch {torch.__version__} (cuda {torch.version.cuda})  cuda available: {torch.cuda.is_available()}"]
    if torch.cuda.is_available():
        for i in range(torch.cuda.device_count()):
            p = torch.cuda.get_device_properties(i)
            lines.append(f"GPU {i}: {p.name} sm_{p.major}{p.minor} {p.total_memory / 2**30:.0f} GiB {p.multi_processor_count} SMs")
    try:
        lines.append(subprocess.run(["nvcc", "--version"], capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[-1])
    except Exception:
        lines.append("nvcc: not found")
    so = os.path.join(os.path.dirname(os.path.dirname(__file__)), "_C.so")
    lines.append(f"kernel library: {so} ({'present' if os.path.exists(so) else 'MISSING'})")
    for k, v in sorted(os.environ.items()):
        if k.startswith(("VLLM_", "NCCL_", "CUDA_")):
            lines.append(f"{k}={v}")
    return "\n".join(lines)
