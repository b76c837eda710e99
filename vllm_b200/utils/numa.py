"""NUMA placement of worker processes (reference: vllm/utils/ numa helpers + ``--numa-bind`` in the worker start-up):
pin a worker to the CPUs (and prefer the memory) of the NUMA node its GPU hangs off, so the pinned staging buffers and
the shm broadcast ring are node-local.  Pure /sys + sched_setaffinity (no libnuma); a no-op where the topology is flat.
"""
from __future__ import annotations

import glob
import os
import re
from typing import Optional


def parse_cpulist(s: str) -> list[int]:
    out: list[int] = []
    for part in s.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            out.extend(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out


def numa_nodes(root: str = "/sys/devices/system/node") -> dict[int, list[int]]:
    nodes = {}
    for d in glob.glob(os.path.join(root, "node[0-9]*")):
        try:
            with open(os.path.join(d, "cpulist")) as f:
                nodes[int(re.sub(r".*node", "", d))] = parse_cpulist(f.read())
        except OSError:
            continue
    return nodes


def gpu_numa_node(gpu_index: int) -> Optional[int]:
    """NUMA node of a GPU from its PCI device entry (NVML bus id → /sys/bus/pci/devices/<id>/numa_node)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(gpu_index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            n = int(f.read().strip())
        return n if n >= 0 else None
    except Exception:
        return None


def bind_to_node(node: Optional[int], nodes: Optional[dict[int, list[int]]] = None) -> Optional[list[int]]:
    """Restrict the calling process to the CPUs of ``node``; returns the CPU list applied (None = left untouched)."""
    nodes = nodes if nodes is not None else numa_nodes()
    if node is None or node not in nodes or len(nodes) < 2:
        return None
    allowed = sorted(set(nodes[node]) & set(os.sched_getaffinity(0)))
    if not allowed:
        return None
    os.sched_setaffinity(0, allowed)
    return allowed


def bind_worker_to_gpu_node(gpu_index: int) -> Optional[list[int]]:
    if os.environ.get("VLLM_B200_NUMA_BIND", "0") != "1":
        return None
    return bind_to_node(gpu_numa_node(gpu_index))
