"""Worker package: the per-device process body and the hooks that let a deployment replace or extend it.

``resolve_worker_class`` honours ``--worker-cls`` (a qualified class name used instead of the built-in worker) and
``--worker-extension-cls`` (a class whose methods are mixed into the worker so ``collective_rpc("name", ...)`` reaches
them on every rank — the RLHF weight-sync pattern).  Reference: vllm/v1/worker/worker_base.py
(``WorkerWrapperBase.init_worker``: ``resolve_obj_by_qualname`` + dynamic ``__bases__`` extension, attribute conflicts
rejected).
"""
from __future__ import annotations

import importlib
from typing import Any


def _by_qualname(name: str) -> Any:
    mod, sep, attr = name.replace(":", ".").rpartition(".")
    if not sep:
        raise ValueError(f"{name!r} is not a qualified name (expected 'package.module.Class')")
    return getattr(importlib.import_module(mod), attr)


def resolve_worker_class(vllm_config) -> type:
    """The class every executor instantiates, one per rank."""
    from vllm_b200.worker.worker import Worker
    pc = vllm_config.parallel_config
    name = getattr(pc, "worker_cls", "auto") or "auto"
    cls = Worker if name == "auto" else _by_qualname(name)
    if not isinstance(cls, type):
        raise TypeError(f"--worker-cls {name!r} does not name a class")
    ext_name = getattr(pc, "worker_extension_cls", "") or ""
    if not ext_name:
        return cls
    ext = _by_qualname(ext_name)
    if not isinstance(ext, type):
        raise TypeError(f"--worker-extension-cls {ext_name!r} does not name a class")
    if issubclass(cls, ext):
        return cls
    clash = sorted(a for a in vars(ext) if not a.startswith("__") and hasattr(cls, a))
    if clash:
        raise ValueError(f"worker extension {ext_name} redefines worker attributes {clash}; rename them")
    # a fresh subclass per (worker, extension) pair keeps the built-in class untouched for other engines in the process
    return type(f"{cls.__name__}With{ext.__name__}", (cls, ext), {"__module__": cls.__module__})
