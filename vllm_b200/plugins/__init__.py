"""Entry-point plugins (reference: vllm/plugins/__init__.py: ``vllm.general_plugins``,
``vllm.platform_plugins``, io-processor / LoRA-resolver groups).

Groups: ``vllm_b200.general_plugins`` (callables run once per process — typically they call
``models.registry.register_model`` / ``kv_transfer.register_connector`` / ``tool_parsers.register``),
``vllm_b200.platform_plugins`` (callable returning a Platform or None).  ``VLLM_PLUGINS`` restricts which
general plugins load (comma-separated names)."""
from __future__ import annotations

import os
from typing import Any, Callable, Optional

from vllm_b200.logger import init_logger

logger = init_logger(__name__)
GENERAL_GROUP = "vllm_b200.general_plugins"
PLATFORM_GROUP = "vllm_b200.platform_plugins"
_loaded = False


def _entry_points(group: str):
    from importlib.metadata import entry_points
    try:
        return list(entry_points(group=group))
    except Exception:  # pragma: no cover
        return []


def load_plugins_by_group(group: str) -> dict[str, Callable[[], Any]]:
    allowed = os.environ.get("VLLM_PLUGINS")
    allow = None if allowed is None else {x for x in allowed.split(",") if x}
    out = {}
    for ep in _entry_points(group):
        if allow is not None and ep.name not in allow:
            continue
        try:
            out[ep.name] = ep.load()
        except Exception:
            logger.exception("failed to load plugin %s from group %s", ep.name, group)
    return out


def load_general_plugins() -> None:
    """Idempotent; called by every engine / worker process before the model is resolved."""
    global _loaded
    if _loaded:
        return
    _loaded = True
    for name, fn in load_plugins_by_group(GENERAL_GROUP).items():
        try:
            fn()
            logger.info("loaded plugin %s", name)
        except Exception:
            logger.exception("plugin %s raised", name)


def load_platform_plugin() -> Optional[Any]:
    for name, fn in load_plugins_by_group(PLATFORM_GROUP).items():
        try:
            plat = fn()
            if plat is not None:
                return plat
        except Exception:
            logger.exception("platform plugin %s raised", name)
    return None
