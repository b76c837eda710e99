"""Logging setup.

Mirrors the behaviour of the reference's ``vllm/logger.py`` (``init_logger``,
``info_once``/``warning_once``, JSON config through an env var) with a much
smaller implementation: one root logger ``vllm_b200`` configured once per
process, optional per-process prefix (cf. ``decorate_logs``).
"""
from __future__ import annotations

import json
import logging
import logging.config
import os
import sys
from functools import lru_cache

_FORMAT = "%(levelname)s %(asctime)s [%(filename)s:%(lineno)d] %(message)s"
_DATE_FORMAT = "%m-%d %H:%M:%S"
_configured = False
_prefix = ""


class _PrefixFilter(logging.Filter):
    def filter(self, record: logging.LogRecord) -> bool:
        if _prefix and not getattr(record, "_prefixed", False):
            record.msg = f"{_prefix}{record.msg}"
            record._prefixed = True
        return True


def _configure() -> None:
    global _configured
    if _configured:
        return
    _configured = True
    cfg_path = os.environ.get("VLLM_LOGGING_CONFIG_PATH")
    if cfg_path and os.path.exists(cfg_path):
        with open(cfg_path) as f:
            logging.config.dictConfig(json.load(f))
        return
    root = logging.getLogger("vllm_b200")
    level = os.environ.get("VLLM_LOGGING_LEVEL", "INFO").upper()
    root.setLevel(getattr(logging, level, logging.INFO))
    if os.environ.get("VLLM_CONFIGURE_LOGGING", "1") != "0" and not root.handlers:
        h = logging.StreamHandler(sys.stderr)
        h.setFormatter(logging.Formatter(_FORMAT, _DATE_FORMAT))
        h.addFilter(_PrefixFilter())
        root.addHandler(h)
        root.propagate = False


class _Logger(logging.LoggerAdapter):
    """Adapter adding ``*_once`` helpers."""

    def process(self, msg, kwargs):
        return msg, kwargs

    @lru_cache(maxsize=4096)
    def _once(self, level: int, msg: str) -> None:
        self.logger.log(level, msg, stacklevel=4)

    def info_once(self, msg: str, *args) -> None:
        self._once(logging.INFO, msg % args if args else msg)

    def warning_once(self, msg: str, *args) -> None:
        self._once(logging.WARNING, msg % args if args else msg)

    def debug_once(self, msg: str, *args) -> None:
        self._once(logging.DEBUG, msg % args if args else msg)


def init_logger(name: str) -> _Logger:
    _configure()
    if not name.startswith("vllm_b200"):
        name = "vllm_b200." + name
    return _Logger(logging.getLogger(name), {})


def set_process_prefix(prefix: str) -> None:
    """Prefix every log line of this process (e.g. ``(Worker_TP3) ``)."""
    global _prefix
    _prefix = prefix
