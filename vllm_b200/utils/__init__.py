"""Small shared helpers (reference: vllm/utils/*)."""
from __future__ import annotations

import itertools
import os
import socket
import time
import uuid
from typing import Any, Iterable, Sequence, TypeVar

T = TypeVar("T")


def cdiv(a: int, b: int) -> int:
    return -(a // -b)


def round_up(a: int, b: int) -> int:
    return cdiv(a, b) * b


def next_power_of_2(n: int) -> int:
    return 1 if n < 1 else 1 << (n - 1).bit_length()


def random_uuid() -> str:
    return uuid.uuid4().hex


def get_open_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def get_open_zmq_ipc_path() -> str:
    base = os.environ.get("VLLM_RPC_BASE_PATH", "/tmp")
    return f"ipc://{base}/vllm_b200_{uuid.uuid4().hex}"


def flatten(lists: Iterable[Iterable[T]]) -> list[T]:
    return list(itertools.chain.from_iterable(lists))


def chunk_list(lst: Sequence[T], n: int):
    for i in range(0, len(lst), n):
        yield lst[i:i + n]


def is_list_of(x: Any, typ: type) -> bool:
    return isinstance(x, list) and all(isinstance(v, typ) for v in x)


class Counter:
    def __init__(self, start: int = 0) -> None:
        self.v = start

    def __next__(self) -> int:
        v = self.v
        self.v += 1
        return v

    def reset(self) -> None:
        self.v = 0


class Timer:
    """``with Timer() as t: ...; t.ms``"""

    def __enter__(self):
        self.t0 = time.perf_counter()
        return self

    def __exit__(self, *a):
        self.ms = (time.perf_counter() - self.t0) * 1e3


def resolve_obj_by_qualname(qualname: str) -> Any:
    import importlib
    mod, _, name = qualname.rpartition(".")
    return getattr(importlib.import_module(mod), name)


GiB = 1 << 30
MiB = 1 << 20
