"""Launcher for the C++ front-end daemon (``vllm_b200/_frontd``, sources in ``native/frontend/``).

Role of the reference's Rust front-end (rust/src/{engine-core-client,llm,text,tokenizer}: a native HTTP tier that
talks ZMQ + msgpack to Python engine cores).  Python only (1) starts the engine-core processes, (2) writes the
id → bytes detokenisation table the daemon mmaps, (3) supervises; no Python runs per request or per token.

``vllm_b200 serve MODEL --native-frontend`` or, for a split deployment, ``--headless --engine-input-addr ...
--engine-output-addr ...`` on the GPU node with the daemon started separately.
"""
from __future__ import annotations

import glob
import os
import signal
import struct
import subprocess
import tempfile
import time
from typing import Optional

from vllm_b200.logger import init_logger

logger = init_logger(__name__)


def find_libzmq() -> str:
    import zmq
    base = os.path.dirname(os.path.dirname(zmq.__file__))
    cands = glob.glob(os.path.join(base, "pyzmq.libs", "libzmq*.so*")) + glob.glob(os.path.join(os.path.dirname(zmq.__file__), "backend", "cython", "libzmq*.so*"))
    if not cands:
        import ctypes.util
        lib = ctypes.util.find_library("zmq")
        if lib:
            return lib
        raise RuntimeError("libzmq shared object not found (pyzmq wheel layout changed?)")
    return cands[0]


def _gpt2_byte_decoder() -> dict[str, int]:
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return {chr(c): b for b, c in zip(bs, cs)}


def token_bytes_table(tok, vocab_size: int) -> list[tuple[bytes, bool]]:
    """id → (raw bytes, is_special) for byte-level BPE (GPT-2 / Llama-3 / Qwen), SentencePiece-style (▁, <0xNN>) and
    the built-in byte tokenizer."""
    special = set(getattr(tok, "all_special_ids", ()) or ())
    if hasattr(tok, "_tok"):                                            # built-in ByteTokenizer
        return [(tok._tok(i), i in special) for i in range(vocab_size)]
    byte_level = False
    try:
        byte_level = "ByteLevel" in tok.backend_tokenizer.to_str()
    except Exception:
        pass
    dec = _gpt2_byte_decoder()
    n = max(vocab_size, len(tok))
    toks = tok.convert_ids_to_tokens(list(range(len(tok)))) + [None] * (n - len(tok))
    out = []
    for i, t in enumerate(toks):
        if t is None:
            out.append((b"", True))
        elif i in special:
            out.append((t.encode(), True))
        elif byte_level and all(ch in dec for ch in t):
            out.append((bytes(dec[ch] for ch in t), False))
        elif len(t) == 6 and t.startswith("<0x") and t.endswith(">"):
            out.append((bytes([int(t[3:5], 16)]), False))
        else:
            out.append((t.replace("▁", " ").encode(), False))
    return out


def write_vocab_table(tok, vocab_size: int, path: str) -> None:
    table = token_bytes_table(tok, vocab_size)
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(table)))
        for b, sp in table:
            f.write(struct.pack("<BI", 1 if sp else 0, len(b)))
            f.write(b)


def spawn_frontd(*, host: str, port: int, in_addr: str, out_addr: str, model: str, max_model_len: int,
                 eos_token_id: Optional[int], vocab_path: Optional[str], num_engines: int = 1,
                 api_key: Optional[str] = None) -> subprocess.Popen:
    from vllm_b200.native import build_frontd
    cmd = [build_frontd(), "--host", host, "--port", str(port), "--in-addr", in_addr, "--out-addr", out_addr,
           "--model", model, "--max-model-len", str(max_model_len), "--libzmq", find_libzmq(),
           "--num-engines", str(num_engines)]
    if eos_token_id is not None:
        cmd += ["--eos-token-id", str(eos_token_id)]
    if vocab_path:
        cmd += ["--vocab", vocab_path]
    if api_key:
        cmd += ["--api-key", api_key]
    return subprocess.Popen(cmd)


def run_native_server(args) -> None:
    """``serve --native-frontend``: daemon + engine cores, supervised until either side exits."""
    from vllm_b200.engine.arg_utils import EngineArgs
    from vllm_b200.engine.core_client import _launch_engines
    from vllm_b200.tokenizers import init_tokenizer_from_config
    from vllm_b200.utils import get_open_zmq_ipc_path
    if args.model_tag:
        args.model = args.model_tag
    ea = EngineArgs.from_cli_args(args)
    cfg = ea.create_engine_config()
    n = cfg.parallel_config.data_parallel_size
    in_addr, out_addr = get_open_zmq_ipc_path(), get_open_zmq_ipc_path()
    tmp = tempfile.mkdtemp(prefix="vllm_b200_frontd_")
    vocab_path = None
    tok = None if ea.skip_tokenizer_init else init_tokenizer_from_config(cfg.model_config)
    if tok is not None:
        vocab_path = os.path.join(tmp, "vocab.bin")
        write_vocab_table(tok, cfg.model_config.vocab_size, vocab_path)
    fd = spawn_frontd(host=args.host if args.host != "0.0.0.0" else "0.0.0.0", port=args.port, in_addr=in_addr,
                      out_addr=out_addr, model=ea.served_model_name or ea.model,
                      max_model_len=cfg.model_config.max_model_len, eos_token_id=cfg.model_config.eos_token_id,
                      vocab_path=vocab_path, num_engines=n, api_key=args.api_key)
    time.sleep(0.2)                                    # the daemon binds before the engines connect (ipc:// needs it)
    procs = _launch_engines(cfg, in_addr, out_addr, n, log_stats=not ea.disable_log_stats)
    stop = {"v": False}
    for sig in (signal.SIGINT, signal.SIGTERM):
        signal.signal(sig, lambda *a: stop.__setitem__("v", True))
    try:
        while not stop["v"]:
            if fd.poll() is not None:
                logger.error("front-end daemon exited with %s", fd.returncode)
                break
            if any(not p.is_alive() for p in procs):
                logger.error("an engine core exited; shutting down")
                break
            time.sleep(0.5)
    finally:
        if fd.poll() is None:
            fd.terminate()
        for p in procs:
            if p.is_alive():
                p.terminate()
        for p in procs:
            p.join(timeout=20)
        try:
            fd.wait(timeout=10)
        except subprocess.TimeoutExpired:
            fd.kill()


def run_headless(args) -> None:
    """``serve --headless``: engine cores only; they connect to a front-end (native or Python) at the given addresses
    (reference: entrypoints/cli/serve.py run_headless)."""
    from vllm_b200.engine.arg_utils import EngineArgs
    from vllm_b200.engine.core_client import _launch_engines
    if args.model_tag:
        args.model = args.model_tag
    if not args.engine_input_addr or not args.engine_output_addr:
        raise SystemExit("--headless needs --engine-input-addr and --engine-output-addr")
    ea = EngineArgs.from_cli_args(args)
    cfg = ea.create_engine_config()
    n_local = args.data_parallel_size_local or cfg.parallel_config.data_parallel_size
    start = args.data_parallel_start_rank or 0
    procs = _launch_engines(cfg, args.engine_input_addr, args.engine_output_addr, n_local,
                            log_stats=not ea.disable_log_stats, start=start)
    stop = {"v": False}
    for sig in (signal.SIGINT, signal.SIGTERM):
        signal.signal(sig, lambda *a: stop.__setitem__("v", True))
    try:
        while not stop["v"] and all(p.is_alive() for p in procs):
            time.sleep(0.5)
    finally:
        for p in procs:
            if p.is_alive():
                p.terminate()
        for p in procs:
            p.join(timeout=20)
