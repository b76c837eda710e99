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
    if type(tok).__name__ == "TiktokenTokenizer":
        inv = {v: k for k, v in tok.special.items()}
        out = []
        for i in range(max(vocab_size, len(tok))):
            if i in inv:
                out.append((inv[i].encode(), True))
            else:
                try:
                    out.append((tok.enc.decode_single_token_bytes(i), False))
                except Exception:
                    out.append((b"", True))
        return out
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


_WS = [(9, 13), (32, 32), (133, 133), (160, 160), (5760, 5760), (8192, 8202), (8232, 8233), (8239, 8239), (8287, 8287),
       (12288, 12288)]                                      # Unicode White_Space (what \s means in the split patterns)
_PATTERNS = {
    0: r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+",
    1: r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+",
    2: r"'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+",
}


def _category_ranges(prefix: str) -> list[tuple[int, int]]:
    import unicodedata
    out, start = [], None
    for cp in range(0x110000):
        hit = unicodedata.category(chr(cp)).startswith(prefix)
        if hit and start is None:
            start = cp
        elif not hit and start is not None:
            out.append((start, cp - 1))
            start = None
    if start is not None:
        out.append((start, 0x10FFFF))
    return out


def encoder_spec(tok) -> Optional[dict]:
    """What the C++ BPE encoder needs, or None when the tokenizer is not a byte-level BPE with a known split pattern:
    {style, add_bos, bos_id, specials {str: id}, ranks [(priority, id, bytes)]}."""
    import json
    norm = lambda p: p.replace("(?i:'s|'t|'re|'ve|'m|'ll|'d)", "(?i:'s|'t|'re|'ve|'m|'ll|'d)").strip()  # noqa: E731
    if type(tok).__name__ == "TiktokenTokenizer":
        style = next((k for k, v in _PATTERNS.items() if norm(tok.enc._pat_str) == v), None)
        if style is None:
            return None
        ranks = [(r, r, b) for b, r in tok.enc._mergeable_ranks.items()]
        return dict(style=style, add_bos=tok.bos_token_id is not None, bos_id=tok.bos_token_id if tok.bos_token_id is not None else -1,
                    specials=dict(tok.special), ranks=ranks)
    bt = getattr(tok, "backend_tokenizer", None)
    if bt is None:
        return None
    cfg = json.loads(bt.to_str())
    model = cfg.get("model") or {}
    if model.get("type") != "BPE" or cfg.get("normalizer") not in (None, {}):
        return None
    pre = cfg.get("pre_tokenizer") or {}
    parts = pre.get("pretokenizers") if pre.get("type") == "Sequence" else [pre]
    style, byte_level = None, False
    for p_ in parts or []:
        if p_.get("type") == "ByteLevel":
            byte_level = True
            if p_.get("use_regex", True) and style is None:
                style = 2
            if p_.get("add_prefix_space"):
                return None
        elif p_.get("type") == "Split" and "Regex" in (p_.get("pattern") or {}):
            style = next((k for k, v in _PATTERNS.items() if norm(p_["pattern"]["Regex"]) == v), None)
            if style is None or p_.get("behavior") not in ("Isolated", None):
                return None
        else:
            return None
    if not byte_level or style is None:
        return None
    dec = _gpt2_byte_decoder()
    vocab = model["vocab"]
    to_b = lambda t: bytes(dec[ch] for ch in t)  # noqa: E731
    try:
        ranks = [(0xFFFFFFFE, i, to_b(t)) for t, i in vocab.items() if len(t) == 1]        # base alphabet: ids only
        pairs = []
        for rank, m in enumerate(model.get("merges") or []):
            a, b = m.split(" ") if isinstance(m, str) else m
            ranks.append((rank, vocab[a + b], to_b(a) + to_b(b)))
            pairs.append((vocab[a], vocab[b], rank, vocab[a + b]))
    except KeyError:
        return None
    specials = {t["content"]: t["id"] for t in cfg.get("added_tokens") or [] if t.get("special", True)}
    probe = tok.encode("a")
    bos = getattr(tok, "bos_token_id", None)
    return dict(style=style, add_bos=bool(probe) and bos is not None and probe[0] == bos, bos_id=bos if bos is not None else -1,
                specials=specials, ranks=ranks, pairs=pairs, whole_piece=bool(model.get("ignore_merges")))


def write_vocab_table(tok, vocab_size: int, path: str) -> None:
    table = token_bytes_table(tok, vocab_size)
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(table)))
        for b, sp in table:
            f.write(struct.pack("<BI", 1 if sp else 0, len(b)))
            f.write(b)
        spec = encoder_spec(tok)
        if spec is None:
            return
        f.write(b"ENC1")
        f.write(struct.pack("<IIq", spec["style"], 1 if spec["add_bos"] else 0, spec["bos_id"]))
        for ranges in (_category_ranges("L"), _category_ranges("N"), _WS):
            f.write(struct.pack("<I", len(ranges)))
            for lo, hi in ranges:
                f.write(struct.pack("<II", lo, hi))
        f.write(struct.pack("<I", len(spec["specials"])))
        for text, tid in spec["specials"].items():
            b = text.encode()
            f.write(struct.pack("<II", tid, len(b)) + b)
        f.write(struct.pack("<I", len(spec["ranks"])))
        for pr, tid, b in spec["ranks"]:
            f.write(struct.pack("<III", pr, tid, len(b)) + b)
        if spec.get("pairs"):                              # HF semantics: explicit pair merges
            f.write(struct.pack("<II", len(spec["pairs"]), 1 if spec.get("whole_piece") else 0))
            for l, r, rank, merged in spec["pairs"]:
                f.write(struct.pack("<IIII", l, r, rank, merged))


def spawn_frontd(*, host: str, port: int, in_addr: str, out_addr: str, model: str, max_model_len: int,
                 eos_token_id: Optional[int], vocab_path: Optional[str], num_engines: int = 1,
                 api_key: Optional[str] = None, chat_template: Optional[str] = None) -> subprocess.Popen:
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
    if chat_template:
        cmd += ["--chat-template", chat_template]
    return subprocess.Popen(cmd)


def _daemon_dialect(spec: Optional[str], model_type: Optional[str]) -> Optional[str]:
    """The chat dialect the C++ daemon renders natively (chatml / llama-3 / neutral role tags): the operator's
    ``--chat-template`` if it names one of them, else the model family's default when that is one of them."""
    native = {"chatml": "chatml", "llama-3": "llama-3", "llama3": "llama-3", "plain": "plain"}
    if spec and spec.lower() in native:
        return native[spec.lower()]
    fam = {"qwen2": "chatml", "qwen3": "chatml", "qwen2_moe": "chatml", "qwen3_moe": "chatml", "smollm3": "chatml"}
    return fam.get(model_type or "")


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
                      vocab_path=vocab_path, num_engines=n, api_key=args.api_key,
                      chat_template=_daemon_dialect(getattr(args, "chat_template", None), cfg.model_config.hf_config.get("model_type")))
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
