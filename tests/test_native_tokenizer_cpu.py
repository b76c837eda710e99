"""The C++ front-end's BPE encoder against its oracles: tiktoken (GPT-4 / Llama-3 and Qwen-2 split patterns) and the
HF ``tokenizers`` library (Split + ByteLevel, and GPT-2's ByteLevel regex) on adversarial Unicode strings."""
import json
import random
import subprocess

import pytest


def _strings(seed: int = 0, n: int = 300) -> list[str]:
    rng = random.Random(seed)
    atoms = ["hello", "World", " the", "  ", "   ", "\n", "\r\n", "\n\n", "\t", " \n ", "123", "4567890", "3.14", "I'm", "we'LL", "DON'T",
             "you've", "'s", "'", "''", "don't", " !?", "...", "—", "¿Qué?", "naïve", "北京", "日本語のテキスト", "𝒳𝒴", "🙂", "🙂🙂",
             "Ünï", "ß", "١٢٣", "Ⅻ", "½", "a_b", "foo(bar)", "x=1;", "   ", " ", "　", "é", "<|x|>", "<s>", "ab" * 9,
             " 1", "1 ", " a1b2 ", "\x00", "\x7f", "co-op", "C++", "#tag", "@me", "100%", "$5", "tab\tsep", "end."]
    out = ["", " ", "  ", "\n", "a", "hello world", "   leading", "trailing   ", "x\n\n\ny", "  \n  \n  z", "1234567", "a'b'c", "'re'RE're"]
    for _ in range(n):
        out.append("".join(rng.choice(atoms) for _ in range(rng.randint(1, 12))))
    return out


def _native_encode(vocab_path: str, texts: list[str]) -> list[list[int]]:
    from vllm_b200.native import build_frontd
    r = subprocess.run([build_frontd(), "--vocab", vocab_path, "--encode-stdin"], input=json.dumps(texts), capture_output=True,
                       text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    return json.loads(r.stdout)


def _ranks():
    ranks = {bytes([b]): b for b in range(256)}
    merges = [b"he", b"ll", b"hell", b"hello", b" w", b" wo", b" wor", b" worl", b" world", b" t", b" th", b" the", b"  ", b"   ", b"\n\n",
              b"12", b"123", b"in", b"ing", b"'t", b"'l", b"'ll", b"\xe5\x8c", b"\xe5\x8c\x97", b"\xf0\x9f", b"\xf0\x9f\x99", b"\xf0\x9f\x99\x82",
              b"..", b"...", b" !", b"ab", b"abab", b"\xc3\xbc", b"do", b"don"]
    for m in merges:
        ranks[m] = len(ranks)
    return ranks


@pytest.mark.parametrize("style", [0, 1])
def test_native_bpe_matches_tiktoken(tmp_path, style):
    from vllm_b200.entrypoints.native_frontend import _PATTERNS, write_vocab_table
    from vllm_b200.tokenizers.tiktoken_tok import TiktokenTokenizer
    ranks = _ranks()
    n = len(ranks)
    tok = TiktokenTokenizer(ranks, {"<|x|>": n, "<|endoftext|>": n + 1}, _PATTERNS[style], eos_token="<|endoftext|>")
    path = str(tmp_path / "vocab.bin")
    write_vocab_table(tok, len(tok), path)
    texts = _strings(style)
    got = _native_encode(path, texts)
    for t, g in zip(texts, got):
        assert g == tok.encode(t), repr(t)


@pytest.mark.parametrize("kind", ["split", "gpt2"])
def test_native_bpe_matches_hf_tokenizers(tmp_path, kind):
    from tokenizers import Regex, Tokenizer, decoders, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    from vllm_b200.entrypoints.native_frontend import _PATTERNS, write_vocab_table
    alphabet = sorted(pre_tokenizers.ByteLevel.alphabet())
    vocab = {ch: i for i, ch in enumerate(alphabet)}
    enc = {b: ch for ch, b in __import__("vllm_b200.entrypoints.native_frontend", fromlist=["x"])._gpt2_byte_decoder().items()}
    merges = []
    for m in [b"he", b"ll", b"hell", b"hello", b" w", b" wo", b" t", b" th", b" the", b"12", b"123", b"in", b"ing", b"..", b"...",
              b"\xe5\x8c", b"\xe5\x8c\x97", b"ab", b"abab", b"  "]:
        # split every merge as (longest existing prefix, rest) the way a trained vocabulary would have produced it
        s = "".join(enc[b] for b in m)
        for cut in range(len(s) - 1, 0, -1):
            if s[:cut] in vocab and s[cut:] in vocab:
                merges.append((s[:cut], s[cut:]))
                vocab[s] = len(vocab)
                break
    vocab["<|x|>"] = len(vocab)
    tk = Tokenizer(models.BPE(vocab={k: v for k, v in vocab.items() if k != "<|x|>"}, merges=merges))
    if kind == "split":
        tk.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Split(Regex(_PATTERNS[0]), behavior="isolated"),
                                                    pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])
    else:
        tk.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=True)
    tk.decoder = decoders.ByteLevel()
    tk.add_special_tokens(["<|x|>"])
    hf = PreTrainedTokenizerFast(tokenizer_object=tk)
    path = str(tmp_path / "vocab.bin")
    write_vocab_table(hf, len(hf), path)
    texts = _strings(7)
    got = _native_encode(path, texts)
    for t, g in zip(texts, got):
        assert g == hf.encode(t, add_special_tokens=False), repr(t)
