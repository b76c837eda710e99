"""ALiBi slopes and attention sinks in the generic paged-attention kernel vs the fp32 PyTorch reference.
(Last in collection order on purpose: newest kernel features run after the established suites.)"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(B, Hq, Hkv, D, BS, lens, qlens, dtype=torch.bfloat16):
    from vllm_b200.forward_context import AttentionMetadata
    dev = "cuda"
    nb = sum((l + BS - 1) // BS for l in lens) + 1
    kv = (torch.randn(nb, 2, Hkv, BS, D, device=dev) * 0.5).to(dtype)
    maxb = max((l + BS - 1) // BS for l in lens)
    bt = torch.zeros(B, maxb, dtype=torch.int32, device=dev)
    nxt = 1
    for i, l in enumerate(lens):
        n = (l + BS - 1) // BS
        bt[i, :n] = torch.arange(nxt, nxt + n, dtype=torch.int32)
        nxt += n
    T = sum(qlens)
    q = (torch.randn(T, Hq, D, device=dev) * 0.5).to(dtype)
    qsl = torch.tensor([0] + list(torch.tensor(qlens).cumsum(0)), dtype=torch.int32, device=dev)
    sl = torch.tensor(lens, dtype=torch.int32, device=dev)
    md = AttentionMetadata(num_reqs=B, num_actual_tokens=T, max_query_len=max(qlens), max_seq_len=max(lens), query_start_loc=qsl,
                           seq_lens=sl, block_table=bt, slot_mapping=torch.zeros(T, dtype=torch.int64, device=dev),
                           decode_only=all(x == 1 for x in qlens), kv_vt=False)
    return q, kv, md, bt, sl, qsl


@pytest.mark.parametrize("decode", [True, False])
@pytest.mark.parametrize("feature", ["alibi", "sinks", "both", "chunk"])
def test_paged_attention_alibi_and_sinks(decode, feature):
    from vllm_b200.ops import cuda as C
    from vllm_b200.ops import ref
    torch.manual_seed(0)
    B, Hq, Hkv, D, BS = 5, 8, 2, 64, 32
    lens = [33, 190, 64, 7, 301]
    qlens = [1] * B if decode else [5, 17, 64, 7, 33]
    q, kv, md, bt, sl, qsl = _setup(B, Hq, Hkv, D, BS, lens, qlens)
    slopes = torch.tensor([2.0 ** (-(i + 1)) for i in range(Hq)], device="cuda") if feature in ("alibi", "both") else None
    sinks = torch.randn(Hq, device="cuda") if feature in ("sinks", "both") else None
    chunk = 40 if feature == "chunk" else 0
    for window in (None, 48):
        out = C.paged_attention(q, kv, md, D ** -0.5, sliding_window=window, alibi_slopes=slopes, sinks=sinks, chunk_size=chunk)
        want = ref.paged_attention(q.float(), kv.float(), bt, sl.cpu(), qsl.cpu(), D ** -0.5, sliding_window=window,
                                   alibi_slopes=slopes, sinks=sinks, chunk_size=chunk)
        err = (out.float() - want).abs().max().item()
        assert err < 3e-2, (feature, decode, window, err)
    if decode:       # split-KV: the sink mass must be counted exactly once
        md.num_splits = 3
        out = C.paged_attention(q, kv, md, D ** -0.5, alibi_slopes=slopes, sinks=sinks, chunk_size=chunk)
        want = ref.paged_attention(q.float(), kv.float(), bt, sl.cpu(), qsl.cpu(), D ** -0.5, alibi_slopes=slopes, sinks=sinks,
                                   chunk_size=chunk)
        assert (out.float() - want).abs().max().item() < 3e-2
