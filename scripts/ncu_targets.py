"""Launch each hot kernel a few times so that `ncu --set full -k regex:<name> -s 3 -c 1` can capture it in isolation.

    ncu --set full --clock-control none --import-source on -k regex:gemm_fp8_kernel -s 3 -c 1 -o gpurun_out/ncu_fp8 \
        python scripts/ncu_targets.py fp8
"""
import sys

import torch

sys.path.insert(0, ".")
from vllm_b200.ops import cuda as cu  # noqa: E402
from vllm_b200.ops import ref  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.manual_seed(0)
dt = torch.bfloat16


def rep(fn, n=5):
    for _ in range(n):
        fn()
    torch.cuda.synchronize()


if which in ("fp8", "all"):
    M, N, K = 64, 7168, 16384
    x = torch.randn(M, K, device="cuda", dtype=dt) * 0.1
    wq, ws = ref.block_quant_weight_fp8(torch.randn(N, K, device="cuda") * 0.02, 128)
    xq, xs = cu.per_token_group_quant_fp8(x, 128, col_major=True)
    ws = ws.contiguous()
    rep(lambda: cu.gemm_fp8_block(xq, xs, wq, ws))
if which in ("gemm", "all"):
    x = torch.randn(64, 4096, device="cuda", dtype=dt) * 0.1
    for N, K in ((6144, 4096), (28672, 4096), (4096, 14336)):
        w = torch.randn(N, K, device="cuda", dtype=dt) * 0.02
        xx = torch.randn(64, K, device="cuda", dtype=dt) * 0.1
        rep(lambda: cu.gemm_tc(xx, w))
if which in ("attn", "mla", "all"):
    from vllm_b200.benchmarks.kernels import _mk_paged
    from vllm_b200.forward_context import AttentionMetadata
    B, BS, ctx = 64, 64, 1040
    if which in ("attn", "all"):
        Hq, Hkv, D = 32, 8, 128
        NB, bt, sl = _mk_paged(B, Hkv, D, BS, ctx)
        cache = (torch.randn(NB, 2, Hkv, BS, D, device="cuda") * 0.3).to(dt)
        q = (torch.randn(B, Hq, D, device="cuda") * 0.3).to(dt)
        md = AttentionMetadata(num_reqs=B, num_actual_tokens=B, max_query_len=1, max_seq_len=ctx,
                               query_start_loc=torch.arange(B + 1, dtype=torch.int32, device="cuda"), seq_lens=sl,
                               block_table=bt, slot_mapping=torch.empty(0), decode_only=True)
        md.kv_vt = True
        rep(lambda: cu.paged_attention(q, cache, md, D ** -0.5))
    if which in ("mla", "all"):
        NB, bt, sl = _mk_paged(B, 1, 576, BS, 4096)
        cache = (torch.randn(NB, BS, 576, device="cuda") * 0.3).to(dt)
        qn = (torch.randn(B, 16, 512, device="cuda") * 0.3).to(dt)
        qp = (torch.randn(B, 16, 64, device="cuda") * 0.3).to(dt)
        rep(lambda: cu.mla_decode_tc(qn, qp, cache, bt, sl, 0.07, 4096))
