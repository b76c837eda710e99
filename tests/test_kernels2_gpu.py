"""Round-2 sm_100a kernels vs fp32 PyTorch references: activation quantisers and their norm / activation fusions,
fused q/k-norm + RoPE + cache insert, cache maintenance, multi-LoRA BGMV (run on the B200 box)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from vllm_b200.ops import ref  # noqa: E402


def cu():
    from vllm_b200.ops import cuda_ops
    return cuda_ops()


def close(a, b, atol=2e-2, rtol=2e-2):
    torch.testing.assert_close(a.float(), b.float(), atol=atol, rtol=rtol)


def deq_close(q, s, expect, rel=0.07):
    """fp8 round trip: |q * s - x| <= rel * max(|x|, s-granularity) elementwise (e4m3 has 3 mantissa bits)."""
    got = q.float() * s
    err = (got - expect.float()).abs()
    tol = rel * expect.float().abs() + 2.0 * torch.as_tensor(s).float().expand_as(got) * 2 ** -9 + 1e-6
    assert bool((err <= tol).all()), f"max err {err.max().item()} (tol {tol.max().item()})"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("T,H", [(1, 128), (7, 1024), (64, 4096), (33, 7168), (3, 16384)])
def test_scaled_fp8_quant(dtype, T, H):
    torch.manual_seed(0)
    x = torch.randn(T, H, device="cuda", dtype=dtype) * 3
    # dynamic per token: identical scales, codes equal to the reference quantiser up to rounding ties
    q, s = cu().scaled_fp8_quant(x, per_token=True)
    rq, rs = ref.scaled_fp8_quant(x, per_token=True)
    close(s, rs, atol=0, rtol=1e-6)
    assert (q.float() - rq.float()).abs().max() <= 32.0          # one e4m3 step at the top binade
    deq_close(q, s, x)
    # dynamic per tensor
    q, s = cu().scaled_fp8_quant(x)
    close(s.reshape(()), ref.scaled_fp8_quant(x)[1].reshape(()), atol=0, rtol=1e-6)
    deq_close(q, s, x)
    # static
    sc = torch.tensor([0.05], device="cuda")
    q, _ = cu().scaled_fp8_quant(x, sc)
    assert torch.equal(q.float(), ref.scaled_fp8_quant(x, sc)[0].float())


@pytest.mark.parametrize("group", [128, 64, 32])
@pytest.mark.parametrize("T,H", [(5, 512), (64, 7168), (17, 2048)])
def test_per_token_group_quant(group, T, H):
    torch.manual_seed(1)
    x = (torch.randn(T, H, device="cuda") * torch.rand(T, 1, device="cuda") * 4).to(torch.bfloat16)
    rq, rs = ref.per_token_group_quant_fp8(x, group)
    q, s = cu().per_token_group_quant_fp8(x, group)
    close(s, rs, atol=0, rtol=1e-6)

    def same_codes(a, b):
        # x * (1 / s) vs x / s may land on different sides of a rounding tie: at most one e4m3 step, on < 0.1 % of values
        d = (a.float() - b.float()).abs()
        assert bool((d <= 0.126 * b.float().abs().clamp_min(2.0 ** -6)).all()) and float((d > 0).float().mean()) < 5e-3
    same_codes(q, rq)
    # column-major scale storage (TMA-friendly) carries the same values
    q2, s2 = cu().per_token_group_quant_fp8(x, group, col_major=True)
    assert s2.shape == (T, H // group) and s2.stride(0) == 1
    close(s2, rs, atol=0, rtol=1e-6)
    assert torch.equal(q2.float(), q.float())
    # UE8M0: power-of-two scales >= the exact ones, dequantisation still within fp8 accuracy
    q3, s3 = cu().per_token_group_quant_fp8(x, group, ue8m0=True)
    assert bool((s3 >= rs * (1 - 1e-6)).all()) and bool((torch.log2(s3) == torch.log2(s3).round()).all())
    deq_close(q3, s3.repeat_interleave(group, 1), x)


def test_scaled_int8_quant():
    torch.manual_seed(2)
    x = torch.randn(19, 4096, device="cuda", dtype=torch.float16) * 2
    q, s = cu().scaled_int8_quant(x)
    rq, rs = ref.scaled_int8_quant(x)
    close(s, rs, atol=0, rtol=1e-6)
    assert (q.int() - rq.int()).abs().max() <= 1
    sc = torch.tensor([0.03], device="cuda")
    q, _ = cu().scaled_int8_quant(x, sc)
    assert (q.int() - ref.scaled_int8_quant(x, sc)[0].int()).abs().max() <= 1


@pytest.mark.parametrize("mode", ["static", "token", "group"])
@pytest.mark.parametrize("with_residual", [False, True])
def test_rms_norm_quant_fusion(mode, with_residual):
    torch.manual_seed(3)
    T, H = 48, 4096
    dt = torch.bfloat16
    x = torch.randn(T, H, device="cuda", dtype=dt)
    r = torch.randn(T, H, device="cuda", dtype=dt)
    w = (1 + 0.1 * torch.randn(H, device="cuda")).to(dt)
    sc = torch.tensor([0.02], device="cuda")
    if with_residual:
        y, r_exp = ref.fused_add_rms_norm(x.clone(), r.clone(), w, 1e-5)
    else:
        y, r_exp = ref.rms_norm(x, w, 1e-5), None
    r2 = r.clone() if with_residual else None
    q, s = cu().rms_norm_quant(x, w, 1e-5, residual=r2, scale=sc if mode == "static" else None, mode=mode)
    if with_residual:
        close(r2, r_exp, atol=1e-2)
    if mode == "static":
        deq_close(q, sc, y.float().clamp(-448 * 0.02, 448 * 0.02), rel=0.08)
    elif mode == "token":
        close(s, ref.scaled_fp8_quant(y, per_token=True)[1], atol=0, rtol=2e-2)
        deq_close(q, s, y, rel=0.08)
    else:
        close(s, ref.per_token_group_quant_fp8(y, 128)[1], atol=0, rtol=2e-2)
        deq_close(q, s.repeat_interleave(128, 1), y, rel=0.08)


@pytest.mark.parametrize("mode", ["static", "token", "group"])
def test_silu_mul_quant_fusion(mode):
    torch.manual_seed(4)
    T, H = 33, 2048
    x = torch.randn(T, 2 * H, device="cuda", dtype=torch.bfloat16)
    y = ref.silu_and_mul(x.float())
    sc = torch.tensor([0.01], device="cuda")
    q, s = cu().act_mul_quant(x, "silu", scale=sc if mode == "static" else None, mode=mode)
    if mode == "static":
        deq_close(q, sc, y.clamp(-4.48, 4.48), rel=0.08)
    elif mode == "token":
        deq_close(q, s, y, rel=0.08)
    else:
        deq_close(q, s.repeat_interleave(128, 1), y, rel=0.08)


def test_nvfp4_and_mxfp4_quant_roundtrip():
    torch.manual_seed(5)
    T, H = 40, 1024
    x = torch.randn(T, H, device="cuda", dtype=torch.bfloat16)
    gs = (448.0 * 6.0) / x.float().abs().max()
    q, sf = cu().scaled_fp4_quant(x, gs.reshape(1).float())
    deq = ref.fp4_dequant(q, sf.view(torch.float8_e4m3fn).float(), 1.0 / float(gs), block=16)
    rel = (deq - x.float()).norm() / x.float().norm()
    assert rel < 0.13, rel                       # e2m1: ~9-10 % relative error on gaussian data
    # MXFP4: block 32, e8m0 scale bytes
    q2, sf2 = cu().scaled_fp4_quant(x, mx=True)
    sc = torch.pow(2.0, sf2.float() - 127.0)
    deq2 = ref.fp4_dequant(q2, sc, 1.0, block=32)
    assert (deq2 - x.float()).norm() / x.float().norm() < 0.16
    # the swizzled scale layout holds the same bytes at the tile offsets
    q3, sfs = cu().scaled_fp4_quant(x, gs.reshape(1).float(), swizzled=True)
    assert torch.equal(q3, q)
    m, k = 37, 21
    off = ((m >> 7) * ((H // 16 + 3) // 4) + (k >> 2)) * 512 + (m & 31) * 16 + ((m >> 5) & 3) * 4 + (k & 3)
    assert int(sfs[off]) == int(sf[m, k])


@pytest.mark.parametrize("neox", [True, False])
@pytest.mark.parametrize("Hq,Hkv,D,rot", [(32, 8, 128, 128), (8, 2, 128, 64), (4, 4, 64, 64), (8, 1, 256, 256)])
def test_fused_qknorm_rope_cache(neox, Hq, Hkv, D, rot):
    torch.manual_seed(6)
    T, BS, NB = 29, 16, 12
    dt = torch.bfloat16
    from vllm_b200.layers.rotary import RotaryEmbedding
    rope = RotaryEmbedding(D, rot, 2048, 10000.0, neox).cuda()
    qkv = torch.randn(T, (Hq + 2 * Hkv) * D, device="cuda", dtype=dt)
    qw = (1 + 0.1 * torch.randn(D, device="cuda")).to(dt)
    kw = (1 + 0.1 * torch.randn(D, device="cuda")).to(dt)
    pos = torch.randint(0, 2048, (T,), device="cuda")
    slots = torch.randperm(NB * BS, device="cuda")[:T].long()
    slots[5] = -1
    q, k, v = qkv.split([Hq * D, Hkv * D, Hkv * D], dim=-1)
    eq, ek = ref.fused_qk_norm_rope(pos, q.clone(), k.clone(), qw, kw, rope.cos_sin_cache, D, 1e-6, neox)
    cache_ref = torch.zeros(NB, 2, Hkv, BS, D, device="cuda", dtype=dt)
    ref.reshape_and_cache(ek.view(T, Hkv, D), v.reshape(T, Hkv, D), cache_ref, slots)
    cache = torch.zeros_like(cache_ref)
    assert cu().qknorm_rope_ok(D, rot, neox)
    oq, ok = cu().fused_qk_norm_rope(pos, q, k, qw, kw, rope.cos_sin_cache, D, 1e-6, neox, v=v, kv_cache=cache,
                                     slot_mapping=slots)
    close(oq, eq, atol=3e-2)
    close(ok, ek, atol=3e-2)
    close(cache, cache_ref, atol=3e-2)


def test_swap_copy_gather_convert():
    torch.manual_seed(7)
    NB, BS, Hkv, D = 24, 16, 4, 128
    cache = torch.randn(NB, 2, Hkv, BS, D, device="cuda", dtype=torch.bfloat16)
    host = torch.zeros(8, 2, Hkv, BS, D, dtype=torch.bfloat16).pin_memory()
    pairs = [(3, 0), (17, 5), (9, 7)]
    cu().swap_blocks(cache, host, pairs)                       # swap out
    torch.cuda.synchronize()
    for s, d in pairs:
        assert torch.equal(host[d], cache[s].cpu())
    dst = torch.zeros_like(cache)
    cu().swap_blocks(host, dst, [(0, 11), (5, 2)])             # swap in
    torch.cuda.synchronize()
    assert torch.equal(dst[11].cpu(), host[0]) and torch.equal(dst[2].cpu(), host[5])
    c2 = cache.clone()
    cu().copy_blocks(c2, torch.tensor([[1, 20], [4, 21]], device="cuda"))
    assert torch.equal(c2[20], cache[1]) and torch.equal(c2[21], cache[4]) and torch.equal(c2[5], cache[5])
    # gather of MLA-style latent rows (bf16 and fp8 source)
    lat = torch.randn(NB, BS, 576, device="cuda", dtype=torch.bfloat16)
    lens = [37, 5, 64]
    bt = torch.randperm(NB, device="cuda")[:3 * 5].int().view(3, 5)
    cu_l = torch.tensor([0, 37, 42, 106], dtype=torch.int32, device="cuda")
    out = torch.zeros(106, 576, device="cuda", dtype=torch.bfloat16)
    cu().gather_cache(lat, bt, cu_l, out, max(lens))
    exp = torch.cat([lat[bt[b].long()].reshape(-1, 576)[:n] for b, n in enumerate(lens)])
    assert torch.equal(out, exp)
    lat8 = (lat.float() / 0.5).clamp(-448, 448).to(torch.float8_e4m3fn)
    out8 = torch.zeros_like(out)
    cu().gather_cache(lat8, bt, cu_l, out8, max(lens), scale=0.5)
    exp8 = torch.cat([lat8[bt[b].long()].reshape(-1, 576)[:n] for b, n in enumerate(lens)]).float() * 0.5
    close(out8, exp8, atol=1e-2)
    # convert
    x = torch.randn(1000, 64, device="cuda", dtype=torch.bfloat16)
    c = cu().convert_fp8(x, 0.25, True)
    assert torch.equal(c.float(), ref.convert_fp8(x, 0.25, True).float())
    close(cu().convert_fp8(c, 0.25, False), ref.convert_fp8(c, 0.25, False), atol=1e-2)


@pytest.mark.parametrize("n_sub,N", [(1, 4096), (3, 6144), (2, 2 * 1408)])
def test_lora_bgmv(n_sub, N):
    torch.manual_seed(8)
    T, K, r, S = 37, 2048, 16, 4
    dt = torch.bfloat16
    x = torch.randn(T, K, device="cuda", dtype=dt)
    y0 = torch.randn(T, N, device="cuda", dtype=dt)
    A = torch.randn(S, n_sub, r, K, device="cuda", dtype=dt) * 0.05
    B = torch.randn(S, N, r, device="cuda", dtype=dt) * 0.05
    slots = torch.randint(-1, S, (T,), device="cuda", dtype=torch.int64)
    if n_sub == 3:
        bounds = [4096, 5120]
    elif n_sub == 2:
        bounds = [1408]
    else:
        bounds = []
    offs = (n_sub, *(bounds + [N] * (3 - len(bounds))))
    y = y0.clone()
    cu().lora_bgmv(x, y, A, B, slots, offs)
    exp = y0.float().clone()
    edges = [0] + bounds + [N]
    for t in range(T):
        s = int(slots[t])
        if s < 0:
            continue
        for j in range(n_sub):
            v = A[s, j].float() @ x[t].float()
            exp[t, edges[j]:edges[j + 1]] += B[s, edges[j]:edges[j + 1]].float() @ v
    close(y, exp, atol=6e-2, rtol=3e-2)
    assert torch.equal(y[slots < 0], y0[slots < 0])


# ------------------------------------------------------------------ SSM / linear attention over the state pool
def _ssm_md(q_lens, ctx_lens, rows):
    from types import SimpleNamespace
    qsl = torch.tensor([0] + list(torch.tensor(q_lens).cumsum(0)), dtype=torch.int32, device="cuda")
    sl = torch.tensor([c + q for c, q in zip(ctx_lens, q_lens)], dtype=torch.int32, device="cuda")
    bt = torch.tensor(rows, dtype=torch.int32, device="cuda").view(-1, 1).repeat(1, 3).contiguous()
    return SimpleNamespace(num_reqs=len(q_lens), query_start_loc=qsl, seq_lens=sl, block_table=bt)


@pytest.mark.parametrize("K", [2, 4])
def test_causal_conv1d_with_cached_window(K):
    torch.manual_seed(10)
    C, q_lens, ctx = 384, [5, 1, 17, 1], [0, 9, 3, 0]
    rows = [3, 7, 1, 5]
    T = sum(q_lens)
    x = torch.randn(T, C, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(C, K, device="cuda", dtype=torch.bfloat16) * 0.5
    b = torch.randn(C, device="cuda", dtype=torch.bfloat16) * 0.1
    extra = 40
    pool = torch.randn(9, C * (K - 1) + extra, device="cuda", dtype=torch.float32)
    pool0 = pool.clone()
    md = _ssm_md(q_lens, ctx, rows)
    out = cu().causal_conv1d(x, w, b, pool, md)
    s = 0
    for i, n in enumerate(q_lens):
        win = pool0[rows[i], :C * (K - 1)].view(C, K - 1) if ctx[i] > 0 else torch.zeros(C, K - 1, device="cuda")
        xin = torch.cat([win.t(), x[s:s + n].float()], 0)                   # [K-1+n, C]
        conv = (xin.unfold(0, K, 1) * w.float()[None]).sum(-1) + b.float()
        close(out[s:s + n], torch.nn.functional.silu(conv), atol=2e-3, rtol=2e-3)
        close(pool[rows[i], :C * (K - 1)].view(C, K - 1), xin[-(K - 1):].t(), atol=1e-6, rtol=0)
        assert torch.equal(pool[rows[i], C * (K - 1):], pool0[rows[i], C * (K - 1):])      # rest of the row untouched
        s += n


def test_selective_scan_mamba1():
    torch.manual_seed(11)
    di, N, q_lens, ctx, rows = 256, 16, [6, 1, 33], [0, 4, 2], [2, 0, 5]
    T = sum(q_lens)
    x = torch.randn(T, di, device="cuda")
    dt = torch.nn.functional.softplus(torch.randn(T, di, device="cuda"))
    B, Cm = torch.randn(T, N, device="cuda"), torch.randn(T, N, device="cuda")
    A = -torch.rand(di, N, device="cuda") - 0.1
    D = torch.randn(di, device="cuda")
    off = 64
    pool = torch.randn(6, off + di * N, device="cuda")
    pool0 = pool.clone()
    y = cu().selective_scan(x, dt, B, Cm, A, D, pool, off, _ssm_md(q_lens, ctx, rows))
    s = 0
    for i, n in enumerate(q_lens):
        h = pool0[rows[i], off:].view(di, N).clone() if ctx[i] > 0 else torch.zeros(di, N, device="cuda")
        for t in range(s, s + n):
            h = torch.exp(dt[t][:, None] * A) * h + (dt[t] * x[t])[:, None] * B[t][None]
            close(y[t], (h * Cm[t][None]).sum(-1) + D * x[t], atol=2e-3, rtol=2e-3)
        close(pool[rows[i], off:].view(di, N), h, atol=2e-3, rtol=2e-3)
        s += n


@pytest.mark.parametrize("N,G", [(128, 1), (64, 2), (32, 4)])
def test_ssd_scan_mamba2_matches_chunked_reference(N, G):
    torch.manual_seed(12)
    nh, hd, q_lens, ctx, rows = 8, 16, [70, 1, 9], [0, 11, 5], [1, 4, 2]
    T = sum(q_lens)
    d_inner = nh * hd
    xbc = torch.randn(T, d_inner + 2 * G * N, device="cuda") * 0.5
    dt_raw = torch.randn(T, nh, device="cuda")
    dt_bias, A_log, D = torch.randn(nh, device="cuda") * 0.3, torch.randn(nh, device="cuda") * 0.3, torch.randn(nh, device="cuda")
    off = 24
    pool = torch.randn(6, off + nh * hd * N, device="cuda") * 0.3
    pool0 = pool.clone()
    y = cu().ssd_scan(xbc, dt_raw, dt_bias, A_log, D, pool, off, _ssm_md(q_lens, ctx, rows), nh, hd, G, N)
    s = 0
    for i, n in enumerate(q_lens):
        h0 = pool0[rows[i], off:].view(nh, hd, N).clone() if ctx[i] > 0 else None
        seg = xbc[s:s + n]
        xs, Bs, Cs = seg.split([d_inner, G * N, G * N], -1)
        dt = torch.nn.functional.softplus(dt_raw[s:s + n] + dt_bias)
        ye, hf = ref.ssd_chunk_scan(xs.view(n, nh, hd), dt, -torch.exp(A_log), Bs.view(n, G, N), Cs.view(n, G, N), 16, h0)
        ye = ye + D[None, :, None] * xs.view(n, nh, hd)
        close(y[s:s + n], ye.reshape(n, d_inner), atol=5e-3, rtol=5e-3)
        close(pool[rows[i], off:].view(nh, hd, N), hf, atol=5e-3, rtol=5e-3)
        s += n


def test_gdn_scan_matches_recurrence():
    torch.manual_seed(13)
    nv, dk, dv, q_lens, ctx, rows = 4, 64, 32, [12, 1, 5], [0, 3, 1], [0, 3, 2]
    T = sum(q_lens)
    q = torch.nn.functional.normalize(torch.randn(T, nv, dk, device="cuda"), dim=-1) * dk ** -0.5
    k = torch.nn.functional.normalize(torch.randn(T, nv, dk, device="cuda"), dim=-1)
    v = torch.randn(T, nv, dv, device="cuda")
    g = -torch.rand(T, nv, device="cuda")
    beta = torch.rand(T, nv, device="cuda")
    off = 16
    pool = torch.randn(5, off + nv * dk * dv, device="cuda") * 0.2
    pool0 = pool.clone()
    o = cu().gdn_scan(q, k, v, g, beta, pool, off, _ssm_md(q_lens, ctx, rows))
    s = 0
    for i, n in enumerate(q_lens):
        S = pool0[rows[i], off:].view(nv, dk, dv).clone() if ctx[i] > 0 else torch.zeros(nv, dk, dv, device="cuda")
        for t in range(s, s + n):
            S = S * g[t].exp()[:, None, None]
            delta = (v[t] - (S * k[t][..., None]).sum(-2)) * beta[t][:, None]
            S = S + k[t][..., None] * delta[:, None, :]
            close(o[t], (S * q[t][..., None]).sum(-2), atol=2e-3, rtol=2e-3)
        close(pool[rows[i], off:].view(nv, dk, dv), S, atol=2e-3, rtol=2e-3)
        s += n


# ------------------------------------------------------------------ tcgen05 fp8 GEMMs (csrc/gemm_fp8.cu)
def _fp8_block_ref(xq, xs, wq, ws):
    K = xq.shape[1]
    xd = xq.float() * xs.float().repeat_interleave(128, dim=1)[:, :K]
    wd = wq.float() * ws.repeat_interleave(128, 0)[:wq.shape[0]].repeat_interleave(128, 1)[:, :K]
    return xd @ wd.t()


@pytest.mark.parametrize("M,N,K", [(1, 512, 512), (16, 7168, 2048), (64, 4096, 7168), (33, 1536, 7168),
                                   (128, 2112, 7168), (100, 576, 1024), (48, 24576, 1536),
                                   (300, 4096, 7168), (1024, 2048, 4096)])
def test_gemm_fp8_block_scaled(M, N, K):
    torch.manual_seed(20)
    x = (torch.randn(M, K, device="cuda") * (1 + torch.rand(M, 1, device="cuda") * 3)).to(torch.bfloat16)
    w = torch.randn(N, K, device="cuda") * 0.05
    wq, ws = ref.block_quant_weight_fp8(w, 128)
    xq, xs = cu().per_token_group_quant_fp8(x, 128, col_major=True)
    y = cu().gemm_fp8_block(xq, xs, wq, ws.contiguous(), torch.bfloat16)
    exp = _fp8_block_ref(xq, xs, wq, ws)
    close(y, exp, atol=2e-2 * exp.abs().max().item() / 8 + 1e-2, rtol=2e-2)
    # end to end against the unquantised product: fp8 accuracy
    full = x.float() @ w.t()
    assert ((y.float() - full).norm() / full.norm()) < 0.06
    # the public op (fused quantiser + GEMM, bias)
    bias = torch.randn(N, device="cuda", dtype=torch.bfloat16)
    y2 = cu().block_fp8_linear(x, wq, ws, 128, bias)
    close(y2, exp + bias.float(), atol=2e-2 * exp.abs().max().item() / 8 + 2e-2, rtol=2e-2)


@pytest.mark.parametrize("M,N,K", [(8, 1024, 1024), (64, 6144, 4096), (128, 4096, 14336), (500, 3072, 2048)])
@pytest.mark.parametrize("per_token,per_channel", [(True, True), (False, False), (True, False)])
def test_gemm_fp8_scaled_mm(M, N, K, per_token, per_channel):
    torch.manual_seed(21)
    a = (torch.randn(M, K, device="cuda") * 2).clamp(-448, 448).to(torch.float8_e4m3fn)
    b = (torch.randn(N, K, device="cuda") * 2).clamp(-448, 448).to(torch.float8_e4m3fn)
    sa = torch.rand(M if per_token else 1, 1, device="cuda") * 0.02 + 0.001
    sb = torch.rand(N if per_channel else 1, 1, device="cuda") * 0.02 + 0.001
    bias = torch.randn(N, device="cuda", dtype=torch.bfloat16)
    y = cu().scaled_mm(a, b, sa, sb, torch.bfloat16, bias)
    exp = ref.scaled_mm(a, b, sa, sb, torch.float32, bias)
    close(y, exp, atol=1e-2 * exp.abs().max().item() + 1e-2, rtol=2e-2)


@pytest.mark.parametrize("E,N,K,R,bn", [(8, 1024, 1024, 64, 16), (4, 4096, 7168, 200, 64), (16, 512, 2048, 37, 32)])
def test_gemm_fp8_grouped(E, N, K, R, bn):
    torch.manual_seed(22)
    x = torch.randn(R, K, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(E, N, K, device="cuda") * 0.05
    wq = torch.empty(E, N, K, device="cuda", dtype=torch.float8_e4m3fn)
    ws = torch.empty(E, (N + 127) // 128, K // 128, device="cuda")
    for e in range(E):
        wq[e], ws[e] = ref.block_quant_weight_fp8(w[e], 128)
    cuts = torch.sort(torch.randint(0, R + 1, (E - 1,))).values.tolist()
    offs = torch.tensor([0] + cuts + [R], dtype=torch.int32, device="cuda")
    xq, xs = cu().per_token_group_quant_fp8(x, 128, col_major=True)
    y = cu().gemm_fp8_grouped(xq, xs, wq, ws, offs, bn)
    o = offs.tolist()
    for e in range(E):
        if o[e + 1] > o[e]:
            exp = _fp8_block_ref(xq[o[e]:o[e + 1]], xs[o[e]:o[e + 1]], wq[e], ws[e])
            close(y[o[e]:o[e + 1]], exp, atol=2e-2 * exp.abs().max().item() / 8 + 1e-2, rtol=2e-2)


# ------------------------------------------------------------------ tcgen05 MLA decode (csrc/mla_decode.cu)
@pytest.mark.parametrize("H", [16, 128, 8, 40])
@pytest.mark.parametrize("BS,splits", [(64, 1), (64, 3), (32, 1), (128, 2)])
def test_mla_decode_tc(H, BS, splits):
    torch.manual_seed(30)
    dt = torch.bfloat16
    seq_lens = [1, 63, 64, 65, 200, 777, 1300]
    B = len(seq_lens)
    max_blocks = max((s + BS - 1) // BS for s in seq_lens) + 1
    NB = B * max_blocks + 3
    cache = (torch.randn(NB, BS, 576, device="cuda") * 0.5).to(dt)
    bt = (torch.randperm(NB - 1, device="cuda")[:B * max_blocks].int() + 1).view(B, max_blocks).contiguous()
    sl = torch.tensor(seq_lens, dtype=torch.int32, device="cuda")
    qn = (torch.randn(B, H, 512, device="cuda") * 0.3).to(dt)
    qp = (torch.randn(B, H, 64, device="cuda") * 0.3).to(dt)
    scale = 0.07
    out, lse = cu().mla_decode_tc(qn, qp, cache, bt, sl, scale, max(seq_lens), num_splits=splits, return_lse=True)
    qsl = torch.arange(B + 1, dtype=torch.int32)
    exp = ref.mla_decode_attention(qn, qp, cache, bt, sl.cpu(), qsl, scale, 512)
    close(out, exp, atol=2e-2, rtol=3e-2)
    # lse against a direct computation for one sequence
    b = 4
    lat = cache[bt[b].long()].reshape(-1, 576)[:seq_lens[b]].float()
    att = (torch.cat([qn[b], qp[b]], -1).float() @ lat.t()) * scale
    close(lse[b], torch.logsumexp(att, -1), atol=2e-2, rtol=1e-2)


# ------------------------------------------------------------------ persistent stream-K decode GEMM (csrc/gemm_sk.cu)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K", [(1, 128, 512), (64, 6144, 4096), (64, 28672, 4096), (64, 4096, 14336), (17, 4096, 4096),
                                   (128, 1280, 8192), (200, 2048, 1024), (33, 1000, 2048), (64, 128256, 4096),
                                   (8, 12800, 256)])
def test_gemm_stream_k(dtype, M, N, K):
    torch.manual_seed(40)
    x = (torch.randn(M, K, device="cuda") * 0.5).to(dtype)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(dtype)
    b = torch.randn(N, device="cuda").to(dtype)
    exp = x.float() @ w.float().t()
    y = cu().gemm_sk(x, w)
    close(y, exp, atol=3e-2, rtol=2e-2)
    close(cu().gemm_sk(x, w, b), exp + b.float(), atol=3e-2, rtol=2e-2)
    # bit-identical run to run (fixed summation order of the stream-K partials), and no stale counters / workspace
    for _ in range(3):
        assert torch.equal(cu().gemm_sk(x, w), y)
    # strided activations (a slice of a fused buffer)
    xb = torch.zeros(M, K + 64, device="cuda", dtype=dtype)
    xb[:, :K] = x
    assert torch.equal(cu().gemm_sk(xb[:, :K], w), y)


@pytest.mark.parametrize("M,I,K", [(64, 14336, 4096), (5, 1792, 1024), (128, 3584, 8192)])
def test_gemm_stream_k_fused_silu_mul(M, I, K):
    torch.manual_seed(41)
    dt = torch.bfloat16
    x = (torch.randn(M, K, device="cuda") * 0.5).to(dt)
    wg = (torch.randn(I, K, device="cuda") * 0.05).to(dt)
    wu = (torch.randn(I, K, device="cuda") * 0.05).to(dt)
    w_il = torch.stack([wg, wu], dim=1).reshape(2 * I, K).contiguous()      # rows: gate_0, up_0, gate_1, up_1, ...
    y = cu().gemm_sk(x, w_il, act=True)
    g, u = x.float() @ wg.float().t(), x.float() @ wu.float().t()
    close(y, torch.nn.functional.silu(g) * u, atol=3e-2, rtol=3e-2)


@pytest.mark.parametrize("M,I,K", [(64, 14336, 4096), (64, 1792, 4096), (7, 896, 1024), (200, 3584, 8192)])
def test_gemm_tc_fused_silu_mul(M, I, K):
    """Tile / cluster split-K kernel with the SiLU·mul epilogue over interleaved (gate, up) rows (both store paths)."""
    torch.manual_seed(42)
    dt = torch.bfloat16
    x = (torch.randn(M, K, device="cuda") * 0.5).to(dt)
    wg = (torch.randn(I, K, device="cuda") * 0.05).to(dt)
    wu = (torch.randn(I, K, device="cuda") * 0.05).to(dt)
    w_il = torch.stack([wg, wu], dim=1).reshape(2 * I, K).contiguous()
    y = cu().gemm_tc(x, w_il, act=True)
    g, u = x.float() @ wg.float().t(), x.float() @ wu.float().t()
    close(y, torch.nn.functional.silu(g) * u, atol=3e-2, rtol=3e-2)
    from vllm_b200 import ops
    close(ops.gated_silu_linear(x, w_il), torch.nn.functional.silu(g) * u, atol=3e-2, rtol=3e-2)
    close(ops.silu_and_mul_interleaved(ops.linear(x, w_il)), torch.nn.functional.silu(g) * u, atol=4e-2, rtol=3e-2)


# ------------------------------------------------------------------ fused MoE combine + residual + norm (one GPU)
@pytest.mark.parametrize("T,E,k,H,I", [(64, 8, 2, 4096, 1792), (7, 16, 4, 1024, 512), (200, 8, 2, 2048, 1024)])
def test_fused_moe_reduce_norm_single_gpu(T, E, k, H, I):
    from vllm_b200.ops.moe import fused_moe_reduce_norm
    torch.manual_seed(50)
    dt = torch.bfloat16
    x = (torch.randn(T, H, device="cuda") * 0.5).to(dt)
    w13 = (torch.randn(E, 2 * I, H, device="cuda") * 0.03).to(dt)
    w2 = (torch.randn(E, H, I, device="cuda") * 0.03).to(dt)
    logits = torch.randn(T, E, device="cuda")
    tw, ids = ref.topk_softmax(logits, k, True)
    res = torch.randn(T, H, device="cuda").to(dt)
    nw = (1 + 0.1 * torch.randn(H, device="cuda")).to(dt)
    exp_moe = ref.fused_moe(x, w13, w2, tw, ids, "silu", None)
    e_norm, e_res = ref.fused_add_rms_norm(exp_moe.clone(), res.clone(), nw, 1e-5)
    for _ in range(3):                       # the accumulator must come back zeroed
        r2 = res.clone()
        out = fused_moe_reduce_norm(x, w13, w2, tw.float(), ids.int(), "silu", None, r2, nw, 1e-5)
        assert out is not None
        close(out[0], e_norm, atol=6e-2, rtol=4e-2)
        close(r2, e_res, atol=6e-2, rtol=4e-2)


def test_fused_moe_fp8_grouped_matches_dequantised_reference():
    from vllm_b200.ops.moe import fused_moe_fp8_grouped
    torch.manual_seed(60)
    T, E, k, H, I = 48, 8, 2, 2048, 1024
    dt = torch.bfloat16
    x = (torch.randn(T, H, device="cuda") * 0.5).to(dt)
    w13 = torch.randn(E, 2 * I, H, device="cuda") * 0.03
    w2 = torch.randn(E, H, I, device="cuda") * 0.03
    q13, s13, q2, s2 = [], [], [], []
    for e in range(E):
        a, b = ref.block_quant_weight_fp8(w13[e], 128); q13.append(a); s13.append(b)
        a, b = ref.block_quant_weight_fp8(w2[e], 128); q2.append(a); s2.append(b)
    q13, s13, q2, s2 = torch.stack(q13), torch.stack(s13), torch.stack(q2), torch.stack(s2)
    tw, ids = ref.topk_softmax(torch.randn(T, E, device="cuda"), k, True)
    y = fused_moe_fp8_grouped(x, q13, s13, q2, s2, tw.float(), ids.int(), None)
    exp = ref.fused_moe(x, w13.to(dt), w2.to(dt), tw, ids, "silu", None)
    rel = (y.float() - exp.float()).norm() / exp.float().norm()
    assert rel < 0.08, rel


def test_gpu_input_prep_kernels_match_numpy_expansion():
    """prep_step / block-table diffs / append_sampled vs the host-side expansion the runner does in numpy."""
    import numpy as np
    torch.manual_seed(70)
    R, L, MB, BS = 32, 512, 16, 32
    rng = np.random.default_rng(0)
    token_table = torch.randint(0, 30000, (R, L), dtype=torch.int32, device="cuda")
    block_rows = torch.zeros((R, MB), dtype=torch.int32, device="cuda")
    bt_np = rng.integers(1, 1000, (R, MB)).astype(np.int32)
    # stage the block tables through diffs: two appends per row
    d_row, d_start, d_cnt, d_off, flat = [], [], [], [], []
    for r in range(R):
        for s, n in ((0, 5), (5, MB - 5)):
            d_row.append(r); d_start.append(s); d_cnt.append(n); d_off.append(len(flat)); flat += bt_np[r, s:s + n].tolist()
    t = lambda a: torch.tensor(a, dtype=torch.int32, device="cuda")   # noqa: E731
    cu().block_table_apply_diffs(block_rows, t(d_row), t(d_start), t(d_cnt), t(d_off), t(flat))
    assert np.array_equal(block_rows.cpu().numpy(), bt_np)
    rows = rng.permutation(R)[:20].astype(np.int32)
    nc = rng.integers(0, 300, 20).astype(np.int32)
    ns = rng.integers(1, 40, 20).astype(np.int32)
    ns[3] = 1
    T = int(ns.sum())
    nblk = int((nc + ns).max() + BS - 1) // BS
    input_ids = torch.zeros(T, dtype=torch.int32, device="cuda")
    positions = torch.zeros(T, dtype=torch.int64, device="cuda")
    qsl = torch.zeros(21, dtype=torch.int32, device="cuda")
    seq_lens = torch.zeros(20, dtype=torch.int32, device="cuda")
    slots = torch.zeros(T, dtype=torch.int64, device="cuda")
    bt = torch.zeros((20, MB), dtype=torch.int32, device="cuda")
    cu().prep_step(t(rows), t(nc), t(ns), token_table, block_rows, input_ids, positions, qsl, seq_lens, slots, bt, nblk, BS)
    tt = token_table.cpu().numpy()
    cum = np.concatenate([[0], np.cumsum(ns)])
    assert np.array_equal(qsl.cpu().numpy(), cum) and np.array_equal(seq_lens.cpu().numpy(), nc + ns)
    for b in range(20):
        pos = nc[b] + np.arange(ns[b])
        sl = slice(cum[b], cum[b + 1])
        assert np.array_equal(input_ids.cpu().numpy()[sl], tt[rows[b], pos])
        assert np.array_equal(positions.cpu().numpy()[sl], pos)
        assert np.array_equal(slots.cpu().numpy()[sl], bt_np[rows[b], pos // BS].astype(np.int64) * BS + pos % BS)
        assert np.array_equal(bt.cpu().numpy()[b, :nblk], bt_np[rows[b], :nblk])
    # mixed-batch bucket padding: 24 request slots, T_pad tokens — empty padding requests, inert padding tokens
    Tp = T + 13
    input_ids2 = torch.full((Tp,), 7, dtype=torch.int32, device="cuda")
    positions2 = torch.full((Tp,), 7, dtype=torch.int64, device="cuda")
    slots2 = torch.full((Tp,), 7, dtype=torch.int64, device="cuda")
    qsl2 = torch.zeros(25, dtype=torch.int32, device="cuda")
    sl2 = torch.zeros(24, dtype=torch.int32, device="cuda")
    bt2 = torch.full((24, MB), 9, dtype=torch.int32, device="cuda")
    cu().prep_step(t(rows), t(nc), t(ns), token_table, block_rows, input_ids2, positions2, qsl2, sl2, slots2, bt2, nblk, BS,
                   pad_b=24, pad_t=Tp)
    assert torch.equal(input_ids2[:T], input_ids) and torch.equal(slots2[:T], slots)
    assert (input_ids2[T:] == 0).all() and (positions2[T:] == 0).all() and (slots2[T:] == -1).all()
    assert qsl2[:21].tolist() == cum.tolist() and (qsl2[21:] == T).all() and (bt2[20:, 0] == 0).all()
    # decode bucket padding: one-token requests on the null block
    ns1 = np.ones(20, dtype=np.int32)
    ids3 = torch.full((24,), 7, dtype=torch.int32, device="cuda")
    pos3 = torch.full((24,), 7, dtype=torch.int64, device="cuda")
    sl3 = torch.full((24,), 7, dtype=torch.int64, device="cuda")
    qsl3 = torch.zeros(25, dtype=torch.int32, device="cuda")
    seq3 = torch.zeros(24, dtype=torch.int32, device="cuda")
    cu().prep_step(t(rows), t(nc), t(ns1), token_table, block_rows, ids3, pos3, qsl3, seq3, sl3, bt2, nblk, BS, pad_b=24)
    assert qsl3.tolist() == list(range(25)) and (seq3[20:] == 1).all() and (sl3[20:] == -1).all() and (ids3[20:] == 0).all()
    assert np.array_equal(seq3[:20].cpu().numpy(), nc + 1)
    sampled = torch.randint(0, 30000, (20,), dtype=torch.int64, device="cuda")
    cu().append_sampled(token_table, t(rows), t(nc + ns), sampled)
    tt2 = token_table.cpu().numpy()
    for b in range(20):
        assert tt2[rows[b], nc[b] + ns[b]] == int(sampled[b])


@pytest.mark.parametrize("fmt", ["nvfp4", "mxfp4"])
@pytest.mark.parametrize("M,N,K", [(1, 256, 512), (7, 384, 1024), (64, 6144, 4096), (64, 4096, 14336), (128, 1000, 2048)])
def test_w4a16_fp4_gemm(fmt, M, N, K):
    """Weight-only FP4 decode GEMM (csrc/gemm_w4.cu: bulk-copied tile-major codes, PRMT expansion, tcgen05 kind::f16) vs
    x @ dequant(W)^T in fp32.  Every product of an e2m1 code and its block scale is exact in bf16, so the only
    difference to the reference is fp32 accumulation order."""
    torch.manual_seed(80 + M)
    dt = torch.bfloat16
    w = torch.randn(N, K, device="cuda") * 0.05
    x = (torch.randn(M, K, device="cuda") * 0.5).to(dt)
    bias = (torch.randn(N, device="cuda") * 0.1).to(dt) if N % 3 == 0 else None
    if fmt == "nvfp4":
        codes, sc, g = ref.fp4_quant(w, 16)
        wd = ref.fp4_dequant(codes, sc, g, 16)
    else:
        codes, sc = ref.mxfp4_quant(w)
        g = 1.0
        wd = ref.mxfp4_dequant(codes, sc)
    tiled = cu().pack_w4_tiles(codes, sc, fmt)
    y = cu().gemm_w4(x, tiled, N, g, fmt, bias)
    exp = x.float() @ wd.t() + (bias.float() if bias is not None else 0.0)
    err = (y.float() - exp).abs().max().item()
    assert err <= 2e-2 * exp.abs().max().item() + 1e-3, (err, exp.abs().max().item())
    # fp16 activations take the fp16 expansion tables
    if fmt == "nvfp4" and M == 7:
        y16 = cu().gemm_w4(x.half(), tiled, N, g, fmt, bias.half() if bias is not None else None)
        assert (y16.float() - exp).abs().max().item() <= 2e-2 * exp.abs().max().item() + 1e-2
