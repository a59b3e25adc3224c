"""Kernel-level parity on the GPU, each kernel called through the C ABI (mmada_b200.ops -> ctypes).

Floating-point kernels (GEMM, attention, RMSNorm, RoPE) are compared with a plain PyTorch fp32
evaluation of the same op on the same bf16-rounded inputs; tolerances are stated per test.
Integer / decision kernels (sampling) are compared bit-exactly with the CPU oracle."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("cta_group", [1, 2], ids=["cg1", "cg2"])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 512, 192), (1000, 768, 256), (4096, 4096, 1024), (1539 * 2, 3072, 1024)])
def test_gemm_bf16_and_f32(M, N, K, cta_group):
    from mmada_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    ref = a.float() @ w.float().t()
    out32 = ops.gemm(a, w, ops.EPI_F32, cta_group=cta_group)
    # fp32 accumulation of exact bf16 products: only summation order differs
    assert _rel(out32, ref) < 1e-5
    out16 = ops.gemm(a, w, ops.EPI_BF16, cta_group=cta_group)
    assert _rel(out16.float(), ref) < 5e-3           # one bf16 rounding of the result


@pytest.mark.parametrize("cta_group", [1, 2], ids=["cg1", "cg2"])
def test_gemm_residual_swiglu_bias(cta_group):
    from mmada_b200 import ops
    M, K, F = 777, 512, 1024
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    wo = (torch.randn(512, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    x = torch.randn(M, 512, device="cuda", generator=g)
    ref = x + a.float() @ wo.float().t()
    out = ops.gemm(a, wo, ops.EPI_RESID_F32, out=x.clone(), aux=x, cta_group=cta_group)
    assert _rel(out, ref) < 1e-5
    xi = x.clone()                                    # in place: out aliases aux
    ops.gemm(a, wo, ops.EPI_RESID_F32, out=xi, aux=xi, cta_group=cta_group)
    assert torch.equal(xi, out)
    wg = (torch.randn(F, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    wu = (torch.randn(F, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    from mmada_b200.modeling_llada import interleave_gate_up
    wgu = interleave_gate_up(wg, wu)
    h = ops.gemm(a, wgu, ops.EPI_SWIGLU_BF16, cta_group=cta_group)
    gate, up = a.float() @ wg.float().t(), a.float() @ wu.float().t()
    ref = torch.nn.functional.silu(gate) * up
    assert h.shape == (M, F)
    assert _rel(h.float(), ref) < 6e-3
    bias = torch.randn(512, device="cuda", generator=g)
    ob = ops.gemm(a, wo, ops.EPI_BIAS_BF16, bias=bias, cta_group=cta_group)
    assert _rel(ob.float(), a.float() @ wo.float().t() + bias) < 5e-3
    x2 = torch.randn(M, 512, device="cuda", generator=g)
    o3 = ops.gemm(a, wo, ops.EPI_BIAS_RESID_F32, aux=x2, bias=bias, cta_group=cta_group)
    assert _rel(o3, a.float() @ wo.float().t() + bias + x2) < 1e-5
    wn = wo[:96].contiguous()                      # narrow output (BN = 128 path for pairs)
    o4 = ops.gemm(a, wn, ops.EPI_BIAS_F32, bias=bias[:96].contiguous(), cta_group=cta_group)
    assert _rel(o4, a.float() @ wn.float().t() + bias[:96]) < 1e-5


@pytest.mark.parametrize("d", [256, 512, 1024, 4096, 384])
def test_rmsnorm(d):
    from mmada_b200 import ops
    from oracle import llada
    g = torch.Generator(device="cuda").manual_seed(d)
    x = torch.randn(333, d, device="cuda", generator=g) * 3
    w = 1 + 0.1 * torch.randn(d, device="cuda", generator=g)
    out = ops.rmsnorm(x, w, 1e-5)
    ref = llada.rms_norm(x.cpu(), w.cpu(), 1e-5)
    assert _rel(out.float().cpu(), ref) < 5e-3        # bf16 output rounding
    rows = torch.tensor([5, 0, 332, 17], device="cuda", dtype=torch.int32)
    sub = ops.rmsnorm(x, w, 1e-5, rows=rows)
    assert torch.equal(sub, out[rows.long()])


@pytest.mark.parametrize("hd,H", [(64, 4), (128, 4)])
def test_rope(hd, H):
    from mmada_b200 import ops
    from oracle import llada
    B, L, d = 2, 77, hd * H
    g = torch.Generator(device="cuda").manual_seed(hd)
    qkv = torch.randn(B * L, 3 * d, device="cuda", generator=g).bfloat16()
    sin, cos = llada.rope_tables(L, hd, 500000.0)
    sin_h, cos_h = sin[0, 0, :, :hd // 2].contiguous().cuda(), cos[0, 0, :, :hd // 2].contiguous().cuda()
    q = qkv[:, :d].view(B, L, H, hd).transpose(1, 2).cpu()
    k = qkv[:, d:2 * d].view(B, L, H, hd).transpose(1, 2).cpu()
    qr, kr = llada.apply_rope(q, k, 500000.0)
    v0 = qkv[:, 2 * d:].clone()
    ops.rope_inplace(qkv, sin_h, cos_h, d, hd, L)
    q2 = qkv[:, :d].view(B, L, H, hd).transpose(1, 2).cpu()
    k2 = qkv[:, d:2 * d].view(B, L, H, hd).transpose(1, 2).cpu()
    # same fp32 op sequence on the same tables: identical up to the sin/cos table (shared) -> bit-exact
    assert torch.equal(q2, qr) and torch.equal(k2, kr)
    assert torch.equal(qkv[:, 2 * d:], v0)


@pytest.mark.parametrize("hd,H,L,B", [(64, 4, 77, 3), (128, 4, 100, 2), (128, 2, 387, 2)])
@pytest.mark.parametrize("cta_group", [1, 2], ids=["cg1", "cg2"])
def test_gemm_qkv_rope_fused_equals_unfused(hd, H, L, B, cta_group):
    """RoPE in the q|k|v GEMM epilogue == bf16 GEMM followed by the stand-alone RoPE kernel, bit for bit
    (same rounding points: bf16 projection output, fp32 rotation, bf16 result)."""
    from mmada_b200 import ops
    from oracle import llada
    d = H * hd
    g = torch.Generator(device="cuda").manual_seed(hd + L)
    a = torch.randn(B * L, d, device="cuda", generator=g).bfloat16()
    w = (torch.randn(3 * d, d, device="cuda", generator=g) / math.sqrt(d)).bfloat16()
    sin, cos = llada.rope_tables(L, hd, 500000.0)
    sin_h, cos_h = sin[0, 0, :, :hd // 2].contiguous().cuda(), cos[0, 0, :, :hd // 2].contiguous().cuda()
    ref = ops.gemm(a, w, ops.EPI_BF16, cta_group=cta_group)
    ops.rope_inplace(ref, sin_h, cos_h, d, hd, L)
    out = ops.gemm_qkv_rope(a, w, sin_h, cos_h, d, hd, L, cta_group=cta_group)
    assert torch.equal(out, ref)


@pytest.mark.parametrize("cta_group", [1, 2], ids=["cg1", "cg2"])
@pytest.mark.parametrize("M,d,F,hd", [(777, 512, 1024, 128), (1539 * 2, 1024, 2816, 64), (130, 256, 512, 64)])
def test_rmsnorm_folded_into_gemms(M, d, F, hd, cta_group):
    """RMSLayerNorm (models/modeling_llada.py:315-329) folded into the GEMMs around it: the producer epilogue
    (residual add + bf16 copy + per-tile row sums of squares), then the two consumers (SwiGLU and q|k|v + RoPE with the
    accumulator rows scaled by rstd) against fp32 torch evaluations of norm -> projection on the same bf16 operands."""
    from mmada_b200 import ops
    from mmada_b200.modeling_llada import interleave_gate_up
    from oracle import llada
    L = M // 2 if M % 2 == 0 else M
    eps = 1e-5
    g = torch.Generator(device="cuda").manual_seed(M + d)
    a = torch.randn(M, F, device="cuda", generator=g).bfloat16()
    wo = (torch.randn(d, F, device="cuda", generator=g) / math.sqrt(F)).bfloat16()
    x0 = torch.randn(M, d, device="cuda", generator=g) * 2
    # --- producer
    x = x0.clone()
    xb = torch.full((M, d), float("nan"), device="cuda", dtype=torch.bfloat16)
    tiles = d // 256
    ssq = torch.full((M, tiles), float("nan"), device="cuda")
    ops.gemm_resid_norm(a, wo, x, xb, ssq, cta_group=cta_group)
    ref_x = ops.gemm(a, wo, ops.EPI_RESID_F32, out=x0.clone(), aux=x0, cta_group=cta_group)
    assert torch.equal(x, ref_x)                                   # same accumulators, same residual add
    assert torch.equal(xb, x.bfloat16())
    ref_ssq = (x.double() ** 2).view(M, tiles, 256).sum(-1)
    assert _rel(ssq, ref_ssq) < 1e-5
    x2 = x0.clone()                                                # deterministic: a second run gives the same bits
    ssq2 = torch.empty_like(ssq)
    ops.gemm_resid_norm(a, wo, x2, torch.empty_like(xb), ssq2, cta_group=cta_group)
    assert torch.equal(ssq2, ssq)
    # --- consumers: norm weight folded into the projection weight's columns
    wn = 1 + 0.1 * torch.randn(d, device="cuda", generator=g)
    rstd = torch.rsqrt((x.double() ** 2).mean(-1, keepdim=True) + eps)
    xn = xb.double() * rstd                                        # what the folded pipeline normalises: bf16(x)
    wg = (torch.randn(F, d, device="cuda", generator=g) / math.sqrt(d))
    wu = (torch.randn(F, d, device="cuda", generator=g) / math.sqrt(d))
    wg_f, wu_f = (wg * wn).bfloat16(), (wu * wn).bfloat16()
    h = ops.gemm_swiglu_rownorm(xb, interleave_gate_up(wg_f, wu_f), ssq, tiles, d, eps, cta_group=cta_group)
    ref_h = torch.nn.functional.silu(xn @ wg_f.double().t()) * (xn @ wu_f.double().t())
    assert _rel(h.double(), ref_h) < 6e-3                          # one bf16 rounding of the result
    # ... and against the un-folded definition (norm in fp32 on x, bf16 weight product, like the reference)
    xn_ref = llada.rms_norm(x.cpu(), wn.cpu(), eps).double().cuda()
    ref_h2 = torch.nn.functional.silu(xn_ref @ wg.bfloat16().double().t()) * (xn_ref @ wu.bfloat16().double().t())
    assert _rel(h.double(), ref_h2) < 2e-2
    wqkv = (torch.randn(3 * d, d, device="cuda", generator=g) / math.sqrt(d))
    wqkv_f = (wqkv * wn).bfloat16()
    sin, cos = llada.rope_tables(L, hd, 500000.0)
    sin_h, cos_h = sin[0, 0, :, :hd // 2].contiguous().cuda(), cos[0, 0, :, :hd // 2].contiguous().cuda()
    qkv = ops.gemm_qkv_rope_rownorm(xb, wqkv_f, sin_h, cos_h, d, hd, L, ssq, tiles, d, eps, cta_group=cta_group)
    # reference: the explicitly normalised operand through the plain fused q|k|v + RoPE GEMM
    ref_qkv = ops.gemm_qkv_rope(xn.float().bfloat16(), wqkv_f, sin_h, cos_h, d, hd, L, cta_group=cta_group)
    assert _rel(qkv.double(), ref_qkv.double()) < 1.5e-2            # two bf16 roundings apart
    # ssq_tiles = 1 (the embedding's layout): total in column 0
    tot = ssq.sum(-1).contiguous()
    qkv1 = ops.gemm_qkv_rope_rownorm(xb, wqkv_f, sin_h, cos_h, d, hd, L, tot, 1, d, eps, cta_group=cta_group)
    assert _rel(qkv1.double(), qkv.double()) < 1e-2


def test_embed_norm():
    from mmada_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    table = torch.randn(1000, 512, device="cuda", generator=g).bfloat16()
    ids = torch.randint(0, 1000, (3, 50), device="cuda", generator=g)
    xb = torch.empty(150, 512, device="cuda", dtype=torch.bfloat16)
    ssq = torch.empty(150, device="cuda")
    out = ops.embed_norm(ids, table, xb, ssq)
    assert torch.equal(out, table[ids.view(-1)].float())
    assert torch.equal(xb, table[ids.view(-1)])
    assert _rel(ssq, (out.double() ** 2).sum(-1)) < 1e-6


def test_gather_rows():
    from mmada_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(2)
    for dtype, d in ((torch.float32, 516), (torch.bfloat16, 4096), (torch.bfloat16, 264)):
        x = torch.randn(700, d, device="cuda", generator=g).to(dtype)
        rows = torch.randint(0, 700, (333,), device="cuda", generator=g).to(torch.int32)
        assert torch.equal(ops.gather_rows(x, rows), x[rows.long()])
        wide = torch.randn(700, 3 * d, device="cuda", generator=g).to(dtype)
        assert torch.equal(ops.gather_rows(wide[:, d:2 * d], rows), wide[rows.long(), d:2 * d])      # strided source


def test_embed():
    from mmada_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    table = torch.randn(1000, 256, device="cuda", generator=g).bfloat16()
    ids = torch.randint(0, 1000, (3, 50), device="cuda", generator=g)
    out = ops.embed(ids, table)
    assert torch.equal(out, table[ids.view(-1)].float())


@pytest.mark.parametrize("hd", [64, 128], ids=["hd64", "hd128"])
@pytest.mark.parametrize("B,H,L", [(1, 2, 128), (2, 3, 387), (2, 2, 1539), (1, 1, 100), (1, 2, 256), (1, 1, 257),
                                   # more work items than CTA pairs: the persistent path (buffer reuse, phase flips)
                                   (2, 40, 700), (3, 60, 130), (5, 32, 1024),
                                   # head_dim 128, B*H >= 148 and L % 256 <= 128: more items than CTAs, one-slot last items
                                   (5, 32, 1539), (2, 80, 300), (4, 40, 640),
                                   # head_dim 128, 1..8 rows behind the last 256-row item: a one-slot item whose first
                                   # warp shares every row between four threads; 9 rows: the thread-per-row path
                                   (2, 3, 264), (3, 5, 773), (1, 2, 513), (2, 2, 521),
                                   # boundaries: one key, one sub-tile +- 1, one tile +- 1, one item +- 1, 33 rows in the
                                   # second slot, a last key sub-tile of 1 / 16 / 17 / 63 keys
                                   (1, 1, 1), (1, 2, 63), (1, 1, 65), (2, 1, 127), (1, 2, 129), (1, 1, 255), (1, 2, 385),
                                   (1, 1, 417), (1, 2, 272), (1, 1, 273), (1, 1, 319)])
def test_attention(hd, B, H, L):
    from mmada_b200 import ops
    d = H * hd
    g = torch.Generator(device="cuda").manual_seed(L + hd)
    qkv = torch.randn(B * L, 3 * d, device="cuda", generator=g).bfloat16()
    qkv[:, :d] *= 2.0       # wider score range so the lazy-rescale path triggers
    out = ops.attention(qkv, B, L, H, hd)
    q, k, v = (qkv[:, i * d:(i + 1) * d].float().view(B, L, H, hd).transpose(1, 2) for i in range(3))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * L, d)
    # P is rounded to bf16 before P.V and the output to bf16: 2^-8 relative steps
    assert _rel(out.float(), ref) < 1.5e-2
    assert float((out.float() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()) < 5e-3


def _sample_case(B, N, C, guidance, seed, frac_known=0.3, temperature=0.7, mask_len_raw=None):
    g = torch.Generator().manual_seed(seed)
    cond = torch.randn(B, N, C, generator=g) * 2
    unc = torch.randn(B, N, C, generator=g) * 2 if guidance > 0 else None
    q = torch.empty(B * N, C).exponential_(1, generator=g)
    u = torch.rand(B, N, generator=g)
    known = torch.full((B, N), 126336, dtype=torch.int64)
    kn = torch.rand(B, N, generator=g) < frac_known
    known[kn] = torch.randint(0, C, (int(kn.sum()),), generator=g)
    if mask_len_raw is None:
        mask_len_raw = float(N // 3)
    return cond, unc, q, u, known, mask_len_raw, temperature


@pytest.mark.parametrize("B,N,C,guidance", [(2, 64, 8192, 3.5), (1, 256, 8192, 0.0), (3, 100, 512, 2.0), (2, 1024, 1024, 3.5)])
def test_t2i_sample_step_bit_exact(B, N, C, guidance):
    from mmada_b200 import ops
    from oracle import denoise
    for seed, frac, T, ml in ((0, 0.0, 1.0, None), (1, 0.3, 0.7, None), (2, 0.9, 0.0, -1.0), (3, 0.5, 0.05, 1e9)):
        cond, unc, q, u, known, mlr, T = _sample_case(B, N, C, guidance, seed, frac, T, ml)
        ref = denoise.t2i_sample_step(cond, unc, guidance, known.clone(), 126336, mlr, T, q, u)
        L = N + 10
        ids = torch.zeros(B, L, dtype=torch.int64, device="cuda")
        kd = known.cuda()
        tickets = torch.zeros(B, dtype=torch.int32, device="cuda")
        sampled, sel, masking = ops.t2i_sample_step(
            cond.cuda().view(B * N, C), None if unc is None else unc.cuda().view(B * N, C), q.cuda(), u.cuda(), kd, ids, 7,
            tickets, guidance, mlr, T, 126336, 126349, want_masking=True)
        assert torch.equal(sampled.cpu(), ref["sampled_ids"])
        assert torch.equal(masking.cpu(), ref["masking"])
        assert torch.equal(kd.cpu(), ref["next_known"])
        exp_ids = torch.where(ref["masking"], 126336, ref["sampled_ids"] + 126349)
        assert torch.equal(ids[:, 7:7 + N].cpu(), exp_ids)
        assert int(tickets.abs().sum()) == 0
        # probabilities: same formula, CPU vs GPU exp/division rounding -> a few ulp
        torch.testing.assert_close(sel.cpu(), ref["selected_probs"], rtol=2e-6, atol=0)


@pytest.mark.parametrize("B,N,C,guidance,frac_known", [(2, 64, 8192, 3.5, 0.3), (3, 100, 512, 2.0, 0.8), (2, 1024, 1024, 0.0, 0.5),
                                                        (1, 256, 8192, 3.5, 0.0)])
def test_t2i_sample_step_on_compact_logits(B, N, C, guidance, frac_known):
    """Logits of the still-masked positions only (compact_masked_rows + slot map) give exactly the step of the full
    logits; the row list names the masked positions of the cond rows, then the uncond rows, padded per row."""
    from mmada_b200 import ops
    cond, unc, q, u, known, ml, temp = _sample_case(B, N, C, guidance, seed=11, frac_known=frac_known)
    cond, q, u = cond.view(B * N, C).cuda(), q.cuda(), u.cuda()
    unc = unc.view(B * N, C).cuda() if unc is not None else None
    tickets = torch.zeros(B, dtype=torch.int32, device="cuda")
    L, off, mask_id = N + 7, 5, 126336
    nb = 2 if unc is not None else 1
    masked = known == mask_id
    cap = min(N, int(masked.sum(1).max()) + 3)            # an upper bound, with some padding
    k_full, k_cmp = known.clone().cuda(), known.clone().cuda()
    rows, slot = ops.compact_masked_rows(k_cmp, L, off, cap, nb, mask_id)
    rows, slot_h = rows.cpu().view(nb, B, cap), slot.cpu()
    for b in range(B):
        pos = masked[b].nonzero().view(-1)
        assert torch.equal(slot_h[b][pos], b * cap + torch.arange(pos.numel(), dtype=torch.int32))
        assert bool((slot_h[b][~masked[b]] == -1).all())
        for r in range(nb):
            assert torch.equal(rows[r, b, :pos.numel()].long(), (r * B + b) * L + off + pos)
            assert bool((rows[r, b, pos.numel():] == (r * B + b) * L + off).all())
    # compact logits: row slot[b, n] holds position (b, n)'s logits, the padding rows garbage
    sel_pos = (slot.view(-1) >= 0).nonzero().view(-1)
    c_cmp = torch.full((B * cap, C), 7.0, device="cuda")
    c_cmp[slot.view(-1)[sel_pos].long()] = cond[sel_pos]
    u_cmp = None
    if unc is not None:
        u_cmp = torch.full((B * cap, C), -3.0, device="cuda")
        u_cmp[slot.view(-1)[sel_pos].long()] = unc[sel_pos]
    a = ops.t2i_sample_step(cond, unc, q, u, k_full, None, 0, tickets, guidance, ml, temp, mask_id, 126349, want_masking=True)
    b_ = ops.t2i_sample_step(c_cmp, u_cmp, q, u, k_cmp, None, 0, tickets, guidance, ml, temp, mask_id, 126349,
                             want_masking=True, slot=slot)
    for x, y in zip(a, b_):
        assert torch.equal(x, y)
    assert torch.equal(k_full, k_cmp)


def test_mask_by_random_topk_golden(golden):
    from mmada_b200 import ops
    gd = golden("sampling")
    probs, u = torch.from_numpy(gd["probs"]).cuda(), torch.from_numpy(gd["u"]).cuda()
    ml = torch.from_numpy(gd["mask_len"])
    for i, T in enumerate(gd["temperatures"]):
        out = ops.mask_by_random_topk(ml, probs, u, float(T))
        assert torch.equal(out.cpu(), torch.from_numpy(gd["masking"][i]))


def test_torch_ops_route_equals_direct_route():
    """torch.ops.mmada_b200.* (dispatcher) and mmada_b200.ops (ctypes) end in the same launchers."""
    import mmada_b200.torch_ops  # noqa: F401
    from mmada_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(300, 512, device="cuda", generator=g).bfloat16()
    w = torch.randn(768, 512, device="cuda", generator=g).bfloat16()
    assert torch.equal(torch.ops.mmada_b200.gemm(a, w, ops.EPI_BF16), ops.gemm(a, w, ops.EPI_BF16))
    x = torch.randn(300, 512, device="cuda", generator=g)
    wt = torch.rand(512, device="cuda", generator=g)
    assert torch.equal(torch.ops.mmada_b200.rmsnorm(x, wt, 1e-5), ops.rmsnorm(x, wt, 1e-5))
    qkv = torch.randn(2 * 150, 3 * 256, device="cuda", generator=g).bfloat16()
    assert torch.equal(torch.ops.mmada_b200.attention(qkv, 2, 150, 2, 128), ops.attention(qkv, 2, 150, 2, 128))
    # an in-place op through the dispatcher: the sampling step mutates known / input_ids like the direct route
    B, N, C = 2, 64, 512
    cond = torch.randn(B * N, C, device="cuda", generator=g)
    q = torch.empty(B * N, C, device="cuda").exponential_(1, generator=g)
    u = torch.rand(B, N, device="cuda", generator=g)
    k1 = torch.full((B, N), 126336, dtype=torch.int64, device="cuda")
    k2 = k1.clone()
    t1, t2 = torch.zeros(B, dtype=torch.int32, device="cuda"), torch.zeros(B, dtype=torch.int32, device="cuda")
    r1 = torch.ops.mmada_b200.t2i_sample_step(cond, None, q, u, k1, None, 0, t1, 0.0, 20.0, 0.5, 126336, 126349)
    r2 = ops.t2i_sample_step(cond, None, q, u, k2, None, 0, t2, 0.0, 20.0, 0.5, 126336, 126349, want_masking=True)
    assert torch.equal(r1[0], r2[0]) and torch.equal(r1[2], r2[2]) and torch.equal(k1, k2) and not torch.equal(k1, torch.full_like(k1, 126336))
    x = torch.randn(300, 512, device="cuda", generator=g)
    x2 = x.clone()
    xb, ssq = torch.empty(300, 512, device="cuda", dtype=torch.bfloat16), torch.empty(300, 2, device="cuda")
    xb2, ssq2 = torch.empty_like(xb), torch.empty_like(ssq)
    w2 = torch.randn(512, 512, device="cuda", generator=g).bfloat16()
    torch.ops.mmada_b200.gemm_resid_norm(a, w2, x, xb, ssq)
    ops.gemm_resid_norm(a, w2, x2, xb2, ssq2)
    assert torch.equal(x, x2) and torch.equal(xb, xb2) and torch.equal(ssq, ssq2)


def test_attention_rejects_bad_arguments():
    """The C entry point returns an error status (raised by the host wrapper) instead of launching: unsupported head
    size, misaligned pointers / pitches, empty problem."""
    from mmada_b200 import ops
    from mmada_b200._lib import MMadaKernelError
    qkv = torch.zeros(64, 3 * 96, device="cuda", dtype=torch.bfloat16)
    with pytest.raises((MMadaKernelError, AssertionError, ValueError)):
        ops.attention(qkv, 1, 64, 1, 96)                         # head_dim 96
    ok = torch.zeros(64, 3 * 128, device="cuda", dtype=torch.bfloat16)
    with pytest.raises((MMadaKernelError, AssertionError, ValueError)):
        ops.attention(ok, 1, 0, 1, 128)                          # L = 0
