"""Thin Python wrappers: torch tensors in, raw pointers over the C ABI, work enqueued on torch's
current CUDA stream.  PyTorch is used for device memory and streams only."""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib

EPI_BF16, EPI_F32, EPI_RESID_F32, EPI_SWIGLU_BF16, EPI_BIAS_BF16, EPI_BIAS_F32, EPI_BIAS_RESID_F32 = range(7)


#: when set to a list, every GEMM launch appends (start_event, end_event, M, N, K, epilogue) — used by
#: bench.py to time the dominant kernel live on the launching stream
GEMM_EVENTS = None
#: the same for the attention launches: (start_event, end_event, batch, seq_len, n_heads, head_dim)
ATT_EVENTS = None


def _call(tensors, name: str, *args) -> None:
    """Launch ``name`` on the current stream of the operands' device (all operands must share one CUDA device; the
    current device is switched for the call when it is another one)."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.MMadaKernelError(f"{name}: CPU tensor passed (mmada_b200 has no CPU path)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise _lib.MMadaKernelError(f"{name}: operands on different devices ({dev} and {t.device})")
    idx = dev.index
    if idx == torch.cuda.current_device():
        _lib.call(name, *args, torch.cuda.current_stream(idx).cuda_stream)
    else:
        with torch.cuda.device(idx):
            _lib.call(name, *args, torch.cuda.current_stream(idx).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise _lib.MMadaKernelError(f"{name} must be a CUDA tensor (mmada_b200 has no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")


def gemm(a: torch.Tensor, w: torch.Tensor, epilogue: int = EPI_BF16, out: Optional[torch.Tensor] = None,
         aux: Optional[torch.Tensor] = None, cta_group: int = 2, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = epilogue(a[M,K] @ w[N,K]^T); a, w bf16 with contiguous K."""
    _chk(a, torch.bfloat16, "a"); _chk(w, torch.bfloat16, "w")
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1]
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    n_out = N // 2 if epilogue == EPI_SWIGLU_BF16 else N
    if out is None:
        dt = torch.float32 if epilogue in (EPI_F32, EPI_RESID_F32, EPI_BIAS_F32, EPI_BIAS_RESID_F32) else torch.bfloat16
        out = torch.empty((M, n_out), dtype=dt, device=a.device)
    assert out.shape == (M, n_out) and out.stride(1) == 1
    if epilogue in (EPI_RESID_F32, EPI_BIAS_RESID_F32):
        assert aux is not None and aux.dtype == torch.float32 and aux.stride(0) == out.stride(0)
    if epilogue in (EPI_BIAS_BF16, EPI_BIAS_F32, EPI_BIAS_RESID_F32):
        assert bias is not None and bias.dtype == torch.float32 and bias.numel() == N
    ev = GEMM_EVENTS
    if ev is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _call((a, w, out, aux, bias,), "mmada_gemm_bf16", a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0),
              _ptr(aux), _ptr(bias), M, N, K, epilogue, cta_group)
    if ev is not None:
        e1.record()
        ev.append((e0, e1, M, N, K, epilogue))
    return out


def gemm_qkv_rope(a: torch.Tensor, wqkv: torch.Tensor, sin: torch.Tensor, cos: torch.Tensor, d_model: int, head_dim: int,
                  seq_len: int, out: Optional[torch.Tensor] = None, cta_group: int = 2) -> torch.Tensor:
    """Fused q|k|v projection + RoPE on the q and k thirds (see mmada_gemm_qkv_rope_bf16)."""
    _chk(a, torch.bfloat16, "a"); _chk(wqkv, torch.bfloat16, "wqkv"); _chk(sin, torch.float32, "sin"); _chk(cos, torch.float32, "cos")
    M, K = a.shape
    N = wqkv.shape[0]
    assert a.stride(1) == 1 and wqkv.stride(1) == 1 and sin.is_contiguous() and cos.is_contiguous()
    assert sin.shape[-1] == head_dim // 2 and sin.shape[0] >= seq_len
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    ev = GEMM_EVENTS
    if ev is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _call((a, wqkv, out, sin, cos,), "mmada_gemm_qkv_rope_bf16", a.data_ptr(), a.stride(0), wqkv.data_ptr(), wqkv.stride(0), out.data_ptr(),
              out.stride(0), sin.data_ptr(), cos.data_ptr(), M, N, K, 2 * d_model, head_dim, seq_len, cta_group)
    if ev is not None:
        e1.record()
        ev.append((e0, e1, M, N, K, 7))
    return out


def _timed(M: int, N: int, K: int, epi: int):
    """GEMM_EVENTS bracket for the folded-norm GEMM entry points (same record format as gemm())."""
    ev = GEMM_EVENTS
    if ev is None:
        return None, None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    return e0, (e1, M, N, K, epi)


def _timed_end(e0, rest):
    if e0 is not None:
        rest[0].record()
        GEMM_EVENTS.append((e0,) + rest)


def gemm_resid_norm(a: torch.Tensor, w: torch.Tensor, x: torch.Tensor, xb: torch.Tensor, ssq: torch.Tensor,
                    cta_group: int = 2) -> torch.Tensor:
    """x fp32 [M,N] += a @ w^T in place; xb bf16 [M,N] = bf16(x); ssq fp32 [M, N/256] = per-tile row sums of x^2
    (mmada_gemm_resid_norm_f32: the producer half of the folded RMSNorm)."""
    _chk(a, torch.bfloat16, "a"); _chk(w, torch.bfloat16, "w"); _chk(x, torch.float32, "x")
    _chk(xb, torch.bfloat16, "xb"); _chk(ssq, torch.float32, "ssq")
    M, K = a.shape
    N = w.shape[0]
    assert a.stride(1) == 1 and w.stride(1) == 1 and x.shape == (M, N) and xb.shape == (M, N)
    assert x.stride(1) == 1 and xb.stride(1) == 1 and N % 256 == 0 and ssq.is_contiguous() and ssq.numel() >= M * (N // 256)
    e0, rest = _timed(M, N, K, EPI_RESID_F32)
    _call((a, w, x, xb, ssq,), "mmada_gemm_resid_norm_f32", a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), x.data_ptr(), x.stride(0),
              xb.data_ptr(), xb.stride(0), ssq.data_ptr(), M, N, K, cta_group)
    _timed_end(e0, rest)
    return x


def gemm_swiglu_rownorm(a: torch.Tensor, w: torch.Tensor, ssq: torch.Tensor, ssq_tiles: int, norm_dim: int, eps: float,
                        out: Optional[torch.Tensor] = None, cta_group: int = 2) -> torch.Tensor:
    """SwiGLU GEMM whose accumulator rows are scaled by the folded RMSNorm's rstd (mmada_gemm_swiglu_rownorm_bf16)."""
    _chk(a, torch.bfloat16, "a"); _chk(w, torch.bfloat16, "w"); _chk(ssq, torch.float32, "ssq")
    M, K = a.shape
    N = w.shape[0]
    assert a.stride(1) == 1 and w.stride(1) == 1 and ssq.is_contiguous() and ssq.numel() >= M * ssq_tiles
    if out is None:
        out = torch.empty((M, N // 2), dtype=torch.bfloat16, device=a.device)
    assert out.shape == (M, N // 2) and out.stride(1) == 1
    e0, rest = _timed(M, N, K, EPI_SWIGLU_BF16)
    _call((a, w, out, ssq,), "mmada_gemm_swiglu_rownorm_bf16", a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(),
              out.stride(0), ssq.data_ptr(), ssq_tiles, norm_dim, float(eps), M, N, K, cta_group)
    _timed_end(e0, rest)
    return out


def gemm_qkv_rope_rownorm(a: torch.Tensor, wqkv: torch.Tensor, sin: torch.Tensor, cos: torch.Tensor, d_model: int,
                          head_dim: int, seq_len: int, ssq: torch.Tensor, ssq_tiles: int, norm_dim: int, eps: float,
                          out: Optional[torch.Tensor] = None, cta_group: int = 2) -> torch.Tensor:
    """gemm_qkv_rope with the folded RMSNorm's row scaling (mmada_gemm_qkv_rope_rownorm_bf16)."""
    _chk(a, torch.bfloat16, "a"); _chk(wqkv, torch.bfloat16, "wqkv"); _chk(sin, torch.float32, "sin"); _chk(cos, torch.float32, "cos")
    _chk(ssq, torch.float32, "ssq")
    M, K = a.shape
    N = wqkv.shape[0]
    assert a.stride(1) == 1 and wqkv.stride(1) == 1 and sin.is_contiguous() and cos.is_contiguous()
    assert sin.shape[-1] == head_dim // 2 and sin.shape[0] >= seq_len and ssq.is_contiguous() and ssq.numel() >= M * ssq_tiles
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    e0, rest = _timed(M, N, K, 7)
    _call((a, wqkv, out, sin, cos, ssq,), "mmada_gemm_qkv_rope_rownorm_bf16", a.data_ptr(), a.stride(0), wqkv.data_ptr(), wqkv.stride(0),
              out.data_ptr(), out.stride(0), sin.data_ptr(), cos.data_ptr(), ssq.data_ptr(), ssq_tiles, norm_dim, float(eps),
              M, N, K, 2 * d_model, head_dim, seq_len, cta_group)
    _timed_end(e0, rest)
    return out


def embed_norm(ids: torch.Tensor, table: torch.Tensor, xb: torch.Tensor, ssq: torch.Tensor) -> torch.Tensor:
    """embed() that also writes the bf16 copy of the rows and their sums of squares (ssq fp32 [M])."""
    _chk(ids, torch.int64, "ids"); _chk(table, torch.bfloat16, "table"); _chk(xb, torch.bfloat16, "xb"); _chk(ssq, torch.float32, "ssq")
    ids = ids.contiguous().view(-1)
    M, d = ids.numel(), table.shape[1]
    assert xb.shape == (M, d) and xb.is_contiguous() and ssq.numel() >= M
    out = torch.empty((M, d), dtype=torch.float32, device=ids.device)
    _call((ids, table, out, xb, ssq,), "mmada_embed_norm_f32", ids.data_ptr(), table.data_ptr(), out.data_ptr(), xb.data_ptr(), ssq.data_ptr(), M, d,
              table.shape[0])
    return out


def embed(ids: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    _chk(ids, torch.int64, "ids"); _chk(table, torch.bfloat16, "table")
    ids = ids.contiguous().view(-1)
    out = torch.empty((ids.numel(), table.shape[1]), dtype=torch.float32, device=ids.device)
    _call((ids, table, out,), "mmada_embed_f32", ids.data_ptr(), table.data_ptr(), out.data_ptr(), ids.numel(), table.shape[1],
              table.shape[0])
    return out


def gather_rows(x: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
    """x[rows] for a 2-D tensor with contiguous rows (any dtype; rows int32)."""
    _chk(rows, torch.int32, "rows")
    assert x.dim() == 2 and x.stride(1) == 1 and x.is_cuda
    out = torch.empty((rows.numel(), x.shape[1]), dtype=x.dtype, device=x.device)
    _call((x, rows, out,), "mmada_gather_rows", x.data_ptr(), x.stride(0) * x.element_size(), rows.data_ptr(), out.data_ptr(),
              rows.numel(), x.shape[1] * x.element_size())
    return out


def cross_entropy_rows(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
    """Per-row ``F.cross_entropy(logits, labels, ignore_index=..., reduction='none')`` (fp32 [R, V] rows, int64 labels):
    logsumexp(row) - row[label], 0 where the label is ``ignore_index`` (modeling_mmada.py:240-243,253-256,264-267)."""
    _chk(logits, torch.float32, "logits"); _chk(labels, torch.int64, "labels")
    assert logits.dim() == 2 and logits.stride(1) == 1 and labels.numel() == logits.shape[0]
    labels = labels.contiguous().view(-1)
    out = torch.empty((logits.shape[0],), dtype=torch.float32, device=logits.device)
    _call((logits, labels, out,), "mmada_cross_entropy_rows_f32", logits.data_ptr(), logits.stride(0), labels.data_ptr(),
          int(ignore_index), out.data_ptr(), logits.shape[0], logits.shape[1])
    return out


def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float, rows: Optional[torch.Tensor] = None,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _chk(x, torch.float32, "x"); _chk(weight, torch.float32, "weight")
    assert x.dim() == 2 and x.is_contiguous()
    m_out = x.shape[0] if rows is None else rows.numel()
    if rows is not None:
        _chk(rows, torch.int32, "rows")
    if out is None:
        out = torch.empty((m_out, x.shape[1]), dtype=torch.bfloat16, device=x.device)
    _call((x, weight, out, rows,), "mmada_rmsnorm_bf16", x.data_ptr(), weight.data_ptr(), out.data_ptr(), _ptr(rows), m_out, x.shape[1],
              float(eps))
    return out


def rope_inplace(qkv: torch.Tensor, sin: torch.Tensor, cos: torch.Tensor, d_model: int, head_dim: int, seq_len: int):
    _chk(qkv, torch.bfloat16, "qkv"); _chk(sin, torch.float32, "sin"); _chk(cos, torch.float32, "cos")
    assert qkv.dim() == 2 and qkv.stride(1) == 1 and sin.is_contiguous() and cos.is_contiguous()
    assert sin.shape[-1] == head_dim // 2 and sin.shape[0] >= seq_len
    _call((qkv, sin, cos,), "mmada_rope_inplace_bf16", qkv.data_ptr(), qkv.stride(0), sin.data_ptr(), cos.data_ptr(), qkv.shape[0],
              d_model, head_dim, seq_len)
    return qkv


def attention(qkv: torch.Tensor, batch: int, seq_len: int, n_heads: int, head_dim: int,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """qkv bf16 [batch*seq_len, 3*d] (q | k | v), returns [batch*seq_len, d]."""
    _chk(qkv, torch.bfloat16, "qkv")
    d = n_heads * head_dim
    assert qkv.shape == (batch * seq_len, 3 * d) and qkv.stride(1) == 1
    if out is None:
        out = torch.empty((batch * seq_len, d), dtype=torch.bfloat16, device=qkv.device)
    es = qkv.element_size()
    ev = ATT_EVENTS
    if ev is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _call((qkv, out,), "mmada_attention_bf16", qkv.data_ptr(), qkv.data_ptr() + d * es, qkv.data_ptr() + 2 * d * es, qkv.stride(0),
              out.data_ptr(), out.stride(0), batch, seq_len, n_heads, head_dim, 1.0 / math.sqrt(head_dim))
    if ev is not None:
        e1.record()
        ev.append((e0, e1, batch, seq_len, n_heads, head_dim))
    return out


def t2i_sample_step(cond: torch.Tensor, uncond: Optional[torch.Tensor], q: torch.Tensor, u: torch.Tensor,
                    known: torch.Tensor, input_ids: Optional[torch.Tensor], img_off: int, tickets: torch.Tensor,
                    guidance: float, mask_len_raw: float, temperature: float, mask_id: int, text_vocab: int,
                    want_masking: bool = False, want_raw: bool = False, no_remask: bool = False,
                    slot: Optional[torch.Tensor] = None):
    """One fused sampling step (see csrc/sampling.cu).  cond/uncond/q: fp32 [B*N, C]; u fp32 [B, N];
    known int64 [B, N] (updated in place); input_ids int64 [B, L] (image slice updated in place).
    Returns (sampled_ids [B,N] int64, selected_probs [B,N] fp32, masking [B,N] bool or None); with
    ``want_raw`` a 4th element: the raw samples at every position (t2m_generate's return value).
    ``slot`` (int32 [B, N], from compact_masked_rows): cond/uncond are then [B*cap, C], the logits of the still-masked
    positions only, and position (b, n) reads row slot[b, n]."""
    B, N = known.shape
    C = cond.shape[-1]
    if slot is not None:
        return _t2i_sample_step_compact(cond, uncond, q, u, known, input_ids, img_off, tickets, guidance, mask_len_raw,
                                        temperature, mask_id, text_vocab, want_masking, no_remask, slot)
    for t, n in ((cond, "cond"), (q, "q"), (u, "u")):
        _chk(t, torch.float32, n)
        assert t.is_contiguous()
    _chk(known, torch.int64, "known"); _chk(tickets, torch.int32, "tickets")
    assert known.is_contiguous() and cond.numel() == B * N * C and q.numel() == B * N * C and u.numel() == B * N
    if uncond is not None:
        _chk(uncond, torch.float32, "uncond")
        assert uncond.is_contiguous() and uncond.numel() == cond.numel()
    ld_ids = 0
    if input_ids is not None:
        _chk(input_ids, torch.int64, "input_ids")
        assert input_ids.stride(1) == 1 and input_ids.shape[0] == B
        ld_ids = input_ids.stride(0)
    sampled = torch.empty((B, N), dtype=torch.int64, device=cond.device)
    sel = torch.empty((B, N), dtype=torch.float32, device=cond.device)
    masking = torch.empty((B, N), dtype=torch.uint8, device=cond.device) if want_masking else None
    raw = torch.empty((B, N), dtype=torch.int64, device=cond.device) if want_raw else None
    # python scalars reach the tensor op as fp32 in the reference ((1 + g) * cond, g * uncond, T * gumbel)
    _call((cond, uncond, q, u, known, input_ids, sampled, sel, masking, raw, tickets,), "mmada_t2i_sample_step", cond.data_ptr(), _ptr(uncond), q.data_ptr(), u.data_ptr(), known.data_ptr(),
              _ptr(input_ids), ld_ids, img_off, sampled.data_ptr(), sel.data_ptr(), _ptr(masking), _ptr(raw),
              1 if no_remask else 0, tickets.data_ptr(), B, N, C, float(1 + guidance), float(guidance),
              float(mask_len_raw), float(temperature), mask_id, text_vocab)
    if want_raw:
        return sampled, sel, (masking.bool() if want_masking else None), raw
    return sampled, sel, (masking.bool() if want_masking else None)


def _t2i_sample_step_compact(cond, uncond, q, u, known, input_ids, img_off, tickets, guidance, mask_len_raw, temperature,
                             mask_id, text_vocab, want_masking, no_remask, slot):
    B, N = known.shape
    C = cond.shape[-1]
    for t, n in ((cond, "cond"), (q, "q"), (u, "u")):
        _chk(t, torch.float32, n)
        assert t.is_contiguous()
    _chk(known, torch.int64, "known"); _chk(tickets, torch.int32, "tickets"); _chk(slot, torch.int32, "slot")
    assert known.is_contiguous() and slot.is_contiguous() and slot.numel() == B * N
    assert q.numel() == B * N * C and u.numel() == B * N and cond.numel() % (B * C) == 0
    if uncond is not None:
        _chk(uncond, torch.float32, "uncond")
        assert uncond.is_contiguous() and uncond.numel() == cond.numel()
    ld_ids = 0
    if input_ids is not None:
        _chk(input_ids, torch.int64, "input_ids")
        assert input_ids.stride(1) == 1 and input_ids.shape[0] == B
        ld_ids = input_ids.stride(0)
    sampled = torch.empty((B, N), dtype=torch.int64, device=cond.device)
    sel = torch.empty((B, N), dtype=torch.float32, device=cond.device)
    masking = torch.empty((B, N), dtype=torch.uint8, device=cond.device) if want_masking else None
    _call((cond, uncond, q, u, known, input_ids, sampled, sel, masking, tickets, slot,), "mmada_t2i_sample_step_compact", cond.data_ptr(), _ptr(uncond), q.data_ptr(), u.data_ptr(), known.data_ptr(),
              _ptr(input_ids), ld_ids, img_off, sampled.data_ptr(), sel.data_ptr(), _ptr(masking),
              1 if no_remask else 0, tickets.data_ptr(), B, N, C, float(1 + guidance), float(guidance),
              float(mask_len_raw), float(temperature), mask_id, text_vocab, slot.data_ptr())
    return sampled, sel, (masking.bool() if want_masking else None)


def compact_masked_rows(known: torch.Tensor, L: int, img_off: int, cap: int, branches: int, mask_id: int):
    """(rows int32 [branches*B*cap], slot int32 [B, N]) of the still-masked positions (csrc/sampling.cu,
    mmada_compact_masked_rows): the token rows to run the last block / ln_f / output head on, and where each
    position's logits then live."""
    _chk(known, torch.int64, "known")
    B, N = known.shape
    assert known.is_contiguous()
    rows = torch.empty(branches * B * cap, dtype=torch.int32, device=known.device)
    slot = torch.empty((B, N), dtype=torch.int32, device=known.device)
    _call((known, rows, slot,), "mmada_compact_masked_rows", known.data_ptr(), rows.data_ptr(), slot.data_ptr(), B, N, L, img_off, cap,
              branches, mask_id)
    return rows, slot


def mask_by_random_topk(mask_len: torch.Tensor, probs: torch.Tensor, u: torch.Tensor, temperature: float) -> torch.Tensor:
    _chk(probs, torch.float32, "probs"); _chk(u, torch.float32, "u")
    B, N = probs.shape
    ml = mask_len.to(device=probs.device).long().reshape(-1).contiguous()
    assert ml.numel() == B
    out = torch.empty((B, N), dtype=torch.uint8, device=probs.device)
    _call((probs, u, ml, out,), "mmada_mask_by_random_topk", probs.contiguous().data_ptr(), u.contiguous().data_ptr(), ml.data_ptr(),
              out.data_ptr(), B, N, float(temperature))
    return out.bool()


def text_sample_rows(logits: torch.Tensor, un_logits: Optional[torch.Tensor], cfg_scale: float, temperature: float,
                     u_noise: Optional[torch.Tensor] = None, seed: int = 0):
    """logits fp32 [R, V] -> (x0 int64 [R], conf fp64 [R]); see csrc/textsample.cu."""
    _chk(logits, torch.float32, "logits")
    assert logits.dim() == 2 and logits.is_contiguous()
    R, V = logits.shape
    if un_logits is not None:
        _chk(un_logits, torch.float32, "un_logits")
        assert un_logits.is_contiguous() and un_logits.shape == logits.shape
    if u_noise is not None:
        _chk(u_noise, torch.float64, "u_noise")
        assert u_noise.is_contiguous() and u_noise.shape == logits.shape
    x0 = torch.empty(R, dtype=torch.int64, device=logits.device)
    conf = torch.empty(R, dtype=torch.float64, device=logits.device)
    _call((logits, un_logits, u_noise, x0, conf,), "mmada_text_sample_rows", logits.data_ptr(), _ptr(un_logits), float(cfg_scale + 1), _ptr(u_noise),
              int(seed) & 0xFFFFFFFFFFFFFFFF, float(temperature), R, V, x0.data_ptr(), conf.data_ptr())
    return x0, conf


def block_mask_count(x: torch.Tensor, lo: int, block: int, mask_id: int) -> torch.Tensor:
    _chk(x, torch.int64, "x")
    assert x.dim() == 2 and x.stride(1) == 1
    cnt = torch.empty(x.shape[0], dtype=torch.int32, device=x.device)
    _call((x, cnt,), "mmada_block_mask_count", x.data_ptr(), x.stride(0), lo, block, x.shape[0], mask_id, cnt.data_ptr())
    return cnt


def text_transfer(x: torch.Tensor, lo: int, block: int, x0: torch.Tensor, conf: Optional[torch.Tensor],
                  cnt: torch.Tensor, steps: int, step: int, mask_id: int, conf_override: Optional[torch.Tensor] = None,
                  want_transfer: bool = False):
    _chk(x, torch.int64, "x"); _chk(x0, torch.int64, "x0"); _chk(cnt, torch.int32, "cnt")
    B = x.shape[0]
    assert x0.numel() == B * block and x0.is_contiguous()
    if conf is not None:
        _chk(conf, torch.float64, "conf")
    if conf_override is not None:
        _chk(conf_override, torch.float64, "conf_override")
        assert conf_override.is_contiguous() and conf_override.numel() == B * block
    tr = torch.empty((B, block), dtype=torch.uint8, device=x.device) if want_transfer else None
    _call((x, x0, conf, conf_override, cnt, tr,), "mmada_text_transfer", x.data_ptr(), x.stride(0), lo, block, x0.data_ptr(), _ptr(conf), _ptr(conf_override),
              cnt.data_ptr(), steps, step, B, mask_id, _ptr(tr))
    return tr.bool() if want_transfer else None


# ---- MAGVIT-v2 decoder ops (NHWC) ---------------------------------------------------------------
def conv_nhwc(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, taps: int, epilogue: int = EPI_BIAS_F32,
              resid: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x bf16 [B,H,W,Cin]; weight bf16 [Cout, taps*Cin]; bias fp32 [Cout] -> [B,H,W,Cout]."""
    _chk(x, torch.bfloat16, "x"); _chk(weight, torch.bfloat16, "weight"); _chk(bias, torch.float32, "bias")
    B, H, W, Cin = x.shape
    Cout = weight.shape[0]
    assert x.is_contiguous() and weight.is_contiguous() and weight.shape[1] == taps * Cin
    if out is None:
        out = torch.empty((B, H, W, Cout), device=x.device, dtype=torch.bfloat16 if epilogue == EPI_BIAS_BF16 else torch.float32)
    if resid is not None:
        _chk(resid, torch.float32, "resid")
        assert resid.is_contiguous() and resid.shape == out.shape
    _call((x, weight, bias, out, resid,), "mmada_conv_nhwc_bf16", x.data_ptr(), weight.data_ptr(), bias.data_ptr(), out.data_ptr(), _ptr(resid), B, H, W,
              Cin, Cout, taps, epilogue)
    return out


def lfq_decode_nhwc(indices: torch.Tensor, pq_weight: torch.Tensor, pq_bias: torch.Tensor, h: int, w: int) -> torch.Tensor:
    _chk(indices, torch.int64, "indices")
    B = indices.shape[0]
    out = torch.empty((B, h, w, 64), device=indices.device, dtype=torch.bfloat16)
    _call((indices, pq_weight, pq_bias, out,), "mmada_lfq_decode_nhwc", indices.contiguous().data_ptr(), pq_weight.data_ptr(), pq_bias.data_ptr(),
              out.data_ptr(), indices.numel())
    return out


def lfq_indices_to_bits(indices: torch.Tensor) -> torch.Tensor:
    _chk(indices, torch.int64, "indices")
    B, N = indices.shape
    out = torch.empty((B, 13, N), device=indices.device, dtype=torch.float32)
    _call((indices, out,), "mmada_lfq_indices_to_bits", indices.contiguous().data_ptr(), out.data_ptr(), B, N)
    return out


def lfq_bits_to_indices(z: torch.Tensor) -> torch.Tensor:
    _chk(z, torch.float32, "z")
    B, N = z.shape[0], z[0, 0].numel()
    out = torch.empty((B, N), device=z.device, dtype=torch.int64)
    _call((z, out,), "mmada_lfq_bits_to_indices", z.contiguous().data_ptr(), out.data_ptr(), B, N)
    return out


def groupnorm_swish(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, sums: torch.Tensor, swish: bool = True,
                    eps: float = 1e-6) -> torch.Tensor:
    """x fp32 NHWC [B,H,W,C] -> bf16 NHWC; sums: fp64 scratch [B,32,2]."""
    _chk(x, torch.float32, "x")
    B, H, W, C = x.shape
    assert x.is_contiguous() and sums.dtype == torch.float64 and sums.numel() >= B * 64
    out = torch.empty((B, H, W, C), device=x.device, dtype=torch.bfloat16)
    _call((x, sums,), "mmada_groupnorm_stats", x.data_ptr(), sums.data_ptr(), B, H * W, C)
    _call((x, sums, gamma, beta, out,), "mmada_groupnorm_apply_bf16", x.data_ptr(), sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(),
              B, H * W, C, float(eps), 1 if swish else 0)
    return out


def upsample2x_nhwc(x: torch.Tensor) -> torch.Tensor:
    _chk(x, torch.float32, "x")
    B, H, W, C = x.shape
    out = torch.empty((B, 2 * H, 2 * W, C), device=x.device, dtype=torch.bfloat16)
    _call((x, out,), "mmada_upsample2x_nhwc_bf16", x.data_ptr(), out.data_ptr(), B, H, W, C)
    return out


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    _chk(x, torch.float32, "x")
    out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    _call((x, out,), "mmada_cast_f32_bf16", x.contiguous().data_ptr(), out.data_ptr(), x.numel())
    return out


def softmax_rows_bf16(x: torch.Tensor, scale: float) -> torch.Tensor:
    _chk(x, torch.float32, "x")
    R, n = x.shape
    out = torch.empty((R, n), device=x.device, dtype=torch.bfloat16)
    _call((x, out,), "mmada_softmax_rows_bf16", x.contiguous().data_ptr(), out.data_ptr(), R, n, float(scale))
    return out


def nhwc_to_nchw(x: torch.Tensor) -> torch.Tensor:
    _chk(x, torch.float32, "x")
    B, H, W, C = x.shape
    out = torch.empty((B, C, H, W), device=x.device, dtype=torch.float32)
    _call((x, out,), "mmada_nhwc_to_nchw_f32", x.contiguous().data_ptr(), out.data_ptr(), B, H * W, C)
    return out


def image_to_uint8(x: torch.Tensor) -> torch.Tensor:
    _chk(x, torch.float32, "x")
    out = torch.empty(x.shape, device=x.device, dtype=torch.uint8)
    _call((x, out,), "mmada_image_to_uint8", x.contiguous().data_ptr(), out.data_ptr(), x.numel())
    return out


def image_to_nhwc64(pixel_values: torch.Tensor) -> torch.Tensor:
    """fp32 NCHW [B,3,H,W] -> bf16 NHWC [B,H,W,64] (channels 3.. zero)."""
    _chk(pixel_values, torch.float32, "pixel_values")
    B, C, H, W = pixel_values.shape
    assert C == 3
    out = torch.empty((B, H, W, 64), device=pixel_values.device, dtype=torch.bfloat16)
    _call((pixel_values, out,), "mmada_image_to_nhwc64_bf16", pixel_values.contiguous().data_ptr(), out.data_ptr(), B, H, W)
    return out


def space_to_depth2(x: torch.Tensor) -> torch.Tensor:
    """fp32 NHWC [B,H,W,C] -> bf16 NHWC [B,H/2,W/2,4C], channel (2*sy+sx)*C + c = pixel (2y+sy, 2x+sx)."""
    _chk(x, torch.float32, "x")
    B, H, W, C = x.shape
    assert x.is_contiguous()
    out = torch.empty((B, H // 2, W // 2, 4 * C), device=x.device, dtype=torch.bfloat16)
    _call((x, out,), "mmada_space_to_depth2_bf16", x.data_ptr(), out.data_ptr(), B, H, W, C)
    return out


def conv1d_gather(x: torch.Tensor, taps: int, dilation: int = 1, upsample: int = 1, relu: bool = False) -> torch.Tensor:
    """x fp32 [B,T,C] -> bf16 [B, T*upsample, taps*C] (see mmada_conv1d_gather_bf16)."""
    _chk(x, torch.float32, "x")
    B, T, C = x.shape
    assert x.is_contiguous()
    out = torch.empty((B, T * upsample, taps * C), device=x.device, dtype=torch.bfloat16)
    _call((x, out,), "mmada_conv1d_gather_bf16", x.data_ptr(), out.data_ptr(), B, T, C, taps, dilation, upsample, 1 if relu else 0)
    return out


def relu_(x: torch.Tensor) -> torch.Tensor:
    _chk(x, torch.float32, "x")
    assert x.is_contiguous()
    _call((x,), "mmada_relu_f32", x.data_ptr(), x.numel())
    return x


def build_prompts(text: torch.Tensor, text_off: torch.Tensor, body: torch.Tensor, text_slots: int, mode: int, task: int,
                  bos: int, eos: int, pad: int, open_tok: int, close_tok: int, end_header: int = -1):
    """Sequence assembly on the device (csrc/prompting.cu, mmada_build_prompts).  text int64 [sum of lengths], text_off
    int64 [B+1], body int64 [B, N].  mode 0 -> (ids [B, text_slots+N+2], attention mask [B, L]); mode 1 -> (ids
    [B, 3+N+text_slots], prompt lengths [B])."""
    _chk(text, torch.int64, "text"); _chk(text_off, torch.int64, "text_off"); _chk(body, torch.int64, "body")
    assert body.dim() == 2 and body.stride(1) == 1 and text.is_contiguous() and text_off.is_contiguous()
    B, N = body.shape
    assert text_off.numel() == B + 1
    L = text_slots + N + 2 if mode == 0 else 3 + N + text_slots
    ids = torch.empty((B, L), dtype=torch.int64, device=body.device)
    mask = torch.empty((B, L) if mode == 0 else (B,), dtype=torch.int64, device=body.device)
    _call((text, text_off, body, ids, mask,), "mmada_build_prompts", text.data_ptr(), text_off.data_ptr(), body.data_ptr(),
          body.stride(0), ids.data_ptr(), mask.data_ptr(), B, N, int(text_slots), int(mode), int(task), int(bos), int(eos),
          int(pad), int(open_tok), int(close_tok), int(end_header))
    return ids, mask
