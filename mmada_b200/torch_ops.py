"""``torch.ops.mmada_b200.*`` — the C-ABI kernels registered as PyTorch custom ops (``torch.library``), for callers
that want dispatcher-visible, stream-correct, ``torch.compile``-traceable operators (fake/meta kernels included).

The package's own hot path does NOT go through the dispatcher (``mmada_b200.ops`` calls libmmada_b200.so directly with
``ctypes``; the dispatcher costs ~10 us per call, ~230 calls per denoising step); both routes end in the same
``extern "C"`` launchers of include/mmada_b200.h on ``torch.cuda.current_stream()``.  Importing this module registers
the ops; calling one without a CUDA tensor raises (there is no CPU implementation).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch.library import custom_op

from . import ops

_F32_EPI = (ops.EPI_F32, ops.EPI_RESID_F32, ops.EPI_BIAS_F32, ops.EPI_BIAS_RESID_F32)


@custom_op("mmada_b200::gemm", mutates_args=(), device_types="cuda")
def gemm(a: torch.Tensor, w: torch.Tensor, epilogue: int, aux: Optional[torch.Tensor] = None,
         bias: Optional[torch.Tensor] = None, cta_group: int = 2) -> torch.Tensor:
    """epilogue(a[M,K] @ w[N,K]^T) — mmada_gemm_bf16 (replaces nn.Linear, modeling_llada.py:901-930)."""
    return ops.gemm(a, w, epilogue, aux=aux, bias=bias, cta_group=cta_group)


@gemm.register_fake
def _(a, w, epilogue, aux=None, bias=None, cta_group=2):
    n = w.shape[0] // 2 if epilogue == ops.EPI_SWIGLU_BF16 else w.shape[0]
    return a.new_empty((a.shape[0], n), dtype=torch.float32 if epilogue in _F32_EPI else torch.bfloat16)


@custom_op("mmada_b200::gemm_qkv_rope", mutates_args=(), device_types="cuda")
def gemm_qkv_rope(a: torch.Tensor, wqkv: torch.Tensor, sin: torch.Tensor, cos: torch.Tensor, d_model: int, head_dim: int,
                  seq_len: int) -> torch.Tensor:
    """q|k|v projection with NeoX RoPE on q and k in the epilogue — mmada_gemm_qkv_rope_bf16."""
    return ops.gemm_qkv_rope(a, wqkv, sin, cos, d_model, head_dim, seq_len)


@gemm_qkv_rope.register_fake
def _(a, wqkv, sin, cos, d_model, head_dim, seq_len):
    return a.new_empty((a.shape[0], wqkv.shape[0]), dtype=torch.bfloat16)


@custom_op("mmada_b200::attention", mutates_args=(), device_types="cuda")
def attention(qkv: torch.Tensor, batch: int, seq_len: int, n_heads: int, head_dim: int) -> torch.Tensor:
    """Bidirectional attention over the fused [B*L, 3d] projection — mmada_attention_bf16 (modeling_llada.py:653)."""
    return ops.attention(qkv, batch, seq_len, n_heads, head_dim)


@attention.register_fake
def _(qkv, batch, seq_len, n_heads, head_dim):
    return qkv.new_empty((qkv.shape[0], n_heads * head_dim))


@custom_op("mmada_b200::rmsnorm", mutates_args=(), device_types="cuda")
def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """RMSLayerNorm (modeling_llada.py:315-329): fp32 rows in, bf16 out — mmada_rmsnorm_bf16."""
    return ops.rmsnorm(x, weight, eps)


@rmsnorm.register_fake
def _(x, weight, eps):
    return x.new_empty(x.shape, dtype=torch.bfloat16)


@custom_op("mmada_b200::mask_by_random_topk", mutates_args=(), device_types="cuda")
def mask_by_random_topk(mask_len: torch.Tensor, probs: torch.Tensor, u: torch.Tensor, temperature: float) -> torch.Tensor:
    """models/sampling.py:31-36 with explicit uniforms ``u`` — mmada_mask_by_random_topk."""
    return ops.mask_by_random_topk(mask_len, probs, u, temperature)


@mask_by_random_topk.register_fake
def _(mask_len, probs, u, temperature):
    return probs.new_empty(probs.shape, dtype=torch.bool)


@custom_op("mmada_b200::conv_nhwc", mutates_args=(), device_types="cuda")
def conv_nhwc(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, taps: int, epilogue: int,
              resid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """3x3 / 1x1 NHWC convolution as an implicit GEMM — mmada_conv_nhwc_bf16 (MAGVIT-v2 encoder / decoder)."""
    return ops.conv_nhwc(x, weight, bias, taps, epilogue, resid=resid)


@conv_nhwc.register_fake
def _(x, weight, bias, taps, epilogue, resid=None):
    return x.new_empty(x.shape[:3] + (weight.shape[0],), dtype=torch.bfloat16 if epilogue == ops.EPI_BIAS_BF16 else torch.float32)


OPS = ("gemm", "gemm_qkv_rope", "attention", "rmsnorm", "mask_by_random_topk", "conv_nhwc")
