"""``torch.ops.mmada_b200.*`` — the C-ABI kernels registered as PyTorch custom ops (``torch.library``), for callers
that want dispatcher-visible, stream-correct, ``torch.compile``-traceable operators (fake/meta kernels included).

Every launcher of include/mmada_b200.h is reachable through an op here (``OPS`` maps op -> entry points), in-place
launchers with ``mutates_args``.  The package's own hot path does NOT go through the dispatcher (``mmada_b200.ops`` calls
libmmada_b200.so directly with ``ctypes``; the dispatcher costs ~10 us per call, ~170 calls per denoising step); both
routes end in the same ``extern "C"`` launchers on the current stream of the operands' device.  Importing this module registers
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


# ---- folded-RMSNorm GEMMs, embeddings, row gathers ---------------------------------------------------------------
@custom_op("mmada_b200::gemm_resid_norm", mutates_args=("x", "xb", "ssq"), device_types="cuda")
def gemm_resid_norm(a: torch.Tensor, w: torch.Tensor, x: torch.Tensor, xb: torch.Tensor, ssq: torch.Tensor,
                    cta_group: int = 2) -> None:
    """x += a @ w^T in place, xb = bf16(x), ssq = per-tile row sums of x^2 — mmada_gemm_resid_norm_f32."""
    ops.gemm_resid_norm(a, w, x, xb, ssq, cta_group=cta_group)


@custom_op("mmada_b200::gemm_swiglu_rownorm", mutates_args=(), device_types="cuda")
def gemm_swiglu_rownorm(a: torch.Tensor, w: torch.Tensor, ssq: torch.Tensor, ssq_tiles: int, norm_dim: int, eps: float,
                        cta_group: int = 2) -> torch.Tensor:
    """silu(r*gate) * (r*up) with the folded RMSNorm's row factor r — mmada_gemm_swiglu_rownorm_bf16."""
    return ops.gemm_swiglu_rownorm(a, w, ssq, ssq_tiles, norm_dim, eps, cta_group=cta_group)


@gemm_swiglu_rownorm.register_fake
def _(a, w, ssq, ssq_tiles, norm_dim, eps, cta_group=2):
    return a.new_empty((a.shape[0], w.shape[0] // 2), dtype=torch.bfloat16)


@custom_op("mmada_b200::gemm_qkv_rope_rownorm", mutates_args=(), device_types="cuda")
def gemm_qkv_rope_rownorm(a: torch.Tensor, wqkv: torch.Tensor, sin: torch.Tensor, cos: torch.Tensor, d_model: int,
                          head_dim: int, seq_len: int, ssq: torch.Tensor, ssq_tiles: int, norm_dim: int, eps: float,
                          cta_group: int = 2) -> torch.Tensor:
    """q|k|v projection + RoPE with the folded RMSNorm's row factor — mmada_gemm_qkv_rope_rownorm_bf16."""
    return ops.gemm_qkv_rope_rownorm(a, wqkv, sin, cos, d_model, head_dim, seq_len, ssq, ssq_tiles, norm_dim, eps,
                                     cta_group=cta_group)


@gemm_qkv_rope_rownorm.register_fake
def _(a, wqkv, sin, cos, d_model, head_dim, seq_len, ssq, ssq_tiles, norm_dim, eps, cta_group=2):
    return a.new_empty((a.shape[0], wqkv.shape[0]), dtype=torch.bfloat16)


@custom_op("mmada_b200::embed", mutates_args=(), device_types="cuda")
def embed(ids: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    """wte(input_ids) (modeling_llada.py:1222): bf16 rows -> fp32 — mmada_embed_f32."""
    return ops.embed(ids, table)


@embed.register_fake
def _(ids, table):
    return table.new_empty((ids.numel(), table.shape[1]), dtype=torch.float32)


@custom_op("mmada_b200::embed_norm", mutates_args=("xb", "ssq"), device_types="cuda")
def embed_norm(ids: torch.Tensor, table: torch.Tensor, xb: torch.Tensor, ssq: torch.Tensor) -> torch.Tensor:
    """embed that also writes the bf16 rows and their sums of squares — mmada_embed_norm_f32."""
    return ops.embed_norm(ids, table, xb, ssq)


@embed_norm.register_fake
def _(ids, table, xb, ssq):
    return table.new_empty((ids.numel(), table.shape[1]), dtype=torch.float32)


@custom_op("mmada_b200::gather_rows", mutates_args=(), device_types="cuda")
def gather_rows(x: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
    """x[rows] for a 2-D tensor — mmada_gather_rows."""
    return ops.gather_rows(x, rows)


@gather_rows.register_fake
def _(x, rows):
    return x.new_empty((rows.numel(), x.shape[1]))


@custom_op("mmada_b200::cross_entropy_rows", mutates_args=(), device_types="cuda")
def cross_entropy_rows(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
    """F.cross_entropy(..., reduction='none') per row (modeling_mmada.py:240-267) — mmada_cross_entropy_rows_f32."""
    return ops.cross_entropy_rows(logits, labels, ignore_index)


@cross_entropy_rows.register_fake
def _(logits, labels, ignore_index=-100):
    return logits.new_empty((logits.shape[0],), dtype=torch.float32)


@custom_op("mmada_b200::rope_inplace", mutates_args=("qkv",), device_types="cuda")
def rope_inplace(qkv: torch.Tensor, sin: torch.Tensor, cos: torch.Tensor, d_model: int, head_dim: int, seq_len: int) -> None:
    """RotaryEmbedding (modeling_llada.py:402-428) in place on the q and k thirds — mmada_rope_inplace_bf16."""
    ops.rope_inplace(qkv, sin, cos, d_model, head_dim, seq_len)


# ---- sampling ------------------------------------------------------------------------------------------------------
@custom_op("mmada_b200::t2i_sample_step", mutates_args=("known", "input_ids", "tickets"), device_types="cuda")
def t2i_sample_step(cond: torch.Tensor, uncond: Optional[torch.Tensor], q: torch.Tensor, u: torch.Tensor,
                    known: torch.Tensor, input_ids: Optional[torch.Tensor], img_off: int, tickets: torch.Tensor,
                    guidance: float, mask_len_raw: float, temperature: float, mask_id: int, text_vocab: int,
                    no_remask: bool = False, slot: Optional[torch.Tensor] = None) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """One fused t2i sampling step (modeling_mmada.py:164-209 + sampling.py:31-36) — mmada_t2i_sample_step, or
    mmada_t2i_sample_step_compact when ``slot`` maps positions to the rows of compacted logits.  Updates ``known`` and the
    image slice of ``input_ids`` in place; returns (sampled_ids, selected_probs, masking)."""
    sampled, sel, masking = ops.t2i_sample_step(cond, uncond, q, u, known, input_ids, img_off, tickets, guidance, mask_len_raw,
                                                temperature, mask_id, text_vocab, want_masking=True, no_remask=no_remask,
                                                slot=slot)
    return sampled, sel, masking


@t2i_sample_step.register_fake
def _(cond, uncond, q, u, known, input_ids, img_off, tickets, guidance, mask_len_raw, temperature, mask_id, text_vocab,
      no_remask=False, slot=None):
    return (known.new_empty(known.shape), known.new_empty(known.shape, dtype=torch.float32),
            known.new_empty(known.shape, dtype=torch.bool))


@custom_op("mmada_b200::compact_masked_rows", mutates_args=(), device_types="cuda")
def compact_masked_rows(known: torch.Tensor, L: int, img_off: int, cap: int, branches: int,
                        mask_id: int) -> tuple[torch.Tensor, torch.Tensor]:
    """Token rows of the still-masked positions and the position -> slot map — mmada_compact_masked_rows."""
    return ops.compact_masked_rows(known, L, img_off, cap, branches, mask_id)


@compact_masked_rows.register_fake
def _(known, L, img_off, cap, branches, mask_id):
    return (known.new_empty((branches * known.shape[0] * cap,), dtype=torch.int32), known.new_empty(known.shape, dtype=torch.int32))


@custom_op("mmada_b200::text_sample_rows", mutates_args=(), device_types="cuda")
def text_sample_rows(logits: torch.Tensor, un_logits: Optional[torch.Tensor], cfg_scale: float, temperature: float,
                     u_noise: Optional[torch.Tensor] = None, seed: int = 0) -> tuple[torch.Tensor, torch.Tensor]:
    """Gumbel-max token + fp64 softmax confidence per candidate row (generate.py:8-19,86-96) — mmada_text_sample_rows."""
    return ops.text_sample_rows(logits, un_logits, cfg_scale, temperature, u_noise, seed=seed)


@text_sample_rows.register_fake
def _(logits, un_logits, cfg_scale, temperature, u_noise=None, seed=0):
    return (logits.new_empty((logits.shape[0],), dtype=torch.int64), logits.new_empty((logits.shape[0],), dtype=torch.float64))


@custom_op("mmada_b200::block_mask_count", mutates_args=(), device_types="cuda")
def block_mask_count(x: torch.Tensor, lo: int, block: int, mask_id: int) -> torch.Tensor:
    """Masked positions per sequence inside the block (get_num_transfer_tokens' input) — mmada_block_mask_count."""
    return ops.block_mask_count(x, lo, block, mask_id)


@block_mask_count.register_fake
def _(x, lo, block, mask_id):
    return x.new_empty((x.shape[0],), dtype=torch.int32)


@custom_op("mmada_b200::text_transfer", mutates_args=("x",), device_types="cuda")
def text_transfer(x: torch.Tensor, lo: int, block: int, x0: torch.Tensor, conf: Optional[torch.Tensor], cnt: torch.Tensor,
                  steps: int, step: int, mask_id: int, conf_override: Optional[torch.Tensor] = None) -> None:
    """Commit the k most confident masked positions of the block (generate.py:98-111) — mmada_text_transfer."""
    ops.text_transfer(x, lo, block, x0, conf, cnt, steps, step, mask_id, conf_override=conf_override)


# ---- MAGVIT-v2 / motion decoder element-wise kernels, prompt assembly ----------------------------------------------------
@custom_op("mmada_b200::lfq_decode_nhwc", mutates_args=(), device_types="cuda")
def lfq_decode_nhwc(indices: torch.Tensor, pq_weight: torch.Tensor, pq_bias: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """code ids -> +-1 bits -> post_quant_conv, NHWC bf16 padded to 64 channels — mmada_lfq_decode_nhwc."""
    return ops.lfq_decode_nhwc(indices, pq_weight, pq_bias, h, w)


@lfq_decode_nhwc.register_fake
def _(indices, pq_weight, pq_bias, h, w):
    return pq_weight.new_empty((indices.shape[0], h, w, 64), dtype=torch.bfloat16)


@custom_op("mmada_b200::lfq_indices_to_bits", mutates_args=(), device_types="cuda")
def lfq_indices_to_bits(indices: torch.Tensor) -> torch.Tensor:
    """LFQuantizer.get_codebook_entry (modeling_magvitv2.py:208-221) — mmada_lfq_indices_to_bits."""
    return ops.lfq_indices_to_bits(indices)


@lfq_indices_to_bits.register_fake
def _(indices):
    return indices.new_empty((indices.shape[0], 13, indices.shape[1]), dtype=torch.float32)


@custom_op("mmada_b200::lfq_bits_to_indices", mutates_args=(), device_types="cuda")
def lfq_bits_to_indices(z: torch.Tensor) -> torch.Tensor:
    """LFQuantizer.get_indices (:201-206) — mmada_lfq_bits_to_indices."""
    return ops.lfq_bits_to_indices(z)


@lfq_bits_to_indices.register_fake
def _(z):
    return z.new_empty((z.shape[0], z[0, 0].numel()), dtype=torch.int64)


@custom_op("mmada_b200::groupnorm_swish", mutates_args=("sums",), device_types="cuda")
def groupnorm_swish(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, sums: torch.Tensor, swish: bool = True,
                    eps: float = 1e-6) -> torch.Tensor:
    """GroupNorm(32) (+ swish), fp32 NHWC -> bf16 — mmada_groupnorm_stats + mmada_groupnorm_apply_bf16."""
    return ops.groupnorm_swish(x, gamma, beta, sums, swish, eps)


@groupnorm_swish.register_fake
def _(x, gamma, beta, sums, swish=True, eps=1e-6):
    return x.new_empty(x.shape, dtype=torch.bfloat16)


@custom_op("mmada_b200::upsample2x_nhwc", mutates_args=(), device_types="cuda")
def upsample2x_nhwc(x: torch.Tensor) -> torch.Tensor:
    """Nearest 2x upsample, fp32 NHWC -> bf16 — mmada_upsample2x_nhwc_bf16."""
    return ops.upsample2x_nhwc(x)


@upsample2x_nhwc.register_fake
def _(x):
    return x.new_empty((x.shape[0], 2 * x.shape[1], 2 * x.shape[2], x.shape[3]), dtype=torch.bfloat16)


@custom_op("mmada_b200::cast_bf16", mutates_args=(), device_types="cuda")
def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> bf16 — mmada_cast_f32_bf16."""
    return ops.cast_bf16(x)


@cast_bf16.register_fake
def _(x):
    return x.new_empty(x.shape, dtype=torch.bfloat16)


@custom_op("mmada_b200::softmax_rows_bf16", mutates_args=(), device_types="cuda")
def softmax_rows_bf16(x: torch.Tensor, scale: float) -> torch.Tensor:
    """Row softmax of the VQ AttnBlock, fp32 -> bf16 — mmada_softmax_rows_bf16."""
    return ops.softmax_rows_bf16(x, scale)


@softmax_rows_bf16.register_fake
def _(x, scale):
    return x.new_empty(x.shape, dtype=torch.bfloat16)


@custom_op("mmada_b200::nhwc_to_nchw", mutates_args=(), device_types="cuda")
def nhwc_to_nchw(x: torch.Tensor) -> torch.Tensor:
    """fp32 NHWC -> NCHW — mmada_nhwc_to_nchw_f32."""
    return ops.nhwc_to_nchw(x)


@nhwc_to_nchw.register_fake
def _(x):
    return x.new_empty((x.shape[0], x.shape[3], x.shape[1], x.shape[2]))


@custom_op("mmada_b200::image_to_uint8", mutates_args=(), device_types="cuda")
def image_to_uint8(x: torch.Tensor) -> torch.Tensor:
    """clamp((x + 1) / 2, 0, 1) * 255 -> uint8 (inference_t2i.py:123-125) — mmada_image_to_uint8."""
    return ops.image_to_uint8(x)


@image_to_uint8.register_fake
def _(x):
    return x.new_empty(x.shape, dtype=torch.uint8)


@custom_op("mmada_b200::image_to_nhwc64", mutates_args=(), device_types="cuda")
def image_to_nhwc64(pixel_values: torch.Tensor) -> torch.Tensor:
    """fp32 NCHW image -> bf16 NHWC padded to 64 channels — mmada_image_to_nhwc64_bf16."""
    return ops.image_to_nhwc64(pixel_values)


@image_to_nhwc64.register_fake
def _(pixel_values):
    B, _, H, W = pixel_values.shape
    return pixel_values.new_empty((B, H, W, 64), dtype=torch.bfloat16)


@custom_op("mmada_b200::space_to_depth2", mutates_args=(), device_types="cuda")
def space_to_depth2(x: torch.Tensor) -> torch.Tensor:
    """fp32 NHWC -> bf16 [B, H/2, W/2, 4C] (stride-2 Downsample as a stride-1 conv) — mmada_space_to_depth2_bf16."""
    return ops.space_to_depth2(x)


@space_to_depth2.register_fake
def _(x):
    return x.new_empty((x.shape[0], x.shape[1] // 2, x.shape[2] // 2, 4 * x.shape[3]), dtype=torch.bfloat16)


@custom_op("mmada_b200::conv1d_gather", mutates_args=(), device_types="cuda")
def conv1d_gather(x: torch.Tensor, taps: int, dilation: int = 1, upsample: int = 1, relu: bool = False) -> torch.Tensor:
    """Conv1d operand gather of the motion decoder — mmada_conv1d_gather_bf16."""
    return ops.conv1d_gather(x, taps, dilation, upsample, relu)


@conv1d_gather.register_fake
def _(x, taps, dilation=1, upsample=1, relu=False):
    return x.new_empty((x.shape[0], x.shape[1] * upsample, taps * x.shape[2]), dtype=torch.bfloat16)


@custom_op("mmada_b200::relu_", mutates_args=("x",), device_types="cuda")
def relu_(x: torch.Tensor) -> None:
    """In-place ReLU — mmada_relu_f32."""
    ops.relu_(x)


@custom_op("mmada_b200::build_prompts", mutates_args=(), device_types="cuda")
def build_prompts(text: torch.Tensor, text_off: torch.Tensor, body: torch.Tensor, text_slots: int, mode: int, task: int,
                  bos: int, eos: int, pad: int, open_tok: int, close_tok: int,
                  end_header: int = -1) -> tuple[torch.Tensor, torch.Tensor]:
    """UniversalPrompting's generation layouts on the device — mmada_build_prompts."""
    return ops.build_prompts(text, text_off, body, text_slots, mode, task, bos, eos, pad, open_tok, close_tok, end_header)


@build_prompts.register_fake
def _(text, text_off, body, text_slots, mode, task, bos, eos, pad, open_tok, close_tok, end_header=-1):
    B, N = body.shape
    L = text_slots + N + 2 if mode == 0 else 3 + N + text_slots
    return body.new_empty((B, L)), body.new_empty((B, L) if mode == 0 else (B,))


#: op name -> the C-ABI entry points (include/mmada_b200.h) it reaches; tests/test_abi_and_host.py checks that every
#: launcher of the header is covered
OPS = {
    "gemm": ("mmada_gemm_bf16",), "gemm_qkv_rope": ("mmada_gemm_qkv_rope_bf16",), "attention": ("mmada_attention_bf16",),
    "rmsnorm": ("mmada_rmsnorm_bf16",), "mask_by_random_topk": ("mmada_mask_by_random_topk",),
    "conv_nhwc": ("mmada_conv_nhwc_bf16",), "gemm_resid_norm": ("mmada_gemm_resid_norm_f32",),
    "gemm_swiglu_rownorm": ("mmada_gemm_swiglu_rownorm_bf16",), "gemm_qkv_rope_rownorm": ("mmada_gemm_qkv_rope_rownorm_bf16",),
    "embed": ("mmada_embed_f32",), "embed_norm": ("mmada_embed_norm_f32",), "gather_rows": ("mmada_gather_rows",),
    "rope_inplace": ("mmada_rope_inplace_bf16",), "cross_entropy_rows": ("mmada_cross_entropy_rows_f32",),
    "t2i_sample_step": ("mmada_t2i_sample_step", "mmada_t2i_sample_step_compact"),
    "compact_masked_rows": ("mmada_compact_masked_rows",), "text_sample_rows": ("mmada_text_sample_rows",),
    "block_mask_count": ("mmada_block_mask_count",), "text_transfer": ("mmada_text_transfer",),
    "lfq_decode_nhwc": ("mmada_lfq_decode_nhwc",), "lfq_indices_to_bits": ("mmada_lfq_indices_to_bits",),
    "lfq_bits_to_indices": ("mmada_lfq_bits_to_indices",),
    "groupnorm_swish": ("mmada_groupnorm_stats", "mmada_groupnorm_apply_bf16"), "upsample2x_nhwc": ("mmada_upsample2x_nhwc_bf16",),
    "cast_bf16": ("mmada_cast_f32_bf16",), "softmax_rows_bf16": ("mmada_softmax_rows_bf16",),
    "nhwc_to_nchw": ("mmada_nhwc_to_nchw_f32",), "image_to_uint8": ("mmada_image_to_uint8",),
    "image_to_nhwc64": ("mmada_image_to_nhwc64_bf16",), "space_to_depth2": ("mmada_space_to_depth2_bf16",),
    "conv1d_gather": ("mmada_conv1d_gather_bf16",), "relu_": ("mmada_relu_f32",), "build_prompts": ("mmada_build_prompts",),
}
