"""Oracle: LLaDA mask-predictor forward, restated functionally over a state dict.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
/root/reference/models/modeling_llada.py:
  * LLaDAModel.forward            :1161-1366  (embedding -> blocks -> ln_f -> ff_out)
  * LLaDALlamaBlock.forward       :886-934
  * LLaDABlock.attention          :662-724    (bias never applied: Appendix A, Q1)
  * RMSLayerNorm.forward          :315-329    (fp32 statistics, weight applied after the down-cast)
  * RotaryEmbedding               :376-428    (NeoX half-split, fp32)
The op sequence is kept identical to the reference's so that on CPU/fp32 it is bit-identical
(asserted by oracle/make_goldens.py against the real reference).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

_P = "model.transformer."


def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    og = x.dtype
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(-1, keepdim=True)
    xf = xf * torch.rsqrt(var + eps)
    return weight * xf.to(og)


def rope_tables(seq_len: int, head_dim: int, theta: float, device="cpu"):
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, device=device, dtype=torch.float) / head_dim))
    seq = torch.arange(seq_len, device=device, dtype=torch.float)
    freqs = torch.einsum("i , j -> i j", seq, inv_freq)
    pos = torch.cat((freqs, freqs), dim=-1)
    return pos.sin()[None, None, :, :], pos.cos()[None, None, :, :]


def _rotate_half(x: torch.Tensor) -> torch.Tensor:
    B, nh, T, hs = x.size()
    x = x.view(B, nh, T, 2, hs // 2)
    x1, x2 = x.unbind(dim=-2)
    return torch.cat((-x2, x1), dim=-1)


def apply_rope(q: torch.Tensor, k: torch.Tensor, theta: float):
    """q, k: (B, H, T, hd) in model dtype -> rotated, same dtype (fp32 math)."""
    q_, k_ = q.float(), k.float()
    T = k_.shape[-2]
    sin, cos = rope_tables(T, q.shape[-1], theta, q.device)
    q_ = ((q_ * cos) + (_rotate_half(q_) * sin)).to(q_.dtype)
    k_ = ((k_ * cos) + (_rotate_half(k_) * sin)).to(k_.dtype)
    return q_.type_as(q), k_.type_as(k)


def block_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], i: int, cfg: dict) -> torch.Tensor:
    b = f"{_P}blocks.{i}."
    H = cfg["n_heads"]
    B, T, C = x.shape
    xn = rms_norm(x, sd[b + "attn_norm.weight"], cfg["rms_norm_eps"])
    q = F.linear(xn, sd[b + "q_proj.weight"])
    k = F.linear(xn, sd[b + "k_proj.weight"])
    v = F.linear(xn, sd[b + "v_proj.weight"])
    q = q.view(B, T, H, C // H).transpose(1, 2)
    k = k.view(B, T, H, C // H).transpose(1, 2)
    v = v.view(B, T, H, C // H).transpose(1, 2)
    q, k = apply_rope(q, k, cfg["rope_theta"])
    att = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
    att = att.transpose(1, 2).contiguous().view(B, T, C)
    x = x + F.linear(att, sd[b + "attn_out.weight"])
    og = x
    h = rms_norm(x, sd[b + "ff_norm.weight"], cfg["rms_norm_eps"])
    g, u = F.linear(h, sd[b + "ff_proj.weight"]), F.linear(h, sd[b + "up_proj.weight"])
    h = F.silu(g) * u
    return og + F.linear(h, sd[b + "ff_out.weight"])


def hidden_states(input_ids: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: dict) -> torch.Tensor:
    """Final-norm hidden states (B, L, d) — everything except the vocabulary projection."""
    x = F.embedding(input_ids, sd[_P + "wte.weight"])
    for i in range(cfg["n_layers"]):
        x = block_forward(x, sd, i, cfg)
    return rms_norm(x, sd[_P + "ln_f.weight"], cfg["rms_norm_eps"])


def forward_logits(input_ids: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: dict,
                   rows: Optional[slice] = None, cols: Optional[slice] = None) -> torch.Tensor:
    """``model(input_ids).logits`` (B, L, V).  ``rows``/``cols`` restrict the vocabulary projection
    to a slice of positions / vocabulary entries; by construction this equals slicing the full
    result (a Linear is row- and column-separable), which is what the CUDA path computes."""
    h = hidden_states(input_ids, sd, cfg)
    w = sd[_P + "ff_out.weight"]
    if rows is not None:
        h = h[:, rows]
    if cols is not None:
        w = w[cols]
    return F.linear(h, w)
