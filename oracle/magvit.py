"""Oracle: MAGVIT-v2 look-up-free quantiser and VQGAN decoder, restated over a state dict.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
  * /root/reference/models/modeling_magvitv2.py:186-221   LFQuantizer tables / get_indices / get_codebook_entry
  * /root/reference/models/modeling_magvitv2.py:365-399   VQGANDecoder.forward
  * /root/reference/models/modeling_magvitv2.py:429-433   MAGVITv2.decode_code
  * /root/reference/models/modeling_magvitv2.py:143-169   VQGANEncoder.forward; :423-427 MAGVITv2.get_code;
    :236-241 the quantiser's sign test; common_modules.py:73-90 Downsample (pad (0,1,0,1) + 3x3 stride 2)
  * /root/reference/models/common_modules.py:16-40,168-211,298-357  swish, GroupNorm(32, eps 1e-6),
    Upsample (nearest 2x + conv3x3), AttnBlock, ResnetBlock
The bit <-> index maps are restated in numpy (integer work); the decoder in fp32 torch.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from .weights import vq_decoder_plan, vq_encoder_plan

CODE_BITS = 13


def lfq_indices_to_bits(indices: np.ndarray) -> np.ndarray:
    """(B, N) ints -> (B, 13, h, w) float32 of -1/+1, channel k = bit (12-k) (MSB first)."""
    indices = np.asarray(indices, dtype=np.int64)
    b, n = indices.shape
    h = w = int(math.sqrt(n))
    shifts = np.arange(CODE_BITS - 1, -1, -1, dtype=np.int64)
    bits = (indices.reshape(-1, 1) >> shifts) & 1                  # (B*N, 13)
    z = bits.astype(np.float32) * 2 - 1
    return np.ascontiguousarray(z.reshape(b, h, w, CODE_BITS).transpose(0, 3, 1, 2))


def lfq_bits_to_indices(z: np.ndarray) -> np.ndarray:
    """(B, 13, h, w) float -> (B, 1, h, w) int64: sum_k 2^(12-k) * [z_k > 0]."""
    pw = (2 ** np.arange(CODE_BITS - 1, -1, -1)).astype(np.float32).reshape(1, -1, 1, 1)
    return (pw * (np.asarray(z) > 0).astype(np.float32)).sum(1, keepdims=True).astype(np.int64)


def _swish(x):
    return x * torch.sigmoid(x)


def _gn(x, sd, key):
    return F.group_norm(x, 32, sd[key + ".weight"], sd[key + ".bias"], eps=1e-6)


def _conv(x, sd, key, pad):
    return F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], stride=1, padding=pad)


def _res(x, sd, key, ci, co):
    h = _conv(_swish(_gn(x, sd, key + ".norm1")), sd, key + ".conv1", 1)
    h = _conv(_swish(_gn(h, sd, key + ".norm2")), sd, key + ".conv2", 1)
    if ci != co:
        x = _conv(x, sd, key + ".nin_shortcut", 0)
    return x + h


def _attn(x, sd, key):
    h_ = _gn(x, sd, key + ".norm")
    q, k, v = (_conv(h_, sd, f"{key}.{n}", 0) for n in ("q", "k", "v"))
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + _conv(h_, sd, key + ".proj_out", 0)


def decoder_forward(z: torch.Tensor, sd: Dict[str, torch.Tensor], taps: Optional[dict] = None) -> torch.Tensor:
    """z (B, 13, h, w) fp32 -> (B, 3, 16h, 16w) fp32.  ``sd`` keys are prefixed 'decoder.'.
    ``taps`` (optional dict) receives the activation after every plan entry."""
    h = z
    for kind, key, ci, co in vq_decoder_plan():
        k = "decoder." + key
        if kind == "conv1":
            h = _conv(h, sd, k, 0)
        elif kind == "conv3":
            h = _conv(h, sd, k, 1)
        elif kind == "res":
            h = _res(h, sd, k, ci, co)
        elif kind == "attn":
            h = _attn(h, sd, k)
        elif kind == "up":
            h = _conv(F.interpolate(h, scale_factor=2.0, mode="nearest"), sd, k + ".conv", 1)
        elif kind == "norm_out":
            h = _swish(_gn(h, sd, k))
        if taps is not None:
            taps[key] = h
    return h


def decode_code(indices: torch.Tensor, sd: Dict[str, torch.Tensor], taps: Optional[dict] = None) -> torch.Tensor:
    z = torch.from_numpy(lfq_indices_to_bits(indices.cpu().numpy()))
    return decoder_forward(z, sd, taps)


def encoder_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], taps: Optional[dict] = None) -> torch.Tensor:
    """pixels (B, 3, H, W) fp32 -> pre-quantisation latents (B, 13, H/16, W/16) fp32 (quant_conv included).
    ``sd`` keys are prefixed 'encoder.'."""
    h = x
    for kind, key, ci, co in vq_encoder_plan():
        k = "encoder." + key
        if kind == "conv1":
            h = _conv(h, sd, k, 0)
        elif kind == "conv3":
            h = _conv(h, sd, k, 1)
        elif kind == "res":
            h = _res(h, sd, k, ci, co)
        elif kind == "attn":
            h = _attn(h, sd, k)
        elif kind == "down":
            h = F.conv2d(F.pad(h, (0, 1, 0, 1), mode="constant", value=0), sd[k + ".conv.weight"], sd[k + ".conv.bias"],
                         stride=2, padding=0)
        elif kind == "norm_out":
            h = _swish(_gn(h, sd, k))
        if taps is not None:
            taps[key] = h
    return h


def get_code(pixels: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """(B, 3, H, W) -> (B, H/16 * W/16) int64: bit k of the code = [latent channel k > 0] (z_q = +-1 by sign)."""
    z = encoder_forward(pixels, sd)
    return torch.from_numpy(lfq_bits_to_indices(z.numpy())).reshape(pixels.shape[0], -1)
