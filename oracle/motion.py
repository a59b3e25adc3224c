"""Oracle: motion VQ-VAE decode (codebook look-up + 1-D convolutional decoder), restated over a state dict.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
  * /root/reference/motion_vqvae/models/vqvae.py:74-81      VQVAE_251.forward_decoder (dequantize, view(1,-1,C), NCT)
  * /root/reference/motion_vqvae/models/quantize_cnn.py:89-91  dequantize = F.embedding(code_idx, codebook)
  * /root/reference/motion_vqvae/models/encdec.py:35-67     Decoder (Conv1d/ReLU, down_t x [Resnet1D, Upsample, Conv1d],
                                                            Conv1d/ReLU, Conv1d)
  * /root/reference/motion_vqvae/models/resnet.py:12-81     ResConv1DBlock (act, conv k3 dilated, act, conv 1x1, +x),
                                                            Resnet1D with reverse_dilation (dilations rate^(depth-1)..1)
Pinned by oracle/make_goldens.py::motion_case against the reference's own ``Decoder`` class (its quantiser module
calls .cuda() in __init__ and cannot be constructed here; the look-up and reshape around it are four lines).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

from .weights import _normal

MOTION = dict(nb_code=512, code_dim=512, width=512, down_t=2, depth=3, rate=3, n_feats=263)   # configs/t2m_test.yaml


def make_motion_decoder_weights(seed: int = 0, cfg: dict = MOTION) -> Dict[str, torch.Tensor]:
    """fp32 state dict with the reference's key names (HumanVQVAE: 'vqvae.quantizer.codebook', 'vqvae.decoder.model.*')."""
    sd: Dict[str, torch.Tensor] = {}
    W = cfg["width"]

    def conv(name, co, ci, k):
        sd[f"vqvae.decoder.model.{name}.weight"] = _normal(seed, "mot." + name + ".w", (co, ci, k), 1.0 / math.sqrt(ci * k))
        sd[f"vqvae.decoder.model.{name}.bias"] = _normal(seed, "mot." + name + ".b", (co,), 0.02)

    sd["vqvae.quantizer.codebook"] = _normal(seed, "mot.codebook", (cfg["nb_code"], cfg["code_dim"]), 1.0)
    conv("0", W, cfg["code_dim"], 3)
    for i in range(cfg["down_t"]):
        for j in range(cfg["depth"]):
            conv(f"{2 + i}.0.model.{j}.conv1", W, W, 3)
            conv(f"{2 + i}.0.model.{j}.conv2", W, W, 1)
        conv(f"{2 + i}.2", W, W, 3)
    conv(f"{2 + cfg['down_t']}", W, W, 3)
    conv(f"{4 + cfg['down_t']}", cfg["n_feats"], W, 3)
    return sd


def decoder_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: dict = MOTION) -> torch.Tensor:
    """x (N, C, T) fp32 -> (N, n_feats, T * 2**down_t)."""
    D = "vqvae.decoder.model."
    c = lambda h, name, pad=1, dil=1: F.conv1d(h, sd[D + name + ".weight"], sd[D + name + ".bias"], stride=1, padding=pad, dilation=dil)
    h = F.relu(c(x, "0"))
    for i in range(cfg["down_t"]):
        for j in range(cfg["depth"]):
            d = cfg["rate"] ** (cfg["depth"] - 1 - j)
            r = c(F.relu(h), f"{2 + i}.0.model.{j}.conv1", pad=d, dil=d)
            h = c(F.relu(r), f"{2 + i}.0.model.{j}.conv2", pad=0) + h
        h = c(F.interpolate(h, scale_factor=2, mode="nearest"), f"{2 + i}.2")
    h = F.relu(c(h, f"{2 + cfg['down_t']}"))
    return c(h, f"{4 + cfg['down_t']}")


def forward_decoder(code_idx: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: dict = MOTION) -> torch.Tensor:
    """All ids as ONE sequence (the reference's view(1, -1, code_dim)) -> (1, numel * 2**down_t, n_feats)."""
    x_d = F.embedding(code_idx, sd["vqvae.quantizer.codebook"])
    x_d = x_d.view(1, -1, cfg["code_dim"]).permute(0, 2, 1).contiguous()
    return decoder_forward(x_d, sd, cfg).permute(0, 2, 1)
