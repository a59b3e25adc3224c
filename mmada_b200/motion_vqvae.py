"""Motion VQ-VAE, token -> pose side — host-side mirror of the reference's ``HumanVQVAE.forward_decoder`` /
``VQVAE_251.forward_decoder`` (/root/reference/motion_vqvae/models/vqvae.py:74-81,115-117): codebook look-up
(``QuantizeEMAReset.dequantize`` quantize_cnn.py:89-91) and the 1-D convolutional ``Decoder`` (encdec.py:35-67,
``Resnet1D`` / ``ResConv1DBlock`` resnet.py:12-81; activation relu, no norm).  It is the decode step after
``MMadaModelLM.t2m_generate`` (BASELINE config 5).

Layout: frames x channels ([B, T, C], channels contiguous) instead of the reference's NCT; the trunk is fp32.  Every
Conv1d is one tcgen05 GEMM (``ops.gemm``, fp32 accumulation, fp32 bias / residual epilogue) over a bf16 operand
gathered by ``ops.conv1d_gather`` (taps side by side; fuses the ReLU in front of the convolution and the nearest
2x upsample).  The reference decodes ONE sequence per call (``view(1, -1, code_dim)``); ``forward_decoder`` keeps
that contract, ``forward_decoder_batched`` decodes B independent sequences in one pass.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops


class HumanVQVAE:
    def __init__(self, nb_code: int = 512, code_dim: int = 512, output_emb_width: int = 512, down_t: int = 2, stride_t: int = 2,
                 width: int = 512, depth: int = 3, dilation_growth_rate: int = 3, dataset_name: str = "t2m", device="cuda"):
        if code_dim != output_emb_width:
            raise ValueError("code_dim must equal output_emb_width (the decoder consumes code vectors)")
        self.device = torch.device(device)
        self.nb_code, self.code_dim, self.width, self.down_t, self.depth = nb_code, code_dim, width, down_t, depth
        self.rate = dilation_growth_rate
        self.n_feats = 251 if dataset_name == "kit" else 263
        self.w: Dict[str, torch.Tensor] = {}
        self.kernel_launches = 0

    # ---- weights ---------------------------------------------------------------------------
    def _conv(self, sd, src: str, dst: str):
        wt = sd[src + ".weight"].to(self.device, torch.float32)                  # [Cout, Cin, k]
        co, ci, k = wt.shape
        self.w[dst + ".w"] = wt.permute(0, 2, 1).reshape(co, k * ci).to(torch.bfloat16).contiguous()   # [Cout, k*Cin], tap-major
        self.w[dst + ".b"] = sd[src + ".bias"].to(self.device, torch.float32).contiguous()

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = False) -> "HumanVQVAE":
        """``sd`` uses the reference's key names (``vqvae.quantizer.codebook``, ``vqvae.decoder.model.*``); a leading
        ``vqvae.`` is optional (VQVAE_251 vs HumanVQVAE checkpoints)."""
        sd = {(k[len("vqvae."):] if k.startswith("vqvae.") else k): v for k, v in sd.items()}
        self.w = {"codebook": sd["quantizer.codebook"].to(self.device, torch.bfloat16).contiguous()}
        D = "decoder.model."
        self._conv(sd, D + "0", "in")
        for i in range(self.down_t):
            for j in range(self.depth):
                self._conv(sd, f"{D}{2 + i}.0.model.{j}.conv1", f"b{i}.r{j}.c1")
                self._conv(sd, f"{D}{2 + i}.0.model.{j}.conv2", f"b{i}.r{j}.c2")
            self._conv(sd, f"{D}{2 + i}.2", f"b{i}.up")
        self._conv(sd, f"{D}{2 + self.down_t}", "mid")
        self._conv(sd, f"{D}{4 + self.down_t}", "out")
        return self

    def init_random(self, seed: int = 0) -> "HumanVQVAE":
        """Random decoder weights with the reference's shapes, drawn on the device (benchmarks: no checkpoint offline)."""
        g = torch.Generator(device=self.device).manual_seed(seed)
        W, cd = self.width, self.code_dim

        def rnd(shape, std):
            return torch.randn(shape, generator=g, device=self.device, dtype=torch.float32) * std

        sd = {"quantizer.codebook": rnd((self.nb_code, cd), 1.0)}
        D = "decoder.model."

        def conv(name, co, ci, k):
            sd[f"{D}{name}.weight"] = rnd((co, ci, k), (ci * k) ** -0.5)
            sd[f"{D}{name}.bias"] = rnd((co,), 0.02)

        conv("0", W, cd, 3)
        for i in range(self.down_t):
            for j in range(self.depth):
                conv(f"{2 + i}.0.model.{j}.conv1", W, W, 3)
                conv(f"{2 + i}.0.model.{j}.conv2", W, W, 1)
            conv(f"{2 + i}.2", W, W, 3)
        conv(f"{2 + self.down_t}", W, W, 3)
        conv(f"{4 + self.down_t}", self.n_feats, W, 3)
        return self.load_state_dict(sd)

    # ---- decoder -----------------------------------------------------------------------------
    def _c(self, x, key, taps, dilation=1, upsample=1, relu=False, resid=None):
        """Conv1d(k = taps, dilation, 'same' padding) of act(x), x fp32 [B,T,C] -> fp32 [B, T*upsample, Cout]."""
        B, T, C = x.shape
        a = ops.conv1d_gather(x, taps, dilation, upsample, relu).view(B * T * upsample, taps * C)
        wt, bias = self.w[key + ".w"], self.w[key + ".b"]
        self.kernel_launches += 2
        if resid is None:
            out = ops.gemm(a, wt, ops.EPI_BIAS_F32, bias=bias)
        else:
            r = resid.view(B * T * upsample, -1)
            out = ops.gemm(a, wt, ops.EPI_BIAS_RESID_F32, out=torch.empty_like(r), aux=r, bias=bias)
        return out.view(B, T * upsample, -1)

    @torch.no_grad()
    def forward_decoder_batched(self, code_idx: torch.Tensor) -> torch.Tensor:
        """(B, T) int64 code ids, every row one sequence -> (B, T * 2**down_t, n_feats) fp32."""
        ids = code_idx.to(self.device, torch.int64)
        B, T = ids.shape
        x = ops.embed(ids, self.w["codebook"]).view(B, T, self.code_dim)          # dequantize
        self.kernel_launches += 1
        x = ops.relu_(self._c(x, "in", 3))                                         # Conv1d, ReLU
        self.kernel_launches += 1
        for i in range(self.down_t):
            for j in range(self.depth):                                            # Resnet1D, dilations rate^(depth-1) .. 1
                d = self.rate ** (self.depth - 1 - j)
                h = self._c(x, f"b{i}.r{j}.c1", 3, dilation=d, relu=True)         # act -> conv1 (k3, dilated)
                x = self._c(h, f"b{i}.r{j}.c2", 1, relu=True, resid=x)            # act -> conv2 (1x1) -> + x
            x = self._c(x, f"b{i}.up", 3, upsample=2)                              # nearest 2x -> Conv1d
        x = ops.relu_(self._c(x, "mid", 3))
        self.kernel_launches += 1
        return self._c(x, "out", 3)

    @torch.no_grad()
    def forward_decoder(self, x: torch.Tensor) -> torch.Tensor:
        """The reference's contract: ALL ids of ``x`` form one sequence -> (1, numel * 2**down_t, n_feats)."""
        return self.forward_decoder_batched(x.reshape(1, -1))
