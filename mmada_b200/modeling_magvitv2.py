"""MAGVIT-v2 tokenizer, token -> pixel side — host-side mirror of the reference's ``MAGVITv2``
(/root/reference/models/modeling_magvitv2.py:402-433: ``decode_code`` -> ``LFQuantizer.get_codebook_entry``
:208-221 -> ``VQGANDecoder.forward`` :365-399, blocks in models/common_modules.py) on the B200 kernels.

Layout: activations are NHWC.  The trunk ``h`` (conv outputs, residual sums) is fp32; every
convolution input is the bf16 output of the fused GroupNorm(+swish) kernel (or a bf16 cast / nearest
upsample of the trunk), so convolutions run as bf16 implicit GEMMs on tcgen05 with fp32 accumulation
and fp32 bias / residual epilogues.  The reference runs this model in fp32 (Q18); the stated
tolerance of the parity tests reflects the bf16 operands.
``get_code`` (pixel -> token side, SURVEY.md 8(f) item 1: ``VQGANEncoder.forward`` :143-169 + the LFQ sign test
:201-206,236-241) reuses the same kernels; its stride-2 ``Downsample`` convolutions run as stride-1 convolutions
over a space-to-depth image with rearranged weights.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import ops

CH, CH_MULT, NUM_RES, Z_CH = 128, (1, 1, 2, 2, 4), (4, 4, 3, 4, 3), 13
ENC_CH_MULT, ENC_NUM_RES = (1, 2, 2, 4, 4), (4, 3, 4, 3, 4)          # VQGANEncoder defaults (modeling_magvitv2.py:49-72)


def encoder_plan():
    """(kind, key, c_in, c_out) in execution order (modeling_magvitv2.py:74-141 / :143-169); no attention in the
    down path (attn_resolutions [5] never matches)."""
    plan = [("conv_in", "conv_in", 3, CH)]
    c = CH
    for lvl in range(5):
        co = CH * ENC_CH_MULT[lvl]
        for j in range(ENC_NUM_RES[lvl]):
            plan.append(("res", f"down.{lvl}.block.{j}", c, co))
            c = co
        if lvl != 4:
            plan.append(("down", f"down.{lvl}.downsample", c, c))
    plan += [("res", "mid.block_1", c, c), ("attn", "mid.attn_1", c, c), ("res", "mid.block_2", c, c),
             ("out", "conv_out", c, Z_CH)]
    return plan


def decoder_plan():
    """(kind, key, c_in, c_out) in execution order (modeling_magvitv2.py:309-362 / :365-399)."""
    c = CH * CH_MULT[-1]
    plan = [("conv_in", "conv_in", Z_CH, c), ("res", "mid.block_1", c, c), ("attn", "mid.attn_1", c, c),
            ("res", "mid.block_2", c, c)]
    for lvl in reversed(range(5)):
        co = CH * CH_MULT[lvl]
        for j in range(NUM_RES[lvl]):
            plan.append(("res", f"up.{lvl}.block.{j}", c, co))
            c = co
        if lvl != 0:
            plan.append(("up", f"up.{lvl}.upsample", c, c))
    plan += [("out", "conv_out", c, 3)]
    return plan


class _LFQ:
    """``vq_model.quantize``: the look-up-free quantiser's index <-> bit maps."""

    codebook_size = 8192
    e_dim = 13

    def get_codebook_entry(self, indices: torch.Tensor, shape=None) -> torch.Tensor:
        b, n = indices.shape
        h, w = (int(math.sqrt(n)),) * 2 if shape is None else shape
        return ops.lfq_indices_to_bits(indices).view(b, 13, h, w)

    def get_indices(self, z_q: torch.Tensor) -> torch.Tensor:
        b, _, h, w = z_q.shape
        return ops.lfq_bits_to_indices(z_q.float().contiguous().view(b, 13, h * w)).view(b, 1, h, w)


class MAGVITv2:
    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        self.quantize = _LFQ()
        self.w: Dict[str, torch.Tensor] = {}
        self.kernel_launches = 0

    # ---- weights ---------------------------------------------------------------------------
    def _conv(self, sd, key, pad_cin_to: Optional[int] = None, src="decoder.", dst=""):
        wt = sd[f"{src}{key}.weight"].to(self.device, torch.float32)             # [Cout, Cin, kh, kw]
        co, ci, kh, kw = wt.shape
        wt = wt.permute(0, 2, 3, 1)                                              # [Cout, kh, kw, Cin]
        if pad_cin_to is not None and pad_cin_to > ci:
            wt = torch.nn.functional.pad(wt, (0, pad_cin_to - ci))
        self.w[dst + key + ".w"] = wt.reshape(co, -1).to(torch.bfloat16).contiguous()
        self.w[dst + key + ".b"] = sd[f"{src}{key}.bias"].to(self.device, torch.float32).contiguous()

    def _norm(self, sd, key, src="decoder.", dst=""):
        self.w[dst + key + ".g"] = sd[f"{src}{key}.weight"].to(self.device, torch.float32).contiguous()
        self.w[dst + key + ".beta"] = sd[f"{src}{key}.bias"].to(self.device, torch.float32).contiguous()

    def _load_encoder(self, sd):
        """``encoder.*`` keys -> self.w["enc." + ...].  Downsample: w[co,ci,ky,kx] (stride 2 on the image padded right /
        bottom) becomes a stride-1 3x3 kernel over the space-to-depth image: pixel (2y+ky, 2x+kx) is sub-pixel
        (ky%2, kx%2) of block (y + ky//2, x + kx//2), i.e. tap (ky//2, kx//2) in {0,1}^2 of the 3x3 window centred on
        the block; taps reaching above / left stay zero.  quant_conv (1x1, linear) is folded into conv_out."""
        E = "enc."
        for kind, key, ci, co in encoder_plan():
            if kind == "conv_in":
                self._conv(sd, key, pad_cin_to=64, src="encoder.", dst=E)
            elif kind == "res":
                self._norm(sd, key + ".norm1", "encoder.", E); self._conv(sd, key + ".conv1", src="encoder.", dst=E)
                self._norm(sd, key + ".norm2", "encoder.", E); self._conv(sd, key + ".conv2", src="encoder.", dst=E)
                if ci != co:
                    self._conv(sd, key + ".nin_shortcut", src="encoder.", dst=E)
            elif kind == "attn":
                self._norm(sd, key + ".norm", "encoder.", E)
                for n in ("q", "k", "v", "proj_out"):
                    self._conv(sd, f"{key}.{n}", src="encoder.", dst=E)
                self.w[E + key + ".qk.w"] = torch.cat([self.w[E + key + ".q.w"], self.w[E + key + ".k.w"]], 0).contiguous()
                self.w[E + key + ".qk.b"] = torch.cat([self.w[E + key + ".q.b"], self.w[E + key + ".k.b"]], 0).contiguous()
            elif kind == "down":
                wt = sd[f"encoder.{key}.conv.weight"].to(self.device, torch.float32)          # [C, C, 3, 3]
                C = wt.shape[0]
                w2 = torch.zeros((C, 3, 3, 4, C), device=self.device, dtype=torch.float32)    # [co, ty, tx, sub, ci]
                for ky in range(3):
                    for kx in range(3):
                        w2[:, ky // 2 + 1, kx // 2 + 1, 2 * (ky % 2) + (kx % 2), :] = wt[:, :, ky, kx]
                self.w[E + key + ".w"] = w2.reshape(C, -1).to(torch.bfloat16).contiguous()
                self.w[E + key + ".b"] = sd[f"encoder.{key}.conv.bias"].to(self.device, torch.float32).contiguous()
            elif kind == "out":
                self._norm(sd, "norm_out", "encoder.", E)
                wc = sd["encoder.conv_out.weight"].to(self.device, torch.float32)             # [13, C, 3, 3]
                bc = sd["encoder.conv_out.bias"].to(self.device, torch.float32)
                wq = sd["encoder.quant_conv.weight"].to(self.device, torch.float32).reshape(Z_CH, Z_CH)
                bq = sd["encoder.quant_conv.bias"].to(self.device, torch.float32)
                wf = torch.einsum("oz,zchw->ochw", wq, wc).permute(0, 2, 3, 1)                # quant_conv o conv_out
                self.w[E + "conv_out.w"] = wf.reshape(Z_CH, -1).to(torch.bfloat16).contiguous()
                self.w[E + "conv_out.b"] = (wq @ bc + bq).contiguous()
        self.has_encoder = True

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = False) -> "MAGVITv2":
        """``sd`` uses the reference's key names (``decoder.*`` and, for ``get_code``, ``encoder.*``; SURVEY.md
        Appendix D).  Either half may be absent."""
        self.w = {}
        self.has_encoder = False
        if any(k.startswith("encoder.") for k in sd):
            self._load_encoder(sd)
        if not any(k.startswith("decoder.") for k in sd):
            return self
        self.w["pq.w"] = sd["decoder.post_quant_conv.weight"].to(self.device, torch.float32).reshape(13, 13).contiguous()
        self.w["pq.b"] = sd["decoder.post_quant_conv.bias"].to(self.device, torch.float32).contiguous()
        for kind, key, ci, co in decoder_plan():
            if kind == "conv_in":
                self._conv(sd, key, pad_cin_to=64)
            elif kind == "res":
                self._norm(sd, key + ".norm1"); self._conv(sd, key + ".conv1")
                self._norm(sd, key + ".norm2"); self._conv(sd, key + ".conv2")
                if ci != co:
                    self._conv(sd, key + ".nin_shortcut")
            elif kind == "attn":
                self._norm(sd, key + ".norm")
                for n in ("q", "k", "v", "proj_out"):
                    self._conv(sd, f"{key}.{n}")
                self.w[key + ".qk.w"] = torch.cat([self.w[key + ".q.w"], self.w[key + ".k.w"]], 0).contiguous()
                self.w[key + ".qk.b"] = torch.cat([self.w[key + ".q.b"], self.w[key + ".k.b"]], 0).contiguous()
            elif kind == "up":
                self._conv(sd, key + ".conv")
            elif kind == "out":
                self._norm(sd, "norm_out"); self._conv(sd, key)
        return self

    @classmethod
    def from_pretrained(cls, path: str, device="cuda", **_ignored) -> "MAGVITv2":
        """``MAGVITv2.from_pretrained(dir)`` (reference inference_t2i.py:67): ``config.json`` (the reference's class takes
        no configuration arguments — the architecture is fixed, models/modeling_magvitv2.py:403-411) and
        ``pytorch_model.safetensors`` | ``pytorch_model.bin`` (models/modeling_utils.py:47-49) with ``encoder.*`` /
        ``decoder.*`` keys."""
        from .checkpoint import ShardedCheckpoint
        return cls(device=device).load_state_dict(ShardedCheckpoint(path))

    def init_random(self, seed: int = 0) -> "MAGVITv2":
        """Random decoder weights on the device (benchmarks: the checkpoint is not available offline)."""
        g = torch.Generator(device=self.device).manual_seed(seed)
        sd = {}

        def conv(name, co, ci, k):
            sd[f"decoder.{name}.weight"] = torch.randn((co, ci, k, k), device=self.device, generator=g) / math.sqrt(ci * k * k)
            sd[f"decoder.{name}.bias"] = torch.randn((co,), device=self.device, generator=g) * 0.02

        def norm(name, c):
            sd[f"decoder.{name}.weight"] = torch.ones(c, device=self.device)
            sd[f"decoder.{name}.bias"] = torch.zeros(c, device=self.device)

        conv("post_quant_conv", 13, 13, 1)
        for kind, key, ci, co in decoder_plan():
            if kind == "conv_in":
                conv(key, co, ci, 3)
            elif kind == "res":
                norm(key + ".norm1", ci); conv(key + ".conv1", co, ci, 3); norm(key + ".norm2", co); conv(key + ".conv2", co, co, 3)
                if ci != co:
                    conv(key + ".nin_shortcut", co, ci, 1)
            elif kind == "attn":
                norm(key + ".norm", ci)
                for n in ("q", "k", "v", "proj_out"):
                    conv(f"{key}.{n}", co, ci, 1)
            elif kind == "up":
                conv(key + ".conv", co, ci, 3)
            elif kind == "out":
                norm("norm_out", ci); conv(key, co, ci, 3)
        return self.load_state_dict(sd)

    # ---- blocks ----------------------------------------------------------------------------
    def _gn(self, x, key, swish=True):
        self.kernel_launches += 2
        return ops.groupnorm_swish(x, self.w[key + ".g"], self.w[key + ".beta"], self._sums, swish=swish)

    def _c(self, x, key, taps, resid=None):
        self.kernel_launches += 1
        epi = ops.EPI_BIAS_F32 if resid is None else ops.EPI_BIAS_RESID_F32
        return ops.conv_nhwc(x, self.w[key + ".w"], self.w[key + ".b"], taps, epi, resid=resid)

    def _res(self, x, key, ci, co):
        h = self._c(self._gn(x, key + ".norm1"), key + ".conv1", 9)
        h = self._gn(h, key + ".norm2")
        if ci != co:
            self.kernel_launches += 1
            x = self._c(ops.cast_bf16(x), key + ".nin_shortcut", 1)
        return self._c(h, key + ".conv2", 9, resid=x)

    @torch.no_grad()
    def _encode_nhwc(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """fp32 NCHW pixels -> pre-quantisation latents fp32 NHWC [B, H/16, W/16, 13] (quant_conv included)."""
        if not getattr(self, "has_encoder", False):
            raise ops._lib.MMadaKernelError("MAGVITv2.get_code: no encoder weights loaded (state dict had no 'encoder.*' keys)")
        px = pixel_values.to(self.device, torch.float32).contiguous()
        B, _, H, W = px.shape
        if H % 16 or W % 16:
            raise ValueError("get_code: image sides must be multiples of 16")
        self._sums = torch.empty((B, 32, 2), device=self.device, dtype=torch.float64)
        x = ops.image_to_nhwc64(px)
        self.kernel_launches += 1
        E = "enc."
        for kind, key, ci, co in encoder_plan():
            if kind == "conv_in":
                x = self._c(x, E + key, 9)
            elif kind == "res":
                x = self._res(x, E + key, ci, co)
            elif kind == "attn":
                x = self._attn(x, E + key)
            elif kind == "down":
                self.kernel_launches += 1
                x = self._c(ops.space_to_depth2(x), E + key, 9)
            elif kind == "out":
                x = self._c(self._gn(x, E + "norm_out"), E + key, 9)
        return x

    def _attn(self, x, key):
        B, H, W, C = x.shape
        P = H * W
        hn = self._gn(x, key + ".norm", swish=False).view(B * P, C)
        qk = ops.gemm(hn, self.w[key + ".qk.w"], ops.EPI_BIAS_BF16, bias=self.w[key + ".qk.b"])          # [B*P, 2C]
        o = torch.empty((B * P, C), device=x.device, dtype=torch.bfloat16)
        for b in range(B):
            r = slice(b * P, (b + 1) * P)
            s = ops.gemm(qk[r, :C], qk[r, C:], ops.EPI_F32)                                              # [P, P] q.k^T
            pm = ops.softmax_rows_bf16(s, float(int(C) ** (-0.5)))
            vt = ops.gemm(self.w[key + ".v.w"], hn[r], ops.EPI_BF16)                                     # [C, P] = (Wv h)^T
            # softmax rows sum to one, so the value bias passes through the attention unchanged
            ops.gemm(pm, vt, ops.EPI_BIAS_BF16, out=o[r], bias=self.w[key + ".v.b"])
        self.kernel_launches += 1 + 4 * B
        return self._c(o.view(B, H, W, C), key + ".proj_out", 1, resid=x)

    # ---- public API ------------------------------------------------------------------------
    @torch.no_grad()
    def _decode_nhwc(self, codebook_indices: torch.Tensor, shape=None) -> torch.Tensor:
        idx = codebook_indices.to(self.device, torch.int64).contiguous()
        B, n = idx.shape
        h, w = (int(math.sqrt(n)),) * 2 if shape is None else shape
        self._sums = torch.empty((B, 32, 2), device=self.device, dtype=torch.float64)
        x = ops.lfq_decode_nhwc(idx, self.w["pq.w"], self.w["pq.b"], h, w)            # bits -> post_quant_conv, bf16, 64 ch
        self.kernel_launches += 1
        for kind, key, ci, co in decoder_plan():
            if kind == "conv_in":
                x = self._c(x, key, 9)
            elif kind == "res":
                x = self._res(x, key, ci, co)
            elif kind == "attn":
                x = self._attn(x, key)
            elif kind == "up":
                self.kernel_launches += 1
                x = self._c(ops.upsample2x_nhwc(x), key + ".conv", 9)
            elif kind == "out":
                x = self._c(self._gn(x, "norm_out"), key, 9)
        return x                                                                       # fp32 [B, 16h, 16w, 3]

    @torch.no_grad()
    def decode_code(self, codebook_indices: torch.Tensor, shape=None) -> torch.Tensor:
        """(B, N) code ids -> reconstructed pixels fp32 (B, 3, 16*sqrt(N), 16*sqrt(N)), like the reference."""
        self.kernel_launches += 1
        return ops.nhwc_to_nchw(self._decode_nhwc(codebook_indices, shape))

    @torch.no_grad()
    def decode_code_uint8(self, codebook_indices: torch.Tensor, shape=None) -> torch.Tensor:
        """decode_code followed by inference_t2i.py:123-125's clamp((x+1)/2)*255 -> uint8, NHWC."""
        self.kernel_launches += 1
        return ops.image_to_uint8(self._decode_nhwc(codebook_indices, shape))

    @torch.no_grad()
    def encode_latents(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """``self.encoder(pixel_values)`` of the reference: fp32 (B, 13, H/16, W/16), before the sign test."""
        self.kernel_launches += 1
        return ops.nhwc_to_nchw(self._encode_nhwc(pixel_values))

    @torch.no_grad()
    def get_code(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """(B, 3, H, W) fp32 pixels -> (B, H/16 * W/16) int64 code ids (modeling_magvitv2.py:423-427):
        bit k of the code = [latent channel k > 0], MSB first."""
        z = self.encode_latents(pixel_values)
        B = z.shape[0]
        self.kernel_launches += 1
        return self.quantize.get_indices(z).reshape(B, -1)
