"""MAGVIT-v2 token -> pixel path on the GPU: LFQ maps bit-exact against the oracle, the NHWC
implicit-GEMM convolution / GroupNorm / upsample kernels against PyTorch fp32 on the same bf16-rounded
operands, and ``decode_code`` against the REAL reference's fp32 output (golden fixture)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def test_lfq_maps_bit_exact():
    from mmada_b200.modeling_magvitv2 import MAGVITv2
    from oracle import magvit
    vq = MAGVITv2()
    idx = torch.arange(8192).view(8, 1024)
    bits = vq.quantize.get_codebook_entry(idx.cuda())
    assert np.array_equal(bits.cpu().numpy(), magvit.lfq_indices_to_bits(idx.numpy()))
    back = vq.quantize.get_indices(bits)
    assert torch.equal(back.cpu().view(8, 1024), idx)
    z = torch.randn(2, 13, 4, 4)
    assert np.array_equal(vq.quantize.get_indices(z.cuda()).cpu().numpy(), magvit.lfq_bits_to_indices(z.numpy()))


@pytest.mark.parametrize("B,H,W,Ci,Co,taps,resid", [
    (2, 16, 16, 64, 128, 9, False), (1, 32, 32, 512, 512, 9, True), (2, 64, 64, 256, 128, 9, False),
    (1, 128, 128, 128, 128, 9, True), (1, 16, 256, 128, 3, 9, False), (2, 32, 32, 512, 256, 1, False),
    (3, 16, 16, 128, 256, 9, True)])
def test_conv_nhwc(B, H, W, Ci, Co, taps, resid):
    from mmada_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(H + Ci + Co)
    k = 3 if taps == 9 else 1
    x = torch.randn(B, H, W, Ci, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Co, Ci, k, k, device="cuda", generator=g) / math.sqrt(Ci * k * k)).bfloat16()
    bias = torch.randn(Co, device="cuda", generator=g)
    r = torch.randn(B, H, W, Co, device="cuda", generator=g) if resid else None
    wk = w.permute(0, 2, 3, 1).reshape(Co, -1).contiguous()
    out = ops.conv_nhwc(x, wk, bias, taps, ops.EPI_BIAS_RESID_F32 if resid else ops.EPI_BIAS_F32, resid=r)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=k // 2).permute(0, 2, 3, 1)
    if resid:
        ref = ref + r
    assert out.shape == ref.shape
    assert _rel(out, ref) < 2e-5            # exact bf16 products, fp32 accumulation: summation order only


@pytest.mark.parametrize("C,H", [(128, 32), (256, 16), (512, 8)])
def test_groupnorm_swish_upsample(C, H):
    from mmada_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(C)
    B = 2
    x = torch.randn(B, H, H, C, device="cuda", generator=g) * 2 + 0.5
    gamma = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
    beta = 0.1 * torch.randn(C, device="cuda", generator=g)
    sums = torch.empty(B, 32, 2, device="cuda", dtype=torch.float64)
    ref = F.group_norm(x.permute(0, 3, 1, 2), 32, gamma, beta, eps=1e-6)
    out = ops.groupnorm_swish(x, gamma, beta, sums, swish=False)
    assert _rel(out.float().permute(0, 3, 1, 2), ref) < 5e-3          # bf16 output rounding
    out = ops.groupnorm_swish(x, gamma, beta, sums, swish=True)
    assert _rel(out.float().permute(0, 3, 1, 2), ref * torch.sigmoid(ref)) < 5e-3
    up = ops.upsample2x_nhwc(x)
    refu = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up, refu.bfloat16())
    assert torch.equal(ops.cast_bf16(x), x.bfloat16())


def test_softmax_rows_and_uint8():
    from mmada_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    s = torch.randn(300, 1024, device="cuda", generator=g) * 20
    out = ops.softmax_rows_bf16(s, 0.0442)
    assert _rel(out.float(), torch.softmax(s * 0.0442, -1)) < 5e-3
    img = torch.randn(2, 8, 8, 3, device="cuda", generator=g)
    ref = (torch.clamp((img + 1.0) / 2.0, 0.0, 1.0) * 255.0).cpu().numpy().astype(np.uint8)
    assert np.array_equal(ops.image_to_uint8(img).cpu().numpy(), ref)
    assert torch.equal(ops.nhwc_to_nchw(img), img.permute(0, 3, 1, 2).contiguous())


def test_decode_code_vs_reference_golden(golden):
    """decode_code on the synthetic decoder weights against the real reference's fp32 output."""
    from mmada_b200.modeling_magvitv2 import MAGVITv2
    from oracle import weights as W
    gd = golden("magvit")
    vq = MAGVITv2().load_state_dict(W.make_vq_decoder_weights(0))
    idx = torch.from_numpy(gd["idx_16x16"]).cuda()
    pix = vq.decode_code(idx)
    assert pix.shape == (1, 3, 256, 256) and pix.dtype == torch.float32
    ref = torch.from_numpy(gd["pix_16x16"])
    mine = pix.cpu()[:, :, ::2, ::2]
    err = float((mine - ref).abs().max()) / float(gd["pix_absmax_16x16"])
    rms = float((mine - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
    print(f"decode_code 16x16: max|d|/max|ref| = {err:.3e}, rel-rms = {rms:.3e}")
    # bf16 conv operands (fp32 accumulate, fp32 trunk) through 60 conv layers vs the fp32 reference
    assert err < 3e-2 and rms < 2e-2
    u8 = vq.decode_code_uint8(idx)
    exp = (torch.clamp((pix + 1.0) / 2.0, 0.0, 1.0) * 255.0).permute(0, 2, 3, 1).cpu().numpy().astype(np.uint8)
    assert np.array_equal(u8.cpu().numpy(), exp)
    # the post-processed image (inference_t2i.py:123-125) against the reference's own fp32 pixels, in uint8 levels: the
    # reference keeps MAGVIT in fp32, this decoder feeds bf16 conv operands
    ref_u8 = (torch.clamp((ref + 1.0) / 2.0, 0.0, 1.0) * 255.0).permute(0, 2, 3, 1).numpy().astype(np.uint8)
    d_u8 = np.abs(u8.cpu().numpy()[:, ::2, ::2, :].astype(np.int32) - ref_u8.astype(np.int32))
    print(f"decode_code uint8 vs reference fp32 image: max {int(d_u8.max())} levels, mean {d_u8.mean():.3f}, "
          f"identical {float((d_u8 == 0).mean()):.3f}, within 1 level {float((d_u8 <= 1).mean()):.3f}")
    assert int(d_u8.max()) <= 6 and float(d_u8.mean()) < 1.0
    with pytest.raises(Exception, match="no encoder weights"):
        vq.get_code(pix)


def test_encoder_input_and_downsample_kernels():
    """image_to_nhwc64 and the space-to-depth formulation of Downsample (pad (0,1,0,1) + 3x3 stride 2,
    models/common_modules.py:73-90) against PyTorch on the same bf16-rounded operands."""
    from mmada_b200 import ops
    from mmada_b200.modeling_magvitv2 import MAGVITv2
    g = torch.Generator(device="cuda").manual_seed(7)
    px = torch.randn(2, 3, 32, 48, device="cuda", generator=g)
    x64 = ops.image_to_nhwc64(px)
    assert x64.shape == (2, 32, 48, 64)
    assert torch.equal(x64[..., :3], px.permute(0, 2, 3, 1).bfloat16()) and int(x64[..., 3:].abs().sum()) == 0
    B, H, W, C = 2, 64, 64, 128
    x = torch.randn(B, H, W, C, device="cuda", generator=g)
    w = torch.randn(C, C, 3, 3, device="cuda", generator=g) / math.sqrt(9 * C)
    b = torch.randn(C, device="cuda", generator=g)
    vq = MAGVITv2()
    sd = {"encoder.down.0.downsample.conv.weight": w, "encoder.down.0.downsample.conv.bias": b}
    # reuse the loader's weight rearrangement for one Downsample layer
    w2 = torch.zeros((C, 3, 3, 4, C), device="cuda")
    for ky in range(3):
        for kx in range(3):
            w2[:, ky // 2 + 1, kx // 2 + 1, 2 * (ky % 2) + (kx % 2), :] = w[:, :, ky, kx]
    s2d = ops.space_to_depth2(x)
    assert s2d.shape == (B, H // 2, W // 2, 4 * C)
    assert torch.equal(s2d[:, 3, 5, 2 * C:3 * C], x[:, 7, 10].bfloat16())          # sub-pixel (1, 0) of block (3, 5)
    out = ops.conv_nhwc(s2d, w2.reshape(C, -1).bfloat16().contiguous(), b, 9, ops.EPI_BIAS_F32)
    xr = F.pad(x.bfloat16().float().permute(0, 3, 1, 2), (0, 1, 0, 1))
    ref = F.conv2d(xr, w.bfloat16().float(), b, stride=2).permute(0, 2, 3, 1)
    assert out.shape == ref.shape and _rel(out, ref) < 2e-5


def test_get_code_vs_reference_golden(golden):
    """MAGVITv2.get_code on the synthetic encoder weights against the real reference (fp32): latents within the
    stated tolerance; the code ids bit-exact given OUR latents (the LFQ sign test is integer work), and their
    agreement with the reference's ids reported (a bf16-operand latent within 1e-2 of zero may flip its bit)."""
    from mmada_b200.modeling_magvitv2 import MAGVITv2
    from oracle import magvit, weights as W
    gd = golden("magvit_encoder")
    vq = MAGVITv2().load_state_dict(W.make_vq_encoder_weights(0))
    px = torch.from_numpy(gd["pixels"]).float().cuda()
    z = vq.encode_latents(px)
    ref = torch.from_numpy(gd["latents"])
    assert z.shape == ref.shape == (1, 13, 16, 16)
    err = float((z.cpu() - ref).abs().max()) / float(gd["latent_absmax"])
    rms = float((z.cpu() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
    codes = vq.get_code(px)
    assert codes.shape == (1, 256) and codes.dtype == torch.int64
    assert np.array_equal(codes.cpu().numpy(), magvit.lfq_bits_to_indices(z.cpu().numpy()).reshape(1, -1))
    ref_codes = torch.from_numpy(gd["codes"])
    bits_equal = 1.0 - float(((codes.cpu() ^ ref_codes).unsqueeze(-1) >> torch.arange(13) & 1).float().mean())
    print(f"get_code 256x256: latents max|d|/max|ref| = {err:.3e}, rel-rms = {rms:.3e}; "
          f"bit agreement with the fp32 reference {bits_equal:.4f}, id agreement {float((codes.cpu() == ref_codes).float().mean()):.3f}")
    # bf16 conv operands (fp32 accumulate, fp32 trunk) through ~50 conv layers vs the fp32 reference
    assert err < 3e-2 and rms < 2e-2
    assert bits_equal > 0.97
