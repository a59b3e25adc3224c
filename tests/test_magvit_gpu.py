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
    with pytest.raises(NotImplementedError):
        vq.get_code(pix)
