"""Motion VQ-VAE token -> pose path on the GPU: the gather kernel against PyTorch, the
decoder against the output of the REFERENCE's own Decoder class (golden fixture), batched against one-by-one."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("taps,dil,up,relu", [(3, 1, 1, False), (3, 9, 1, True), (3, 3, 1, True), (1, 1, 1, True), (3, 1, 2, False)])
def test_conv1d_gather_gemm_equals_conv1d(taps, dil, up, relu):
    from mmada_b200 import ops
    B, T, C, Co = 3, 37, 512, 263
    g = torch.Generator(device="cuda").manual_seed(taps * 100 + dil + up)
    x = torch.randn(B, T, C, device="cuda", generator=g)
    w = torch.randn(Co, C, taps, device="cuda", generator=g) / (C * taps) ** 0.5
    b = torch.randn(Co, device="cuda", generator=g)
    a = ops.conv1d_gather(x, taps, dil, up, relu)
    assert a.shape == (B, T * up, taps * C)
    out = ops.gemm(a.view(-1, taps * C), w.permute(0, 2, 1).reshape(Co, -1).bfloat16().contiguous(), ops.EPI_BIAS_F32, bias=b)
    xin = x.permute(0, 2, 1)
    if up == 2:
        xin = F.interpolate(xin, scale_factor=2, mode="nearest")
    if relu:
        xin = F.relu(xin)
    ref = F.conv1d(xin.bfloat16().float(), w.bfloat16().float(), b, padding=dil * (taps // 2), dilation=dil).permute(0, 2, 1)
    assert _rel(out.view(B, T * up, Co), ref) < 2e-5       # exact bf16 products, fp32 accumulation: summation order only
    y = torch.randn(8, 100, device="cuda", generator=g)
    assert torch.equal(ops.relu_(y.clone()), F.relu(y))


def test_forward_decoder_vs_reference_golden(golden):
    from mmada_b200.motion_vqvae import HumanVQVAE
    from oracle import motion
    gd = golden("motion_decoder")
    sd = motion.make_motion_decoder_weights(0)
    vq = HumanVQVAE().load_state_dict(sd)
    ids = torch.from_numpy(gd["ids"]).cuda()
    pose = vq.forward_decoder(ids)
    ref = torch.from_numpy(gd["pose"])
    assert pose.shape == ref.shape == (1, 196, 263) and pose.dtype == torch.float32
    err = float((pose.cpu() - ref).abs().max()) / float(gd["pose_absmax"])
    rms = float((pose.cpu() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
    print(f"motion forward_decoder: max|d|/max|ref| = {err:.3e}, rel-rms = {rms:.3e}")
    # bf16 conv operands (fp32 accumulate, fp32 trunk) through 18 conv layers vs the fp32 reference
    assert err < 2e-2 and rms < 1e-2
    # the reference's contract: a (B, T) input is ONE sequence of B*T tokens
    two = torch.cat([ids, ids], 0)
    assert vq.forward_decoder(two).shape == (1, 2 * 196, 263)
    # batched decode = every row on its own
    g = torch.Generator().manual_seed(5)
    batch = torch.randint(0, 512, (5, 32), generator=g).cuda()
    out = vq.forward_decoder_batched(batch)
    assert out.shape == (5, 128, 263)
    for i in range(5):
        assert torch.equal(out[i:i + 1], vq.forward_decoder(batch[i:i + 1]))
