"""One-off stress run (not part of the suite): the head_dim-128 attention kernel on random (batch, heads, length) shapes
against torch SDPA in fp32 (tolerances of tests/test_kernels_gpu.py::test_attention) + run-to-run determinism.
usage: python scripts/stress_attention_parity.py [n_cases]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmada_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
g = torch.Generator().manual_seed(7)
bad = 0
worst = 0.0
for case in range(n):
    hd = 128
    B = int(torch.randint(1, 5, (1,), generator=g))
    H = int(torch.randint(1, 48, (1,), generator=g))
    L = int(torch.randint(129, 1700, (1,), generator=g)) if case % 4 else [257, 513, 1539, 1025, 383, 641, 897, 1281, 1537, 264][case // 4 % 10]
    d = H * hd
    gg = torch.Generator(device="cuda").manual_seed(case)
    qkv = torch.randn(B * L, 3 * d, device="cuda", generator=gg).bfloat16()
    qkv[:, :d] *= float(torch.rand(1, generator=g)) * 3 + 0.5
    out = ops.attention(qkv, B, L, H, hd)
    out2 = ops.attention(qkv, B, L, H, hd)
    q, k, v = (qkv[:, i * d:(i + 1) * d].float().view(B, L, H, hd).transpose(1, 2) for i in range(3))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * L, d)
    rel = float((out.float() - ref).abs().max() / ref.abs().max())
    rms = float((out.float() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
    worst = max(worst, rel)
    ok = rel < 1.5e-2 and rms < 5e-3 and torch.equal(out, out2) and bool(torch.isfinite(out.float()).all())
    if not ok:
        bad += 1
        print(f"case {case}: B={B} H={H} L={L}: max {rel:.3e} rms {rms:.3e} deterministic {torch.equal(out, out2)}")
print(f"{n} shapes, {bad} failures, worst max-normalised error {worst:.3e}")
sys.exit(1 if bad else 0)
