import os, sys
sys.path.insert(0, os.getcwd())
import torch
from mmada_b200 import ops
for (B, H, L) in ((1, 2, 5000), (1, 1, 8195), (2, 3, 4096)):
    hd = 128; d = H * hd
    g = torch.Generator(device="cuda").manual_seed(L)
    qkv = torch.randn(B * L, 3 * d, device="cuda", generator=g).bfloat16()
    out = ops.attention(qkv, B, L, H, hd)
    q, k, v = (qkv[:, i * d:(i + 1) * d].float().view(B, L, H, hd).transpose(1, 2) for i in range(3))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * L, d)
    rel = float((out.float() - ref).abs().max() / ref.abs().max())
    print(B, H, L, f"max-normalised error {rel:.3e}", "OK" if rel < 1.5e-2 else "FAIL")
