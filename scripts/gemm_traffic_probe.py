"""One launch each of the config-2 GEMM shapes (for ncu DRAM-traffic probes under different MMADA_GEMM_GROUP_M /
MMADA_GEMM_HINTS settings): gate_up (SwiGLU), ff_out (+residual), qkv, attn_out."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmada_b200 import ops
M, d, f = 24624, 4096, 12288
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, d, device="cuda", generator=g).bfloat16()
hbuf = torch.randn(M, f, device="cuda", generator=g).bfloat16()
w_gu = torch.randn(2 * f, d, device="cuda", generator=g).bfloat16() * d ** -0.5
w_dn = torch.randn(d, f, device="cuda", generator=g).bfloat16() * f ** -0.5
w_qkv = torch.randn(3 * d, d, device="cuda", generator=g).bfloat16() * d ** -0.5
w_o = torch.randn(d, d, device="cuda", generator=g).bfloat16() * d ** -0.5
res = torch.zeros(M, d, device="cuda")
for _ in range(int(os.environ.get("REPS", "2"))):
    ops.gemm(x, w_gu, ops.EPI_SWIGLU_BF16, out=hbuf)
    ops.gemm(hbuf, w_dn, ops.EPI_RESID_F32, out=res, aux=res)
    ops.gemm(x, w_qkv, ops.EPI_BF16)
    ops.gemm(x, w_o, ops.EPI_RESID_F32, out=res, aux=res)
torch.cuda.synchronize()
