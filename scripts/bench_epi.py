import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmada_b200 import ops
from bench_kernels import timeit
M, N, K = 24624, 4096, 4096
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
ws = [torch.randn(N, K, device="cuda", generator=g).bfloat16() * K ** -0.5 for _ in range(3)]
for name, epi, dt in (("bf16", ops.EPI_BF16, torch.bfloat16), ("f32", ops.EPI_F32, torch.float32), ("resid_inplace", ops.EPI_RESID_F32, torch.float32),
                      ("resid_separate", ops.EPI_RESID_F32, torch.float32)):
    out = torch.zeros(M, N, device="cuda", dtype=dt)
    aux = None
    if name == "resid_inplace":
        aux = out
    if name == "resid_separate":
        aux = torch.zeros(M, N, device="cuda", dtype=dt)
    i = [0]
    def run():
        ops.gemm(a, ws[i[0] % 3], epi, out=out, aux=aux); i[0] += 1
    ms = timeit(run)
    print(name, f"{ms:.3f} ms", f"{2.0*M*N*K/ms/1e9:.0f} TF")
