import os, sys
sys.path.insert(0, os.getcwd())
import torch
from mmada_b200 import ops
R, V = 512, 134656
g = torch.Generator(device="cuda").manual_seed(0)
lg = torch.randn(R, V, device="cuda", generator=g) * 3
un = torch.randn(R, V, device="cuda", generator=g) * 3
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, fn, nbytes in (("T=0 no cfg", lambda: ops.text_sample_rows(lg, None, 0.0, 0.0, None), R * V * 4),
                 ("T=0 cfg", lambda: ops.text_sample_rows(lg, un, 1.5, 0.0, None), 2 * R * V * 4),
                 ("T=1 philox", lambda: ops.text_sample_rows(lg, None, 0.0, 1.0, None, seed=1), R * V * 4)):
    ms = t(fn)
    print(f"text_sample_rows {name}: {ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s")
