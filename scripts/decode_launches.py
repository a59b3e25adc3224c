"""decode_code_uint8 of 8 x 1024 codes (512 x 512 images) — for an ncu launch list / CUDA-event timing of the decoder."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmada_b200.modeling_magvitv2 import MAGVITv2
vq = MAGVITv2(device="cuda").init_random(seed=7)
codes = torch.randint(0, 8192, (8, 1024), device="cuda")
for _ in range(2):
    out = vq.decode_code_uint8(codes)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = vq.decode_code_uint8(codes)
e1.record()
torch.cuda.synchronize()
print("decode_code_uint8 8x512x512:", e0.elapsed_time(e1) / 5, "ms")
