"""Debug: clock64 timeline of CTA 0 of the attention kernel (needs the trace build:
nvcc ... -DMMADA_ATT_TRACE -> mmada_b200/csrc/build/libattn_trace.so, see scripts/attn_trace.sh)."""
import ctypes, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = ctypes.CDLL(os.path.join(os.path.dirname(__file__), "..", "mmada_b200", "csrc", "build", "libattn_trace.so"))
B, L, H, hd = 16, 1539, 32, 128
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * L, 3 * H * hd, device="cuda", generator=g).bfloat16()
out = torch.empty(B * L, H * hd, device="cuda", dtype=torch.bfloat16)
trace = torch.zeros(4 * 64 * 8, dtype=torch.int64, device="cuda")
lib.mmada_attention_set_trace(ctypes.c_void_p(trace.data_ptr()))
lib.mmada_attention_pair_set_trace(ctypes.c_void_p(trace.data_ptr()))
f = lib.mmada_attention_bf16
f.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_int] * 4 + [ctypes.c_float, ctypes.c_void_p]
d = H * hd
for _ in range(3):
    st = f(qkv.data_ptr(), qkv.data_ptr() + 2 * d, qkv.data_ptr() + 4 * d, 3 * d, out.data_ptr(), d, B, L, H, hd, hd ** -0.5,
           torch.cuda.current_stream().cuda_stream)
    assert st == 0
torch.cuda.synchronize()
tr = trace.cpu().view(4, 64, 8)
t0 = int(tr[tr > 0].min())
T = 40
print("t | softmax0: wait_begin s_ready ld_done max_done exp_done st_done arrived | softmax1 ... | mma: (wait0 got0 issued0) (wait1 got1 issued1)")
for t in range(T):
    r = lambda role, evs: " ".join(f"{int(tr[role, t, e]) - t0:7d}" if tr[role, t, e] > 0 else "      -" for e in evs)
    print(f"{t:2d} | {r(0, range(7))} | {r(1, range(7))} | {r(2, [0, 1, 2, 3])} | issue_s: {r(3, [0, 1, 2, 3])}")
