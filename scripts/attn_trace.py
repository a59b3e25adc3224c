"""Debug: clock64 timeline of CTA 0 of the head_dim-128 attention kernel (attention_duo.cu; needs the trace build:
nvcc ... -DMMADA_ATT_TRACE -> mmada_b200/csrc/build/libattn_trace.so, see scripts/attn_trace.sh).
Columns per key tile g of CTA 0 (clocks since the first event):
  softmax slot s: wait_begin | scores ready | loaded | max (+ rescale) done | exponentials done | P stored + arrived
  mma slot s:     PV(g): wait_begin | P and V ready | issued ;  S(g): operands ready | issued"""
import ctypes, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = ctypes.CDLL(os.path.join(os.path.dirname(__file__), "..", "mmada_b200", "csrc", "build", os.environ.get("ATT_TRACE_LIB", "libattn_trace.so")))
B, L, H, hd = 16, 1539, 32, 128
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * L, 3 * H * hd, device="cuda", generator=g).bfloat16()
out = torch.empty(B * L, H * hd, device="cuda", dtype=torch.bfloat16)
trace = torch.zeros(4 * 64 * 8, dtype=torch.int64, device="cuda")
lib.mmada_attention_duo_set_trace(ctypes.c_void_p(trace.data_ptr()))
lib.mmada_attention_quad_set_trace(ctypes.c_void_p(trace.data_ptr()))
lib.mmada_attention_duo64_set_trace(ctypes.c_void_p(trace.data_ptr()))
lib.mmada_attention_pair64_set_trace(ctypes.c_void_p(trace.data_ptr()))
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
print("g | softmax0: wait sready loaded max exp arrived ofull epi_done | softmax1: ... | mma0: pv_wait pv_ready pv_issued s_ready s_issued | mma1: ...")
for t in range(int(sys.argv[1]) if len(sys.argv) > 1 else 42):
    r = lambda role, evs: " ".join(f"{int(tr[role, t, e]) - t0:7d}" if tr[role, t, e] > 0 else "      -" for e in evs)
    print(f"{t:2d} | {r(0, range(8))} | {r(1, range(6))} | {r(2, range(5))} | {r(3, range(5))}")
