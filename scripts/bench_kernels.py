"""Micro-benchmarks of the individual kernels at BASELINE config-2 shapes (CUDA events, L2-cold:
operands of consecutive launches rotate through buffers larger than L2)."""
import argparse
import json
import math
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmada_b200 import ops


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="gemm,attn,norm,rope,sample")
    ap.add_argument("--cta-groups", default="2")
    ap.add_argument("--M", type=int, default=16 * 1539)
    args = ap.parse_args()
    M, d, f = args.M, 4096, 12288
    res = {}
    g = torch.Generator(device="cuda").manual_seed(0)
    if "gemm" in args.what:
        shapes = [("qkv", 3 * d, d, ops.EPI_BF16), ("attn_out", d, d, ops.EPI_RESID_F32),
                  ("gate_up", 2 * f, d, ops.EPI_SWIGLU_BF16), ("ff_out", d, f, ops.EPI_RESID_F32),
                  ("head", 8192, d, ops.EPI_F32)]
        for name, N, K, epi in shapes:
            Mx = 16 * 1024 if name == "head" else M
            a = torch.randn(Mx, K, device="cuda", generator=g).bfloat16()
            ws = [torch.randn(N, K, device="cuda", generator=g).bfloat16() * K ** -0.5 for _ in range(3)]
            n_out = N // 2 if epi == ops.EPI_SWIGLU_BF16 else N
            out = torch.empty(Mx, n_out, device="cuda", dtype=torch.float32 if epi in (ops.EPI_F32, ops.EPI_RESID_F32) else torch.bfloat16)
            aux = out if epi == ops.EPI_RESID_F32 else None
            if aux is not None:
                out.zero_()
            for cg in [int(c) for c in args.cta_groups.split(",")]:
                i = [0]
                def run():
                    ops.gemm(a, ws[i[0] % 3], epi, out=out, aux=aux, cta_group=cg); i[0] += 1
                ms = timeit(run)
                res[f"gemm_{name}_cg{cg}"] = dict(ms=ms, tflops=2.0 * Mx * N * K / ms / 1e9)
            ref = timeit(lambda: torch.matmul(a, ws[0].t()))
            res[f"cublas_{name}"] = dict(ms=ref, tflops=2.0 * Mx * N * K / ref / 1e9)
    if "attn" in args.what:
        B, L, H, hd = 16, 1539, 32, 128
        qkv = torch.randn(B * L, 3 * H * hd, device="cuda", generator=g).bfloat16()
        out = torch.empty(B * L, H * hd, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: ops.attention(qkv, B, L, H, hd, out=out))
        fl = 4.0 * L * L * H * hd * B
        res["attention"] = dict(ms=ms, tflops=fl / ms / 1e9)
        q, k, v = (qkv[:, i * H * hd:(i + 1) * H * hd].view(B, L, H, hd).transpose(1, 2) for i in range(3))
        ms = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
        res["torch_sdpa"] = dict(ms=ms, tflops=fl / ms / 1e9)
    if "norm" in args.what:
        x = torch.randn(M, d, device="cuda", generator=g)
        w = torch.ones(d, device="cuda")
        out = torch.empty(M, d, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: ops.rmsnorm(x, w, 1e-5, out=out))
        res["rmsnorm"] = dict(ms=ms, gbs=M * d * 6 / ms / 1e6)
    if "rope" in args.what:
        qkv = torch.randn(M, 3 * d, device="cuda", generator=g).bfloat16()
        fr = torch.rand(2048, 64, device="cuda")
        ms = timeit(lambda: ops.rope_inplace(qkv, fr, fr, d, 128, 1539))
        res["rope"] = dict(ms=ms, gbs=M * 2 * d * 2 * 2 / ms / 1e6)
    if "sample" in args.what:
        B, N, C = 8, 1024, 8192
        cond = torch.randn(B * N, C, device="cuda", generator=g)
        unc = torch.randn(B * N, C, device="cuda", generator=g)
        q = torch.empty(B * N, C, device="cuda").exponential_(1, generator=g)
        u = torch.rand(B, N, device="cuda", generator=g)
        tickets = torch.zeros(B, dtype=torch.int32, device="cuda")
        def run():
            known = torch.full((B, N), 126336, dtype=torch.int64, device="cuda")
            ops.t2i_sample_step(cond, unc, q, u, known, None, 0, tickets, 3.5, 500.0, 0.5, 126336, 126349)
        ms = timeit(run)
        res["t2i_sample_all_masked"] = dict(ms=ms, gbs=3 * B * N * C * 4 / ms / 1e6)
        # the loop's own form: logits of the still-masked positions only, read through the slot map (all masked here)
        known0 = torch.full((B, N), 126336, dtype=torch.int64, device="cuda")
        _, slot = ops.compact_masked_rows(known0, N + 2, 1, N, 2, 126336)
        def run_c():
            known = torch.full((B, N), 126336, dtype=torch.int64, device="cuda")
            ops.t2i_sample_step(cond, unc, q, u, known, None, 0, tickets, 3.5, 500.0, 0.5, 126336, 126349, slot=slot)
        ms = timeit(run_c)
        res["t2i_sample_all_masked_compact"] = dict(ms=ms, gbs=3 * B * N * C * 4 / ms / 1e6)
    for k, v in res.items():
        print(k, json.dumps(v))


if __name__ == "__main__":
    main()
