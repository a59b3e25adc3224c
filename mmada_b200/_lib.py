"""ctypes binding of libmmada_b200.so (the C ABI declared in include/mmada_b200.h).

There is no fallback: if the shared library is missing or a call fails, this raises.  Build it with
``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C mmada_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
#: MMADA_B200_LIB: another build of the same library (A/B runs of kernel variants, scripts/ab_attention.sh)
LIB_PATH = os.environ.get("MMADA_B200_LIB") or os.path.join(_HERE, "libmmada_b200.so")

_lib: Optional[C.CDLL] = None

_p, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

#: name -> argtypes (restype is always int); mirrors include/mmada_b200.h
SIGNATURES = {
    "mmada_abi_version": [],
    "mmada_device_arch": [],
    "mmada_gemm_bf16": [_p, _i64, _p, _i64, _p, _i64, _p, _p, _i, _i, _i, _i, _i, _p],
    "mmada_gemm_qkv_rope_bf16": [_p, _i64, _p, _i64, _p, _i64, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "mmada_gemm_resid_norm_f32": [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _i, _i, _i, _i, _p],
    "mmada_gemm_swiglu_rownorm_bf16": [_p, _i64, _p, _i64, _p, _i64, _p, _i, _i, _f, _i, _i, _i, _i, _p],
    "mmada_gemm_qkv_rope_rownorm_bf16": [_p, _i64, _p, _i64, _p, _i64, _p, _p, _p, _i, _i, _f, _i, _i, _i, _i, _i, _i,
                                         _i, _p],
    "mmada_embed_norm_f32": [_p, _p, _p, _p, _p, _i, _i, _i64, _p],
    "mmada_embed_f32": [_p, _p, _p, _i, _i, _i64, _p],
    "mmada_rmsnorm_bf16": [_p, _p, _p, _p, _i, _i, _f, _p],
    "mmada_gather_rows": [_p, _i64, _p, _p, _i, _i, _p],
    "mmada_cross_entropy_rows_f32": [_p, _i64, _p, _i64, _p, _i, _i, _p],
    "mmada_rope_inplace_bf16": [_p, _i64, _p, _p, _i, _i, _i, _i, _p],
    "mmada_attention_bf16": [_p, _p, _p, _i64, _p, _i64, _i, _i, _i, _i, _f, _p],
    "mmada_t2i_sample_step": [_p, _p, _p, _p, _p, _p, _i64, _i64, _p, _p, _p, _p, _i, _p, _i, _i, _i, _f, _f, _f, _f,
                              _i64, _i64, _p],
    "mmada_compact_masked_rows": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i64, _p],
    "mmada_t2i_sample_step_compact": [_p, _p, _p, _p, _p, _p, _i64, _i64, _p, _p, _p, _i, _p, _i, _i, _i, _f, _f, _f, _f,
                                      _i64, _i64, _p, _p],
    "mmada_mask_by_random_topk": [_p, _p, _p, _p, _i, _i, _f, _p],
    "mmada_text_sample_rows": [_p, _p, _f, _p, C.c_uint64, _f, _i, _i, _p, _p, _p],
    "mmada_block_mask_count": [_p, _i64, _i, _i, _i, _i64, _p, _p],
    "mmada_text_transfer": [_p, _i64, _i, _i, _p, _p, _p, _p, _i, _i, _i, _i64, _p, _p],
    "mmada_conv_nhwc_bf16": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "mmada_lfq_decode_nhwc": [_p, _p, _p, _p, _i, _p],
    "mmada_lfq_indices_to_bits": [_p, _p, _i, _i, _p],
    "mmada_lfq_bits_to_indices": [_p, _p, _i, _i, _p],
    "mmada_groupnorm_stats": [_p, _p, _i, _i, _i, _p],
    "mmada_groupnorm_apply_bf16": [_p, _p, _p, _p, _p, _i, _i, _i, _f, _i, _p],
    "mmada_upsample2x_nhwc_bf16": [_p, _p, _i, _i, _i, _i, _p],
    "mmada_cast_f32_bf16": [_p, _p, _i64, _p],
    "mmada_softmax_rows_bf16": [_p, _p, _i, _i, _f, _p],
    "mmada_nhwc_to_nchw_f32": [_p, _p, _i, _i, _i, _p],
    "mmada_image_to_uint8": [_p, _p, _i64, _p],
    "mmada_image_to_nhwc64_bf16": [_p, _p, _i, _i, _i, _p],
    "mmada_space_to_depth2_bf16": [_p, _p, _i, _i, _i, _i, _p],
    "mmada_conv1d_gather_bf16": [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "mmada_relu_f32": [_p, _i64, _p],
    "mmada_build_prompts": [_p, _p, _p, _i64, _p, _p, _i, _i, _i, _i, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _p],
}


class MMadaKernelError(RuntimeError):
    pass


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise MMadaKernelError(
                f"{LIB_PATH} not found: the CUDA extension is not built (run __graft_entry__.build()). "
                "mmada_b200 has no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            if not hasattr(lib, name):
                continue            # reported by check_symbols(); calling it raises below
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = C.c_int
        _lib = lib
    return _lib


#: profiling hook: when set to a list, every kernel call appends (name, start_event, end_event) recorded on
#: torch's current stream (scripts/bench_step_breakdown.py)
EVENTS = None


def call(name: str, *args) -> None:
    lib = load()
    fn = getattr(lib, name, None)
    if fn is None:
        raise MMadaKernelError(f"{name} is not exported by {LIB_PATH}")
    if EVENTS is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st = fn(*args)
        e1.record()
        EVENTS.append((name, e0, e1))
    else:
        st = fn(*args)
    if st != 0:
        detail = ""
        if st >= 1000:
            try:
                import torch
                detail = f" (cudaError {st - 1000})"
                torch.cuda.synchronize()
            except Exception as e:  # surface the asynchronous error text if there is one
                detail += f": {e}"
        raise MMadaKernelError(f"{name} failed with status {st}{detail}")
