"""ctypes binding of liblossycomp_b200.so (include/lossycomp_b200.h).

There is no CPU fallback: if the CUDA library is missing or no CUDA device is present the
codec raises -- the product path is the GPU path or nothing.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LOSSYCOMP_LIB") or os.path.join(_HERE, "liblossycomp_b200.so")   # LOSSYCOMP_LIB: experiment builds

PREC_FP32, PREC_BF16, PREC_FP16 = 0, 1, 2
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "fp16": PREC_FP16}


class NativeError(RuntimeError):
    pass


class LcLayer(C.Structure):
    _fields_ = [("kind", C.c_int32), ("k", C.c_int32), ("stride", C.c_int32), ("cin", C.c_int32),
                ("cout", C.c_int32), ("relu", C.c_int32), ("res_save", C.c_int32), ("res_add", C.c_int32),
                ("h_kernel", C.c_void_p), ("h_bias", C.c_void_p)]


_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_f = C.c_float
_d = C.c_double
_sz = C.c_size_t

# name -> (restype, argtypes); mirrors include/lossycomp_b200.h one to one
SIGNATURES = {
    "lc_last_error": (C.c_char_p, []),
    "lc_version": (_i, []),
    "lc_create": (_i, [C.POINTER(LcLayer), _i, _i, _i, _i, _i, C.POINTER(_p)]),
    "lc_destroy": (_i, [_p]),
    "lc_latent_floats": (_i, [_p]),
    "lc_layer_backend": (_i, [_p, _i]),
    "lc_profile_enable": (_i, [_p, _i]),
    "lc_profile_read": (_i, [_p, _p, _p, _i]),
    "lc_chunk_grid": (_i, [_i, _i, _i, C.POINTER(_i * 3)]),
    "lc_chunk_window": (_i, [_i, _i, _i, _i, C.POINTER(_i * 6)]),
    "lc_stats_workspace_bytes": (_sz, []),
    "lc_stats": (_i, [_p, _i64, _i, _p, _p, _p]),
    "lc_stats_blocks": (_i, [_p, _i64, _i64, _i, _p, _p, _p]),
    "lc_ae_workspace_bytes": (_sz, [_p, _i]),
    "lc_ae_workspace_init": (_i, [_p, _p, _sz, _p]),
    "lc_ae_workspace_release": (_i, [_p, _p]),
    "lc_encode": (_i, [_p, _p, _i, _i, _i, _i, _f, _f, _i, _i, _p, _p, _sz, _p]),
    "lc_decode": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "lc_scan_workspace_bytes": (_sz, [_i64]),
    "lc_quantize": (_i, [_p, _i, _p, _i64, _f, _f, _f, _f, _p, _p, _p, _p, _p]),
    "lc_quantize_checked": (_i, [_p, _i, _p, _i64, _f, _f, _f, _f, _d, _d, _p, _p, _p, _p, _p]),
    "lc_latent_quantize": (_i, [_p, _i64, _p, _p, _p]),
    "lc_latent_dequantize": (_i, [_p, _p, _i64, _p, _p]),
    "lc_diff_i32": (_i, [_p, _i64, _p, _p]),
    "lc_diff_i8": (_i, [_p, _i64, _p, _p]),
    "lc_cumsum_i32": (_i, [_p, _i64, _p, _p, _p]),
    "lc_cumsum_i8": (_i, [_p, _i64, _p, _p, _p]),
    "lc_reconstruct": (_i, [_p, _p, _p, _i64, _i64, _f, _f, _d, _p, _p, _p]),
    "lc_verify_bound": (_i, [_p, _i, _p, _i64, _d, _p, _p, _p]),
    "lc_hist_u8": (_i, [_p, _i64, _p, _p]),
    "lc_hist_codes": (_i, [_p, _i64, _i, _i, _i, _p, _p]),
    "lc_symbolize_i32": (_i, [_p, _i64, _i, _i, _i, _p, _p, _p, _p, _p]),
    "lc_unsymbolize_i32": (_i, [_p, _i64, _i, _i, _i, _p, _p, _p, _p]),
    "lc_symbolize_plain_i32": (_i, [_p, _i64, _i, _i, _p, _p]),
    "lc_pack_trits": (_i, [_p, _i64, _p, _p]),
    "lc_pack_trits_hist": (_i, [_p, _i64, _p, _p, _p]),
    "lc_unpack_trits": (_i, [_p, _i64, _p, _p]),
    "lc_rle_workspace_bytes": (_sz, [_i64]),
    "lc_rle_tokens": (_i, [_p, _i64, _i, _i, _p, _p, _p, _p, _p, _p]),
    "lc_rle_expand": (_i, [_p, _i64, _i, _i, _p, _i64, _p, _p]),
    "lc_huff_build_table": (_i, [_p, _i, _p, _p, _p, _p]),
    "lc_huff_build_decode": (_i, [_p, _p, _p, _i, _p, _p, _p]),
    "lc_huff_max_bytes": (_sz, [_i64]),
    "lc_huff_workspace_bytes": (_sz, [_i64, _i]),
    "lc_huff_encode": (_i, [_p, _i64, _p, _p, _i, _p, _p, _p, _p, _p]),
    "lc_huff_decode": (_i, [_p, _i64, _i64, _i, _p, _p, _i, _p, _p, _p]),
    "lc_synth_field": (_i, [_p, _i, _i, _i, _i, _i, C.c_uint64, _p]),
    "lc_chunk_gather": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "lc_chunk_merge": (_i, [_p, _i, _i, _i, _i, _p, _p]),
}

_lib = None


def lib():
    """Load the shared library (once).  Raises NativeError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C lossy-ml_b200/csrc). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().lc_last_error().decode(errors="replace")
        if rc == -1 and "Chunks size cannot be larger" in msg:
            raise AssertionError(msg.split(" ", 1)[-1])     # dataLoader.py:447
        raise NativeError(f"{what} failed ({rc}): {msg}")


_malloc_tuned = False


def tune_host_malloc():
    """Keep freed multi-megabyte blocks inside the process heap (glibc mallopt).

    The codec returns its payloads as Python `bytes`; a fresh 8 MB allocation served by mmap costs
    2.5-3.9 ms of first-touch page faults per copy on the benchmark box against 0.45 ms into recycled
    heap pages (tools/host_copy_probe.py) -- a quarter of the whole compress step.  Raising
    M_MMAP_THRESHOLD to its 32 MB maximum and M_TRIM_THRESHOLD / M_TOP_PAD lets those allocations
    reuse warm pages.  Process-wide and glibc-only; LOSSYCOMP_NO_MALLOPT=1 leaves malloc untouched."""
    global _malloc_tuned
    if _malloc_tuned or os.environ.get("LOSSYCOMP_NO_MALLOPT") == "1":
        return
    _malloc_tuned = True
    try:
        libc = C.CDLL("libc.so.6")
        M_TRIM_THRESHOLD, M_TOP_PAD, M_MMAP_THRESHOLD = -1, -2, -3
        libc.mallopt(M_MMAP_THRESHOLD, 32 << 20)
        libc.mallopt(M_TRIM_THRESHOLD, 1 << 30)
        libc.mallopt(M_TOP_PAD, 64 << 20)
    except Exception:
        pass


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise NativeError("lossycomp (B200 build) needs a CUDA device; there is no CPU fallback.")
    tune_host_malloc()
    return torch
