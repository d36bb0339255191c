"""chunk_data / merge_data -- drop-ins for the codec-path helpers of the reference's
lossycomp/dataLoader.py:437-532, executed on the GPU (lc_chunk_gather / lc_chunk_merge).

Inside compress()/decompress() these index maps are fused into the encoder's loader and the
decoder's epilogue; the standalone functions exist for callers that used them directly
(notebooks/Chunking_data.ipynb) and for the parity tests.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N


def _dev(a):
    torch = N.require_cuda()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def chunk_data(data, size=(16, 48, 48)):
    """(time, lat, lon, attr) -> (N, 16, 48, 48, attr); only the reference's (16,48,48) chunk."""
    assert tuple(size[:3]) == (16, 48, 48), "the B200 codec is built for (16, 48, 48) chunks"
    assert (data.shape[0] >= size[0] and data.shape[1] >= size[1] and data.shape[2] >= size[2]), \
        "Chunks size cannot be larger than the data size."
    torch = N.require_cuda()
    lib = N.lib()
    T, H, W, Cc = data.shape
    g = (C.c_int * 3)()
    N.check(lib.lc_chunk_grid(T, H, W, C.byref(g)), "lc_chunk_grid")
    x = _dev(data)
    out = torch.empty((g[0] * g[1] * g[2], 16, 48, 48, Cc), dtype=torch.float32, device=x.device)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    N.check(lib.lc_chunk_gather(C.c_void_p(x.data_ptr()), T, H, W, Cc, C.c_void_p(out.data_ptr()), s),
            "lc_chunk_gather")
    return out.cpu().numpy().astype(data.dtype, copy=False)


def merge_data(data, size):
    """(N, 16, 48, 48, attr) -> (time, lat, lon, attr) of shape size[:3] + (attr,)."""
    torch = N.require_cuda()
    lib = N.lib()
    T, H, W = (int(v) for v in size[:3])
    Cc = data.shape[4]
    x = _dev(data)
    out = torch.empty((T, H, W, Cc), dtype=torch.float32, device=x.device)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    N.check(lib.lc_chunk_merge(C.c_void_p(x.data_ptr()), T, H, W, Cc, C.c_void_p(out.data_ptr()), s),
            "lc_chunk_merge")
    return out.cpu().numpy().astype(data.dtype, copy=False)
