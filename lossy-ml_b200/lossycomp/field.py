"""ERA5-shaped driver around the codec: a `(time, level, lat, lon)` variable is
`level` independent `compress()` problems -- "level" never enters the model (compress.py:55 takes
`(time, lat, lon, 2)` with the land-sea mask as channel 1, testing_compressor.py:53-59).  The container is a
pickled dict of per-level blobs, each a regular 9-slot blob that `decompress()` reads on its own.
"""
from __future__ import annotations

import pickle
from typing import Sequence

import numpy as np

from . import _native as N
from ._engine import Engine
from .compress import compress_device
from .decompress import decompress_device

FIELD_MAGIC = "lossycomp-field-v1"


def compress_field(var, lsm, abs_error, *, levels: Sequence[int] | None = None, precision=None, model=None,
                   device=None, mode="strict", codec="huff") -> bytes:
    """var: (T, L, H, W) float array; lsm: (H, W) land-sea mask in [0, 1]; abs_error: scalar or one value per
    level.  Returns the pickled container."""
    torch = N.require_cuda()
    var = np.asarray(var)
    assert var.ndim == 4, "expected (time, level, lat, lon)"
    T, L, H, W = var.shape
    levels = list(range(L)) if levels is None else list(levels)
    errs = np.broadcast_to(np.asarray(abs_error, dtype=np.float64), (len(levels),))
    eng = Engine.get(model, precision, device)
    lsm_d = torch.from_numpy(np.ascontiguousarray(lsm, dtype=np.float32)).to(eng.dev)
    x = torch.empty((T, H, W, 2), dtype=torch.float32, device=eng.dev)
    x[..., 1] = lsm_d
    blobs, infos = {}, {}
    for lev, err in zip(levels, errs):
        x[..., 0] = torch.from_numpy(np.ascontiguousarray(var[:, lev], dtype=np.float32)).to(eng.dev)
        slots, info = compress_device(eng, x, float(err), mode=mode, codec=codec)
        blobs[lev] = pickle.dumps(slots)
        infos[lev] = {k: v for k, v in info.items() if not hasattr(v, "shape")}
    return pickle.dumps({"magic": FIELD_MAGIC, "shape": (T, L, H, W), "levels": levels, "blobs": blobs,
                         "info": infos})


def decompress_field(container: bytes, *, precision=None, model=None, device=None) -> np.ndarray:
    """Inverse of compress_field: (T, len(levels), H, W) float32."""
    N.require_cuda()
    c = pickle.loads(container)
    assert c.get("magic") == FIELD_MAGIC, "not a lossycomp field container"
    T, _, H, W = c["shape"]
    out = np.empty((T, len(c["levels"]), H, W), dtype=np.float32)
    eng = None
    for i, lev in enumerate(c["levels"]):
        slots = pickle.loads(c["blobs"][lev])
        if eng is None:
            from .decompress import _peek_meta
            eng = Engine.get(model, precision or _peek_meta(slots[0]).get("precision"), device)
        out[:, i] = decompress_device(eng, slots).cpu().numpy()
    return out
