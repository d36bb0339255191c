"""Decompressor -- drop-in for the reference's lossycomp/decompress.py.

    decompress(compressed_data, verbose=False) -> np.ndarray float32 (time, lat, lon, 1)   (decompress.py:21)

Reads the 9-slot blob of compress.py:168 / decompress.py:54.  Slots 6/7 may be the reference's bz2
streams (first-differenced int32 / int8) or this package's 'LCH1' Huffman containers; slot 0 must be an
'LCL1' latent container written by this package (it names the decoder precision the blob needs).
"""
from __future__ import annotations

import bz2
import pickle

import numpy as np

from . import _hostio
from . import _native as N
from ._engine import HUFF_MAGIC, Engine
from .dist import MULTI_MAGIC


def decompress_device(eng: Engine, slots, *, inject=None, batch=None, rows=None):
    """Returns the reconstruction as a torch float32 CUDA tensor (T,H,W).  `rows=(t0, t1)` (a time
    shard in whole 16-step chunks, see dist.shard_time) decodes only those rows."""
    torch = eng.torch
    comp_data, lshape, array_size, error_shape, mean, std, comp_error, comp_mask, thr2 = slots
    T, H, W = (int(v) for v in array_size[:3])
    latent, meta = eng.unpack_latent(comp_data, lshape)                              # decompress.py:58-60
    T_all = T
    if rows is not None:
        t0, t1 = rows
        assert t0 % 16 == 0 and (t1 % 16 == 0 or t1 == T_all) and t1 - t0 >= 16
        per_tc = latent.shape[0] // (-(-T_all // 16))
        c0 = (t0 // 16) * per_tc
        c1 = c0 + (-(-(t1 - t0) // 16)) * per_tc
        latent = latent[c0:c1].contiguous()
        T = t1 - t0
    if latent.shape[1] != eng.latent_floats:
        raise N.NativeError(f"latent has {latent.shape[1]} values per chunk, the engine's decoder expects "
                            f"{eng.latent_floats}")
    if inject and "recon_std" in inject:
        recon = inject["recon_std"]
    else:
        recon = eng.decode(latent, T, H, W, batch=batch)                             # decompress.py:66-72
    n = T_all * H * W
    huff_mask = comp_mask[:4] in (HUFF_MAGIC, MULTI_MAGIC)
    side_done = None
    if huff_mask and comp_error[:4] in (HUFF_MAGIC, MULTI_MAGIC):
        # The two Huffman streams are independent and each decode kernel occupies a fraction of the GPU (one thread
        # per 1024-symbol segment, a serial chain per thread): the mask goes to the side stream once the decoder's
        # convolutions are through (beside them it would cost the convolution CTAs their residency), the codes stay
        # on the main stream, and reconstruct() waits for both.
        main = torch.cuda.current_stream(eng.dev)
        side = eng.side_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            mask = eng.unpack_mask(comp_mask)
            side_done = torch.cuda.Event()
            side_done.record(side)
        mask.record_stream(main)
    if comp_error[:4] in (HUFF_MAGIC, MULTI_MAGIC):
        q = eng.unpack_codes(comp_error)
    else:                                                                            # decompress.py:75-78
        dq = np.frombuffer(bz2.decompress(comp_error), dtype=np.int32).reshape(error_shape)
        q = eng.cumsum(torch.from_numpy(dq.copy()).to(eng.dev))
    if side_done is not None:
        torch.cuda.current_stream(eng.dev).wait_event(side_done)
    elif huff_mask:
        mask = eng.unpack_mask(comp_mask)
    else:                                                                            # decompress.py:81-84
        dm = np.frombuffer(bz2.decompress(comp_mask), dtype=np.int8).reshape((n,))
        mask = eng.cumsum(torch.from_numpy(dm.copy()).to(eng.dev))
    if q.numel() != int(error_shape[0]) or mask.numel() != n:
        raise N.NativeError("corrupt blob: the decoded streams do not have the lengths the blob declares")
    if rows is not None:                 # this shard's part of the global mask / code streams
        lo, hi = t0 * H * W, t1 * H * W
        off = int((mask[:lo] > 0).sum().item())
        cnt = int((mask[lo:hi] > 0).sum().item())
        mask = mask[lo:hi].contiguous()
        q = q[off:off + cnt].contiguous()
    return eng.reconstruct(recon, mask, q, mean, std, thr2)                          # decompress.py:86-104


def _open_blob(compressed_data, precision=None, model=None):
    """-> (slot list, precision, model path) the blob asks for (decompress.py:54 + the 'LCL1' header of slot 0)."""
    slots = pickle.loads(compressed_data)                                            # decompress.py:54
    try:
        meta = _peek_meta(slots[0])
    except Exception:
        meta = {}
    if precision is None:
        precision = meta.get("precision")
    if model is None and meta.get("model"):
        # the blob names the architecture it was written with: use the shipped weights of that name when there are
        # any (Engine.check_latent_meta then verifies architecture and weight hash, and refuses a mismatch)
        from .models import default_model_path, shipped_model_path
        cand = shipped_model_path(meta["model"])
        if cand and cand != default_model_path():
            model = cand
    return slots, precision, model


def _decompress_on(eng: Engine, slots, out=None):
    T, H, W = (int(v) for v in slots[2][:3])
    pending = ()
    if out is None:            # fresh result: its page faults are taken by the thread pool while the GPU decodes
        out, pending = _hostio.prefault((T, H, W, 1))
    dev_out = decompress_device(eng, slots)
    res = _hostio.d2h(eng, dev_out, (T, H, W, 1), out=out, wait=pending)             # decompress.py:104-108
    eng.check_reconstruct()              # the copy synchronised the stream: outlier count of the mask == code count
    return res


def decompress(compressed_data, verbose=False, *, precision=None, model=None, device=None, out=None):
    """Decompression method (reference: decompress.py:21).  Returns the reconstructed data.
    `out` (optional, float32, T*H*W elements) is filled and returned instead of a new array."""
    N.require_cuda()
    slots, precision, model = _open_blob(compressed_data, precision, model)
    res = _decompress_on(Engine.get(model, precision, device), slots, out)
    if verbose:
        print("Done.")
    return res


def decompress_stream(blobs, *, precision=None, model=None, device=None, lanes=None, outs=None):
    """Decompress a sequence of blobs, yielding one array per blob, in order (each identical to decompress(blob)).

    Bulk-job counterpart of compress_stream(): consecutive blobs run in alternating lanes (lossycomp.lanes), so the
    host work of one blob -- unpickling, Huffman decode tables, the upload of its payloads and above all the
    device->host copy of its 100 MB result -- runs behind the decoder kernels of the other.  All blobs must come
    from the same model and precision (the first one decides; a mismatch raises like in decompress()).
    outs: optional sequence of float32 arrays to fill, one per blob."""
    from collections import deque
    from .lanes import LaneSet
    N.require_cuda()
    ls = None
    outs = iter(outs) if outs is not None else None

    def job(lane, slots, out):
        return _decompress_on(lane.eng, slots, out)

    inflight = deque()
    for blob in blobs:
        slots, p, m = _open_blob(blob, precision, model)
        if ls is None:
            precision, model = p, m
            ls = LaneSet.get(model, precision, device, lanes)
        out = None
        if outs is not None:
            try:
                out = next(outs)
            except StopIteration:
                raise ValueError("decompress_stream: fewer `outs` arrays than blobs") from None
        inflight.append(ls.submit(job, slots, out))
        while len(inflight) > len(ls):
            yield inflight.popleft().result()
    while inflight:
        yield inflight.popleft().result()


def _peek_meta(comp_data: bytes) -> dict:
    """Header of the (first) 'LCL1' container of blob slot 0."""
    import json
    import struct
    from .dist import unpack_multipart
    first = unpack_multipart(comp_data)[0]
    (lm,) = struct.unpack("<I", first[4:8])
    return json.loads(bytes(first[8:8 + lm]))
