"""Compressor -- drop-in for the reference's lossycomp/compress.py.

    compress(array, err_threshold, verbose=False) -> bytes            (compress.py:21)

Same call, same 9-slot pickled blob (compress.py:168):
    [comp_data, comp_data_shape, array.shape, error_shape, mean, std, comp_error, comp_mask, 2*thr]
but every stage runs as hand-written sm_100a CUDA behind include/lossycomp_b200.h; there is no
TensorFlow and no CPU fallback.  Keyword-only extensions (defaults keep the reference call working):

  mode   "strict" (default): guard-banded threshold thr' = abs_err - guard so that the hard bound
         |x - xhat| <= abs_err holds on every element after the float32 rounding of the final add
         (the reference itself exceeds it by <= 1/2 ulp); slot 8 = 2*thr'.
         "parity": exactly the reference arithmetic, slot 8 = 2*abs_err.
  codec  "huff" (default): GPU Huffman containers in slots 6/7 (run tokens first when one symbol dominates);
         "bz2": the reference's bz2 level 9 of the first-differenced int32 / int8 streams (byte-identical
         slots 6/7 for identical codes).
Slot 0 (fpzip in the reference) is an 'LCL1' container: half-precision latent codes with a Huffman-coded high
byte plane on the 16-bit engines (the decoder consumes the rounded latent on both sides), raw fp32 on the fp32
engine, bz2 of the fp32 latent with codec="bz2".  compress_stream() is the bulk-job variant of compress().
"""
from __future__ import annotations

import bz2
import io
import pickle

import numpy as np

from . import _hostio
from . import _native as N
from ._engine import Engine


def pickle_parts(slots) -> list:
    """The pickle stream of the 9-slot list (compress.py:168) as a list of buffers whose concatenation
    pickle.loads() turns back into the list.  A byte-string slot may be given as a list of bytes-like pieces
    (its concatenation is the slot): the multi-GPU path assembles the blob straight from the pieces it received.

    The stream is assembled by hand -- PROTO 4, EMPTY_LIST, MARK, one BINBYTES8 opcode per bytes object with the
    object itself as the operand, the small items as protocol-2 fragments, APPENDS, STOP -- so that the
    multi-megabyte payloads are copied exactly once (pickle.dumps grows its buffer while appending them and then
    copies the result again: 1.0 ms against 0.45 ms for the 7.7 MB benchmark blob)."""
    import struct
    parts = [b"\x80\x04", b"]", b"("]
    for item in slots:
        if isinstance(item, (bytes, bytearray, memoryview)):
            parts += [b"\x8e", struct.pack("<Q", len(item)), item]
        elif isinstance(item, list) and all(isinstance(p, (bytes, bytearray, memoryview)) for p in item):
            parts += [b"\x8e", struct.pack("<Q", sum(len(p) for p in item))] + item
        else:
            buf = io.BytesIO()
            pk = pickle.Pickler(buf, protocol=2)
            pk.fast = True                                          # no memo opcodes: fragments must not share keys
            pk.dump(item)
            parts.append(buf.getvalue()[2:-1])                      # strip the PROTO header and the STOP opcode
    parts += [b"e", b"."]
    return parts


def pickle_frame(slots) -> tuple:
    """(bytes before slot 6, bytes after slot 7) of the blob's pickle stream, given the slot list with slots 6 and 7
    still to come: Engine.pack_streams(frame=...) writes the two entropy containers between them, straight into the
    finished blob."""
    parts = pickle_parts(list(slots[:6]) + [b"", b""] + list(slots[8:]))
    # the two empty placeholders are each [opcode, length, payload]; cut them out
    i = next(k for k in range(len(parts) - 5) if parts[k] == b"\x8e" and parts[k + 2] == b"" and parts[k + 3] == b"\x8e"
             and parts[k + 5] == b"")
    return b"".join(bytes(p) for p in parts[:i]), b"".join(bytes(p) for p in parts[i + 6:])


def dumps_slots(slots) -> bytes:
    """pickle.dumps(slots) for the 9-slot list (compress.py:168), written with a single copy of the payloads;
    pickle.loads() of the result is the same list, which is all decompress.py:54 relies on."""
    return b"".join(pickle_parts(slots))


def _guard(max_abs: float, abs_err: float) -> float:
    """Width taken off the threshold in strict mode: covers the float32 roundings between the
    exact residual and the decoded value (final add, residual, de-quantisation)."""
    return 2.0 * float(np.spacing(np.float32(max_abs + abs_err)))


class _StageTimer:
    """Wall-clock per stage with a device sync at every mark (diagnostics only)."""

    def __init__(self, eng):
        import time
        self.eng, self.time, self.ms = eng, time, {}
        eng.torch.cuda.synchronize()
        self.t = time.perf_counter()

    def mark(self, name):
        self.eng.torch.cuda.synchronize()
        now = self.time.perf_counter()
        self.ms[name] = self.ms.get(name, 0.0) + (now - self.t) * 1e3
        self.t = now
        return True


def compress_device(eng: Engine, x, err_threshold, *, mode="strict", codec="huff", verify=None,
                    inject=None, latent_codec=None, batch=None, return_parts=False, timing=False,
                    device_out=False, guard_scale=1.0, retry=True, as_blob=False):
    """x: torch float32 CUDA tensor (T,H,W,2).  Returns (slots list, info dict).
    device_out: slots 0/6/7 come back as _engine.DevPieces whose payloads are still in device memory (the
    multi-GPU path sends them over NCCL from there).  guard_scale multiplies strict mode's guard band; retry=False
    reports a bound violation in info["violations"] instead of widening the band and quantising again
    (dist.compress_sharded makes that decision for all ranks together).  as_blob: return the finished blob (bytes,
    what compress() returns) in place of the slot list -- the entropy payloads are then copied from the device straight
    into it.

    = compress_finish(compress_begin(...)): begin enqueues everything up to and including the quantiser without
    waiting for it, finish needs the host (outlier count, verdict, Huffman tables, payloads).  A caller with several
    fields may begin the next one (on another engine) before it finishes this one."""
    st = compress_begin(eng, x, err_threshold, mode=mode, codec=codec, verify=verify, inject=inject,
                        latent_codec=latent_codec, batch=batch, timing=timing, guard_scale=guard_scale)
    return compress_finish(st, return_parts=return_parts, device_out=device_out, retry=retry, as_blob=as_blob)


def compress_begin(eng: Engine, x, err_threshold, *, mode="strict", codec="huff", verify=None, inject=None,
                   latent_codec=None, batch=None, timing=False, guard_scale=1.0):
    """First half of compress_device(): statistics (one host sync unless they are injected), encoder, latent
    rounding, decoder and the quantiser's launch, all enqueued on the current stream.  Returns the state for
    compress_finish()."""
    assert err_threshold != 0, "The absolute error can't be 0."                      # compress.py:34
    assert x.dim() == 4 and x.shape[3] == eng.model.arch.in_channels, \
        "Input should have %d channels." % eng.model.arch.in_channels                # compress.py:55
    T, H, W, _ = x.shape
    inject = inject or {}
    abs_err = float(err_threshold)
    tm = _StageTimer(eng) if timing else None
    if "mean" in inject and "max_abs" in inject:     # global statistics of a time-sharded field: nothing to reduce here
        mean, std, max_abs = np.float32(inject["mean"]), np.float32(inject["std"]), float(inject["max_abs"])
    else:
        mean, std, max_abs = eng.stats(x)                                            # compress.py:64-65
        if "mean" in inject:             # parity tests
            mean, std = np.float32(inject["mean"]), np.float32(inject["std"])
    tm and tm.mark("stats")
    if latent_codec is None:
        # "f16": latent quantised to half precision + Huffman (16-bit engines); compat codec: lossless host coder
        # for the latent too (fpzip's role, compress.py:144); fp32 engine: raw fp32
        latent_codec = "bz2" if codec == "bz2" else ("f16" if eng.precision != "fp32" else "raw")
    latent = eng.encode(x, mean, std, batch=batch)                                   # compress.py:66-86
    planes = staged = None
    if latent_codec == "f16":
        planes = eng.quantize_latent(latent)     # in place: the decoder below consumes the rounded latent
    else:
        staged = eng.stage_latent(latent)        # its D2H copy runs behind the decoder; the quantiser's sync covers it
    tm and tm.mark("encode")
    if "recon_std" in inject:
        recon = inject["recon_std"]
    else:
        recon = eng.decode(latent, T, H, W, batch=batch)                             # compress.py:99-111
    tm and tm.mark("decode")
    thr = abs_err
    if mode == "strict":
        g = _guard(max_abs, abs_err) * float(guard_scale)
        assert abs_err > 2 * g, "abs_error is below the float32 resolution of the data"
        thr = abs_err - g
    elif mode != "parity":
        raise ValueError("mode must be 'strict' or 'parity'")
    if codec not in ("huff", "bz2"):
        raise ValueError("codec must be 'huff' or 'bz2'")
    # verify: None/True = bound check fused into the quantisation pass (strict mode's default);
    # "full" = run the decoder's reconstruct kernel and an independent fp64 comparison; False = none
    verify = (mode == "strict") if verify is None else verify
    pending = None
    if verify != "full":       # launched here, read in compress_finish()
        pending = eng.quantize_launch(x, recon, mean, std, thr, check_abs_err=abs_err if verify else None)
    return dict(eng=eng, x=x, abs_err=abs_err, mode=mode, codec=codec, verify=verify, latent_codec=latent_codec,
                mean=mean, std=std, thr=thr, latent=latent, planes=planes, staged=staged, recon=recon,
                pending=pending, tm=tm)


def compress_finish(st, *, return_parts=False, device_out=False, retry=True, as_blob=False):
    """Second half of compress_device(): latent packing (side stream), the quantiser's result (strict mode: verdict,
    wider guard band and another pass on a violation), entropy coding, the slot list / blob."""
    eng, x, abs_err, mode, codec, verify = st["eng"], st["x"], st["abs_err"], st["mode"], st["codec"], st["verify"]
    latent_codec, mean, std, thr, latent, recon, tm = (st["latent_codec"], st["mean"], st["std"], st["thr"],
                                                       st["latent"], st["recon"], st["tm"])
    T, H, W, _ = x.shape
    comp_data = None
    if latent_codec == "f16":
        comp_data = eng.pack_latent_f16(*st["planes"], device_out=device_out)        # side stream, behind the decoder
    info = {"mode": mode, "codec": codec}
    pending = st["pending"]
    while True:
        thr2 = thr * 2.0                                                             # compress.py:129
        if verify == "full":
            mask, q = eng.quantize(x, recon, mean, std, thr)                         # compress.py:114-135
            xhat = eng.reconstruct(recon, mask, q, mean, std, thr2)
            viol, max_err = eng.verify(x, xhat, abs_err)
        elif verify:
            if pending is None:
                pending = eng.quantize_launch(x, recon, mean, std, thr, check_abs_err=abs_err)
            mask, q, viol, max_err = eng.quantize_result(pending)
            pending = None
        else:
            mask, q = eng.quantize_result(pending)
            break
        info.update(violations=viol, max_err=max_err)
        if viol == 0 or mode != "strict" or not retry:
            break
        thr = abs_err - 2.0 * (abs_err - thr)    # widen the guard band and redo the cheap stages
        assert thr > 0.5 * abs_err, "cannot meet the bound in float32"
    tm and tm.mark("quantize+verify")
    n_out = int(q.numel())
    if comp_data is None:
        comp_data = eng.pack_latent(latent, latent_codec, st["staged"])              # compress.py:142-144
    tm and tm.mark("latent_pack")
    # compress.py:142 squeezes the latent's time axis, which is 1 after four reducing convolutions; the 3-conv
    # variant keeps its (2, 6, 6, F) bottleneck whole
    ls = tuple(eng.model.arch.latent_shape())
    lshape = (latent.shape[0],) + (ls[1:] if ls[0] == 1 else ls)
    blob = None
    if codec == "bz2":
        dq = eng.diff(q).cpu().numpy()                                               # compress.py:147-151
        dm = eng.diff(mask).cpu().numpy()                                            # compress.py:159-163
        comp_error = bz2.compress(dq.tobytes(), 9)                                   # compress.py:156
        comp_mask = bz2.compress(dm.tobytes(), 9)                                    # compress.py:166
    else:
        if as_blob and not device_out and n_out and mask.numel():
            head = [comp_data, lshape, (T, H, W, int(x.shape[3])), (n_out,), mean, std, None, None, thr2]
            blob, (ne_b, nm_b) = eng.pack_streams(mask, q, frame=pickle_frame(head))
            comp_error, comp_mask = range(ne_b), range(nm_b)           # only their lengths are reported below
        else:
            comp_error, comp_mask = eng.pack_streams(mask, q, device_out=device_out)
        tm and tm.mark("pack_streams")
    slots = [comp_data, lshape, (T, H, W, int(x.shape[3])), (n_out,), mean, std, comp_error, comp_mask, thr2]
    info.update(n_out=n_out, thr=thr, latent_bytes=len(comp_data), error_bytes=len(comp_error),
                mask_bytes=len(comp_mask))
    if tm:
        info["timing_ms"] = tm.ms
    if return_parts:
        info.update(latent=latent, recon_std=recon, mask=mask, q=q)
    if as_blob:
        return (blob if codec == "huff" and blob is not None else dumps_slots(slots)), info
    return slots, info


def compress(array, err_threshold, verbose=False, *, mode="strict", codec="huff", precision=None,
             model=None, device=None, verify=None):
    """Compression algorithm using a deep convolutional autoencoder (reference: compress.py:21).

    array: 4D numpy array (time, lat, lon, 2); err_threshold: absolute error.  Returns bytes."""
    torch = N.require_cuda()
    assert err_threshold != 0, "The absolute error can't be 0."
    if array.dtype != np.float32:                                                    # compress.py:37-38
        array = np.array(array, dtype=np.float32)
    eng = Engine.get(model, precision, device)
    # compress.py:55 for the default 2-channel model; the 1- and 5-channel models (compress_test.py:90-115) take
    # their own channel count
    nch = eng.model.arch.in_channels
    assert array.ndim == 4 and array.shape[3] == nch, \
        "Input should have 2 channels." if nch == 2 else "Input should have %d channels." % nch
    assert array.shape[0] >= 16 and array.shape[1] >= 48 and array.shape[2] >= 48, \
        "Chunks size cannot be larger than the data size."                           # dataLoader.py:447
    x = _hostio.h2d(eng, array)
    out, info = compress_device(eng, x, err_threshold, mode=mode, codec=codec, verify=verify, as_blob=True)   # compress.py:168
    if verbose:
        tot = info["latent_bytes"] + info["error_bytes"] + info["mask_bytes"]
        print("Latent space (%):", info["latent_bytes"] / tot)
        print("Error space (%):", info["error_bytes"] / tot)
        print("Error mask (%):", info["mask_bytes"] / tot)
        print("Compression factor:", array[:, :, :, 0].nbytes / len(out))            # compress.py:175
        print("Done.", {k: v for k, v in info.items() if not hasattr(v, "shape")})
    return out


def compress_stream(arrays, err_threshold, *, mode="strict", codec="huff", precision=None, model=None, device=None,
                    lanes=None):
    """Compress a sequence of fields, yielding one blob per field, in order (each identical to compress(field, ...)).

    Extension for bulk jobs (a month of 24 h slices).  Two things overlap the kernels of a field: the host->device
    copy of the next one, on a copy stream into a ring of input buffers (3.6 ms of a 12 ms call on the benchmark
    field), and the host round trips of its neighbour -- consecutive fields run in alternating lanes
    (lossycomp.lanes: an engine, a stream and a host thread each; default 2, LOSSYCOMP_LANES), so that while one
    field waits for its statistics, outlier count or Huffman tables the other one's convolutions keep the GPU busy.
    Pinned inputs overlap; pageable ones are staged like in compress()."""
    from collections import deque
    from .lanes import LaneSet
    torch = N.require_cuda()
    ls = LaneSet.get(model, precision, device, lanes)
    eng = ls.lanes[0].eng
    copy = torch.cuda.Stream(device=eng.dev)
    nbuf = len(ls) + 1                    # one field per lane in flight + the one being uploaded
    bufs = [None] * nbuf

    def upload(slot, array):
        """-> (device tensor, event to wait for or None, object to keep alive)"""
        assert err_threshold != 0, "The absolute error can't be 0."
        if array.dtype != np.float32:
            array = np.array(array, dtype=np.float32)
        assert array.ndim == 4 and array.shape[3] == eng.model.arch.in_channels, \
            "Input should have %d channels." % eng.model.arch.in_channels
        assert array.shape[0] >= 16 and array.shape[1] >= 48 and array.shape[2] >= 48, \
            "Chunks size cannot be larger than the data size."
        src = torch.from_numpy(np.ascontiguousarray(array))
        if not src.is_pinned():
            return _hostio.h2d(eng, array), None, None
        # the job that read this buffer nbuf fields ago has been collected (at most len(ls) jobs are in flight)
        if bufs[slot] is None or bufs[slot].shape != src.shape:
            bufs[slot] = torch.empty(src.shape, dtype=torch.float32, device=eng.dev)
        with torch.cuda.stream(copy):
            bufs[slot].copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy)
        return bufs[slot], ev, src                # the pinned source stays alive until the job has run

    def job(lane, x, ev, _keep):
        if ev is not None:
            lane.stream.wait_event(ev)
        return compress_device(lane.eng, x, err_threshold, mode=mode, codec=codec, as_blob=True)[0]

    inflight = deque()
    for k, array in enumerate(arrays):
        inflight.append(ls.submit(job, *upload(k % nbuf, array)))
        while len(inflight) > len(ls):
            yield inflight.popleft().result()
    while inflight:
        yield inflight.popleft().result()
