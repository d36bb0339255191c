"""Host <-> device transfers of whole fields for the numpy-facing compress() / decompress().

A 100 MB result copied into a freshly allocated numpy array costs ~20 ms on the benchmark box, almost
all of it first-touch page faults taken by one thread (tools/d2h_probe.py: 1.8 ms for the PCIe copy into
pinned memory, 4.7 ms into a warm pageable array).  Both directions therefore go through a reused
pinned staging buffer in slices, and the pageable side of every slice is touched by a small thread pool
(numpy copies release the GIL) while the next slice is on the bus.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_SLICES = 8
_pool = None


def _n_workers():
    """Copy threads of this process: half the cores, shared between the ranks of the node (eight ranks with
    eight threads each on 32 cores made the 8-GPU decompress path 2.5x slower per rank than a single rank)."""
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 2
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
    return max(1, min(_SLICES, cores // (2 * max(local_world, 1))))


def _threads():
    global _pool
    if _pool is None:
        _pool = ThreadPoolExecutor(max_workers=_n_workers(), thread_name_prefix="lossycomp-io")
    return _pool


_big = None


def gather_pool():
    """Copy threads of the ONE rank per node that assembles (or distributes) the blob of a time-sharded field: it moves
    every rank's payload through host memory, so it gets up to half the node's cores instead of its 1/N share."""
    global _big
    if _big is None:
        try:
            cores = len(os.sched_getaffinity(0))
        except Exception:
            cores = os.cpu_count() or 2
        _big = ThreadPoolExecutor(max_workers=max(2, min(16, cores // 2)), thread_name_prefix="lossycomp-gather")
    return _big


def _bounds(n, parts):
    step = -(-n // parts)
    step = -(-step // 1024) * 1024
    return [(a, min(a + step, n)) for a in range(0, n, step)]


def prefault(shape):
    """Allocate the result array and start touching its pages in the thread pool (one write per 4 KB page).  Called
    before the device work of decompress() so that the first-touch page faults of the 100 MB result are taken while
    the GPU decodes; pass the returned (array, futures) to d2h()."""
    host = np.empty(shape, dtype=np.float32)
    flat = host.reshape(-1)
    jobs = [_threads().submit(_touch, flat[a:b]) for a, b in _bounds(flat.size, _SLICES)]
    return host, jobs


def _touch(part):
    part[::1024] = 0.0


def d2h(eng, t_dev, shape, out=None, wait=()):
    """Device float32 tensor -> numpy float32 array of `shape` (`out` if given, else freshly allocated)."""
    torch = eng.torch
    n = t_dev.numel()
    src = t_dev.reshape(-1)
    for j in wait:
        j.result()
    if out is not None:
        assert out.dtype == np.float32 and out.size == n and out.flags.c_contiguous, \
            "out must be a C-contiguous float32 array of the field's size"
        host = out
    else:
        host = np.empty(shape, dtype=np.float32)
    flat = host.reshape(-1)
    pin = eng._pinned("d2h", 4 * n)[:4 * n].view(torch.float32)
    pin_np = pin.numpy()
    jobs = []
    for a, b in _bounds(n, _SLICES):
        pin[a:b].copy_(src[a:b], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(eng.dev))

        def work(a=a, b=b, ev=ev):
            ev.synchronize()
            np.copyto(flat[a:b], pin_np[a:b])
        jobs.append(_threads().submit(work))
    for j in jobs:
        j.result()
    return host.reshape(shape)


def h2d(eng, array):
    """numpy float32 array -> device tensor.  Pinned inputs are copied directly; pageable ones are staged."""
    torch = eng.torch
    array = np.ascontiguousarray(array)
    src = torch.from_numpy(array)
    if src.is_pinned():
        return src.to(eng.dev, non_blocking=False)
    n = array.size
    dst = torch.empty(array.shape, dtype=torch.float32, device=eng.dev)
    dflat = dst.reshape(-1)
    flat = array.reshape(-1)
    pin = eng._pinned("h2d", 4 * n)[:4 * n].view(torch.float32)
    pin_np = pin.numpy()
    bounds = _bounds(n, _SLICES)

    def stage(a, b):
        np.copyto(pin_np[a:b], flat[a:b])
    jobs = [_threads().submit(stage, a, b) for a, b in bounds]
    for (a, b), j in zip(bounds, jobs):
        j.result()
        dflat[a:b].copy_(pin[a:b], non_blocking=True)
    torch.cuda.current_stream(eng.dev).synchronize()      # the staging buffer is reused by the next call
    return dst
