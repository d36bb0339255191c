"""Lanes: several fields in flight on one GPU.

compress_device() needs the host four times per field (mean/std, the outlier count and strict mode's verdict, the
code histograms for the Huffman tables, the finished payloads), and each of those round trips leaves the GPU idle
while Python runs: 1.4 of the 8.9 ms of a benchmark field.  A lane is an Engine of its own (weights, autoencoder
workspace, staging buffers), a CUDA stream and one host thread; consecutive fields go to alternating lanes, so the
kernels of one field fill the host gaps of the other.  Nothing is shared between lanes, every field is still
compressed by exactly the single-lane code, and the blobs are identical (tests/test_gpu_lanes.py).

Time-sharded fields (lossycomp.dist) use the engines, streams and turn-taking of a LaneSet but NOT its threads: one
thread begins field k+1 on the other engine and then finishes field k (dist.compress_sharded_stream), so every
collective has one issuer and one order on every rank.  Lane threads with a process group each (LaneSet.groups())
deadlocked over NCCL: a device-synchronising call made with the GIL held waits for the other thread's NCCL kernel,
which waits for the peer rank, whose other thread is kept from launching the matching call by the same situation.
"""
from __future__ import annotations

import contextlib
import os
import threading
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Dict, List, Optional

from . import _native as N
from ._engine import DEFAULT_PRECISION, Engine

DEFAULT_LANES = 2


class GpuTurn:
    """Serialises the autoencoder passes of the lanes of one device.  A pass holds the host lock only while its
    kernels are being enqueued (they are asynchronous); on the GPU it starts after the event the previous pass left
    behind, whichever lane that was."""

    def __init__(self):
        self._lock = threading.Lock()
        self._last = None

    @contextlib.contextmanager
    def section(self, eng):
        torch = eng.torch
        with self._lock:
            stream = torch.cuda.current_stream(eng.dev)
            if self._last is not None:
                stream.wait_event(self._last)
            try:
                yield
            finally:
                ev = torch.cuda.Event()
                ev.record(stream)
                self._last = ev


class Lane:
    def __init__(self, eng: Engine, index: int):
        torch = eng.torch
        self.eng, self.index = eng, index
        self.stream = torch.cuda.Stream(device=eng.dev)
        self.group = None                     # process group of this lane (LaneSet.groups())
        self._pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="lossycomp-lane%d" % index)

    def submit(self, fn, *args, **kwargs) -> Future:
        """Run fn(lane, *args, **kwargs) in the lane's thread with the lane's stream current."""
        torch = self.eng.torch

        def run():
            if torch.cuda.current_device() != self.eng.device:      # the current device is per thread
                torch.cuda.set_device(self.eng.device)
            with torch.cuda.stream(self.stream):
                return fn(self, *args, **kwargs)
        return self._pool.submit(run)


class LaneSet:
    """n lanes on one device; submit() deals jobs round-robin.  Lane 0 uses the default Engine of the device."""

    _cache: Dict[tuple, "LaneSet"] = {}

    @classmethod
    def get(cls, model_path: Optional[str] = None, precision: Optional[str] = None, device: Optional[int] = None,
            n: Optional[int] = None) -> "LaneSet":
        torch = N.require_cuda()
        precision = precision or os.environ.get("LOSSYCOMP_PRECISION", DEFAULT_PRECISION)
        device = torch.cuda.current_device() if device is None else device
        n = n_lanes(n)
        key = (model_path or "default", precision, device, n)
        if key not in cls._cache:
            cls._cache[key] = LaneSet([Engine.get(model_path, precision, device, lane=i) for i in range(n)])
        return cls._cache[key]

    def __init__(self, engines: List[Engine]):
        self.lanes = [Lane(e, i) for i, e in enumerate(engines)]
        if len(engines) > 1 and os.environ.get("LOSSYCOMP_LANE_TURNS", "1") != "0":
            turn = GpuTurn()
            for e in engines:
                e.turn = turn
        self._next = 0
        self._groups_of = None

    def __len__(self):
        return len(self.lanes)

    @property
    def engines(self):
        return [l.eng for l in self.lanes]

    def submit(self, fn, *args, **kwargs) -> Future:
        lane = self.lanes[self._next % len(self.lanes)]
        self._next += 1
        return lane.submit(fn, *args, **kwargs)

    def groups(self, group=None):
        """Give every lane its own process group over the ranks of `group` (collective: every rank must call it, in
        the same order).  The communicators are created and used once here, one after the other, so that no two
        threads ever initialise NCCL at the same time.  For experiments with threaded lanes over gloo; over NCCL see
        the module docstring -- the sharded codec does not use this."""
        import torch.distributed as dist
        key = id(group) if group is not None else None
        if self._groups_of is not None and self._groups_of[0] == key:
            return self
        ranks = dist.get_process_group_ranks(group) if group is not None else list(range(dist.get_world_size()))
        torch = self.lanes[0].eng.torch
        for lane in self.lanes:
            lane.group = dist.new_group(ranks=ranks, backend=dist.get_backend(group))
            dev = lane.eng.dev if dist.get_backend(lane.group) == "nccl" else "cpu"
            lane.submit(lambda ln, d=dev: dist.all_reduce(torch.zeros(1, device=d), group=ln.group)).result()
        self._groups_of = (key, group)
        return self

    def launches(self) -> int:
        return sum(l.eng.launches for l in self.lanes)


def n_lanes(n: Optional[int] = None) -> int:
    if n is None:
        n = int(os.environ.get("LOSSYCOMP_LANES", DEFAULT_LANES))
    return max(1, int(n))
