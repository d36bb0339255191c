"""Host-side logic of the multi-GPU path, run on CPU: time sharding, statistics combination, the
multi-part entropy container and the blob assembly; the collectives with 2 gloo ranks."""
import os
import pickle
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lossycomp import dist as D


def test_shard_time_whole_chunks():
    # 45 time chunks on 8 ranks -> 6,6,6,6,6,5,5,5; never split a 16-step chunk
    sh = D.shard_time(720, 8)
    assert [(b - a) // 16 for a, b in sh] == [6, 6, 6, 6, 6, 5, 5, 5]
    assert sh[0][0] == 0 and sh[-1][1] == 720 and all(a % 16 == 0 for a, _ in sh)
    assert [b - a for a, b in D.shard_time(720, 2)] == [368, 352]
    assert [b - a for a, b in D.shard_time(720, 4)] == [192, 176, 176, 176]
    # ragged tail goes to the last non-empty rank; empty ranks when there are fewer chunks than ranks
    assert D.shard_time(24, 2) == [(0, 24), (24, 24)]
    assert D.shard_time(50, 2) == [(0, 32), (32, 50)]
    assert D.shard_time(33, 4) == [(0, 16), (16, 33), (33, 33), (33, 33)]
    for T in (16, 17, 31, 32, 100, 721):
        for w in (1, 2, 3, 4, 8):
            sh = D.shard_time(T, w)
            assert sh[0][0] == 0 and max(b for _, b in sh) == T
            for a, b in sh:
                assert b == a or (a % 16 == 0 and b - a >= 16)
    with pytest.raises(AssertionError):
        D.shard_time(15, 2)


def test_combine_stats_matches_whole_array():
    rng = np.random.default_rng(0)
    x = (265 + 16 * rng.standard_normal(100000)).astype(np.float32)
    parts = []
    for a in np.array_split(x, 3):
        a64 = a.astype(np.float64)
        parts.append((a64.sum(), (a64 * a64).sum(), float(np.abs(a).max()), float(a.size)))
    mean, std, mx = D.combine_stats(parts)
    assert mean.dtype == np.float32 and abs(float(mean) - float(x.mean(dtype=np.float64))) < 1e-4
    assert abs(float(std) - float(x.std(dtype=np.float64))) < 1e-4 and mx == float(np.abs(x).max())


def test_multipart_roundtrip_and_blob_assembly():
    parts = [b"LCH1" + bytes(range(40)), b"LCH1", b"LCH1" + b"\x00" * 1000]
    assert D.unpack_multipart(D.pack_multipart(parts)) == parts
    assert D.pack_multipart(parts[:1]) == parts[0] and D.unpack_multipart(parts[0]) == [parts[0]]
    mean, std = np.float32(265.0), np.float32(16.0)
    a = [b"A", (20, 1, 3, 3), (32, 96, 96, 2), (100,), mean, std, b"LCH1x", b"LCH1y", 0.2]
    b = [b"B", (10, 1, 3, 3), (16, 96, 96, 2), (50,), mean, std, b"LCH1z", b"LCH1w", 0.2]
    s = D.assemble_blob([a, b], lambda xs: b"".join(xs))
    assert s[0] == b"AB" and s[1] == (30, 1, 3, 3) and s[2] == (48, 96, 96, 2) and s[3] == (150,)
    assert D.unpack_multipart(s[6]) == [b"LCH1x", b"LCH1z"] and s[8] == 0.2
    assert D.unframe_slots(b"".join(D.frame_slots(a))) == a
    b[8] = 0.3
    with pytest.raises(AssertionError):
        D.assemble_blob([a, b], lambda xs: b"".join(xs))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        x = (265 + 16 * rng.standard_normal(4000)).astype(np.float64)
        lo, hi = (0, 1500) if rank == 0 else (1500, 4000)
        a = x[lo:hi]
        tot = D.all_reduce_stats((a.sum(), (a * a).sum(), float(np.abs(a).max()), float(a.size)))
        payload = [bytes([rank]) * 10, b"", bytes([rank]) * (1000 * rank)]     # a list is sent concatenated
        got = D.gather_bytes(payload, dst=0)
        empty = D.gather_bytes(b"" if rank == 1 else b"xy", dst=0)
        q.put((rank, tot, got, empty))
    finally:
        dist.destroy_process_group()


def test_collectives_two_gloo_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(7)
    x = (265 + 16 * rng.standard_normal(4000)).astype(np.float64)
    for rank, tot, got, empty in res:
        assert abs(tot[0] - x.sum()) < 1e-6 and abs(tot[1] - (x * x).sum()) < 1e-3
        assert tot[2] == float(np.abs(x).max()) and tot[3] == 4000.0
        if rank == 0:
            assert got == [bytes([0]) * 10, bytes([1]) * 1010] and empty == [b"xy", b""]
        else:
            assert got is None and empty is None
