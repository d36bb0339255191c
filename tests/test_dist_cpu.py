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
    s = D.assemble_blob([a, b])
    assert D.unpack_multipart(s[0]) == [b"A", b"B"] and s[1] == (30, 1, 3, 3)
    assert s[2] == (48, 96, 96, 2) and s[3] == (150,)
    assert D.unpack_multipart(s[6]) == [b"LCH1x", b"LCH1z"] and s[8] == 0.2
    for bad in (s[6][:-1], s[6][:12], b"LCM1" + b"\xff" * 12):
        with pytest.raises(ValueError):
            D.unpack_multipart(bad)
    assert D.unframe_slots(b"".join(D.frame_slots(a))) == a
    b[8] = 0.3
    with pytest.raises(AssertionError):
        D.assemble_blob([a, b])


def test_device_frames_roundtrip():
    """frame_layout / parse_frame: slots whose big parts are lists of host and "device" pieces (CPU tensors stand in
    for CUDA tensors here) survive the trip as the same bytes, piece by piece."""
    from lossycomp._engine import DevPieces
    mean, std = np.float32(265.0), np.float32(16.0)
    t = lambda b: torch.frombuffer(bytearray(b), dtype=torch.uint8)
    slots = [DevPieces([b"LCL1hdr", t(b"lowplane"), b"LCH1head", t(b"hibits")]), (20, 3, 3, 23), (32, 96, 96, 2), (7,),
             mean, std, DevPieces([b"LCH1codes", t(b"\x01\x02\x03"), b"esc!"]), b"LCH1mask-on-host", 0.2]
    head, host, devp = D.frame_layout(slots)
    frame = b"".join([len(head).to_bytes(4, "little"), head] + [bytes(h) for h in host] +
                     [bytes(d.numpy()) for d in devp])
    back = D.parse_frame(frame)
    assert b"".join(bytes(p) for p in back[0]) == b"LCL1hdrlowplaneLCH1headhibits"
    assert b"".join(bytes(p) for p in back[6]) == b"LCH1codes\x01\x02\x03esc!"
    assert b"".join(bytes(p) for p in back[7]) == b"LCH1mask-on-host"
    assert back[1:6] == slots[1:6] and back[8] == 0.2
    assert len(slots[0]) == 29 and slots[6].tobytes() == b"LCH1codes\x01\x02\x03esc!"


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
        # per-block partials: rank 0 owns blocks 0..2 (500 values each), rank 1 blocks 3..7
        blk = torch.tensor([[b.sum(), (b * b).sum(), float(np.abs(b).max()), float(b.size)]
                            for b in np.split(a, (hi - lo) // 500)], dtype=torch.float64)
        tot = D.all_reduce_blocks(blk, lo // 500, 8)
        payload = [bytes([rank]) * 10, b"", bytes([rank]) * (1000 * rank)]     # a list is sent concatenated
        got = D.gather_bytes(payload, dst=0)
        empty = D.gather_bytes(b"" if rank == 1 else b"xy", dst=0)
        # the gather the sharded compressor uses: tensors in, one receive buffer + (offset, length, meta) out
        fr = torch.full((100 + 900 * rank,), rank + 1, dtype=torch.uint8)
        g2 = D.gather_tensors(fr, 7 + rank, dst=1)
        if rank == 1:
            recv, table = g2
            assert [(n, m) for _, n, m in table] == [(100, 7), (1000, 8)]
            assert all(bool((recv[o:o + n] == r + 1).all()) for r, (o, n, _) in enumerate(table))
        else:
            assert g2 is None
        assert D.all_reduce_max(5 * rank) == 5
        # strict mode's collective verdict: one rank's violation makes every rank retry, nothing is sent
        assert D.gather_tensors(fr, 0, dst=1, veto=3 * rank) == "veto"
        g3 = D.gather_tensors(fr, 0, dst=1, veto=0)
        assert (g3 is None) == (rank == 0)
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
    want = [(b.sum(), (b * b).sum(), float(np.abs(b).max()), 500.0) for b in np.split(x, 8)]
    for rank, tot, got, empty in res:
        assert tot == want                  # exact: disjoint entries, x + 0 in any order
        if rank == 0:
            assert got == [bytes([0]) * 10, bytes([1]) * 1010] and empty == [b"xy", b""]
        else:
            assert got is None and empty is None
