"""Multi-GPU (one process per GPU, torch.distributed) compress / decompress of a time-sharded field.

The reference is single-process; what shards is the chunk list: chunks are independent
(dataLoader.py:437-481, no halo), and the chunk index is time-major, so cutting the time axis in
whole 16-step chunks gives every rank a contiguous chunk range whose local chunk grid (including
the re-anchored tail window, which is anchored at the END of the axis both locally and globally)
is exactly its part of the global one.  To produce the blob a single compress() call would have
produced, only three things cross ranks:

  1. the statistics of channel 0 (compress.py:64-65 are global): all-reduce of {sum, sum^2, n} (SUM)
     and max|x| (MAX) -- 32 bytes;
  2. the outlier counts (slot 3) -- 8 bytes per rank;
  3. the blob parts (latent rows, entropy-coded code / mask streams) gathered to rank 0.

The entropy containers of slots 6/7 become multi-part ('LCM1': one self-contained 'LCH1' stream per
rank, first differences restarting at every part), so no seam values and no shared Huffman table are
needed; the decoded mask / codes / reconstruction are bit-identical to the single-GPU ones.
NCCL carries the collectives on GPUs; the same code runs on gloo with CPU tensors for the host-logic
tests.
"""
from __future__ import annotations

import math
import pickle
import struct
from typing import List, Optional, Sequence, Tuple

import numpy as np

MULTI_MAGIC = b"LCM1"
CHUNK_T = 16


# ------------------------------------------------------------------------------------------
# pure host logic (unit-tested on CPU)
# ------------------------------------------------------------------------------------------
def shard_time(T: int, world: int) -> List[Tuple[int, int]]:
    """Rows [begin, end) of every rank: whole 16-step chunks, as even as possible (45 chunks on 8
    ranks -> 6,6,6,6,6,5,5,5).  A ragged tail (T % 16 != 0) goes to the last non-empty rank, whose
    shard is then long enough (>= 16 rows) to hold the re-anchored tail window."""
    assert T >= CHUNK_T, "Chunks size cannot be larger than the data size."
    n_full = T // CHUNK_T
    tail = T - n_full * CHUNK_T
    base, extra = divmod(n_full, world)
    out, t = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((t, t + n * CHUNK_T))
        t += n * CHUNK_T
    if tail:
        last = max(r for r in range(world) if out[r][1] > out[r][0])
        out[last] = (out[last][0], out[last][1] + tail)
        for r in range(last + 1, world):
            out[r] = (T, T)
    return out


def combine_stats(parts: Sequence[Sequence[float]]):
    """[(sum, sum^2, max|x|, n)] per rank -> (mean float32, std float32, max|x|), the same formula the
    single-GPU path applies to its own sums (Engine.stats)."""
    s = sum(p[0] for p in parts)
    s2 = sum(p[1] for p in parts)
    mx = max(p[2] for p in parts)
    n = sum(p[3] for p in parts)
    mean64 = s / n
    var = max(s2 / n - mean64 * mean64, 0.0)
    return np.float32(mean64), np.float32(math.sqrt(var)), mx


def pack_multipart(parts: Sequence[bytes]) -> bytes:
    """Several self-contained entropy streams -> one slot payload."""
    if len(parts) == 1:
        return parts[0]
    return b"".join([MULTI_MAGIC, struct.pack("<I", len(parts)),
                     struct.pack(f"<{len(parts)}Q", *(len(p) for p in parts))] + list(parts))


def unpack_multipart(blob: bytes) -> List[bytes]:
    if blob[:4] != MULTI_MAGIC:
        return [blob]
    (n,) = struct.unpack("<I", blob[4:8])
    lens = struct.unpack(f"<{n}Q", blob[8:8 + 8 * n])
    out, p = [], 8 + 8 * n
    for ln in lens:
        out.append(blob[p:p + ln])
        p += ln
    return out


def assemble_blob(rank_slots: Sequence[list], latent_join) -> list:
    """Per-rank slot lists (each describing its own time shard) -> the slot list of the whole field.
    `latent_join(list of slot-0 payloads) -> payload` concatenates the latent containers."""
    first = rank_slots[0]
    n_chunks = sum(s[1][0] for s in rank_slots)
    T = sum(s[2][0] for s in rank_slots)
    n_out = sum(s[3][0] for s in rank_slots)
    for s in rank_slots:
        assert s[4] == first[4] and s[5] == first[5] and s[8] == first[8], "ranks disagree on mean/std/threshold"
        assert tuple(s[2][1:]) == tuple(first[2][1:]) and tuple(s[1][1:]) == tuple(first[1][1:])
    return [latent_join([s[0] for s in rank_slots]), (n_chunks,) + tuple(first[1][1:]),
            (T,) + tuple(first[2][1:]), (n_out,), first[4], first[5],
            pack_multipart([s[6] for s in rank_slots]), pack_multipart([s[7] for s in rank_slots]), first[8]]


def frame_slots(slots) -> list:
    """Slot list -> [small pickled header, payload 0, payload 6, payload 7]: the three large byte strings
    travel unpickled, so framing costs no extra copy of the payloads."""
    big = (0, 6, 7)
    head = list(slots)
    for i in big:
        head[i] = len(slots[i])
    h = pickle.dumps(head)
    return [struct.pack("<I", len(h)), h] + [slots[i] for i in big]


def unframe_slots(buf) -> list:
    """Inverse of frame_slots on a bytes-like object."""
    mv = memoryview(buf)
    (hl,) = struct.unpack("<I", mv[:4])
    head = pickle.loads(mv[4:4 + hl])
    p = 4 + hl
    for i in (0, 6, 7):
        n = head[i]
        head[i] = bytes(mv[p:p + n])
        p += n
    return head


# ------------------------------------------------------------------------------------------
# collectives (NCCL on GPUs, gloo on CPU tensors)
# ------------------------------------------------------------------------------------------
def all_reduce_stats(local, group=None, device=None):
    """local = (sum, sum^2, max|x|, n) -> the same tuple over all ranks."""
    import torch
    import torch.distributed as dist
    a = torch.tensor([local[0], local[1], local[3]], dtype=torch.float64, device=device)
    m = torch.tensor([local[2]], dtype=torch.float64, device=device)
    dist.all_reduce(a, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    a = a.cpu().tolist()
    return (a[0], a[1], float(m.item()), a[2])


_pinned = {}


def _stage(parts, device):
    """Concatenate byte strings straight into a (reused) pinned host tensor and start the copy to `device`."""
    import torch
    total = sum(len(p) for p in parts)
    use_pin = device is not None and torch.device(device).type == "cuda"
    key = str(device)
    host = _pinned.get(key)
    if host is None or host.numel() < total:
        host = torch.empty(max(int(total * 1.25), 1 << 16), dtype=torch.uint8, pin_memory=use_pin)
        _pinned[key] = host
    hv = host.numpy()
    off = 0
    for p in parts:
        n = len(p)
        if n:
            hv[off:off + n] = np.frombuffer(p, dtype=np.uint8)
        off += n
    return host, total


def gather_bytes(payload, dst: int = 0, group=None, device=None, to_host: bool = True):
    """Variable-length byte strings of all ranks -> list on rank `dst` (None elsewhere).  `payload` is a
    bytes object or a list of bytes objects (sent as their concatenation).  Lengths travel in an
    all-gather, payloads in one padded gather (NCCL over NVLink on GPUs).  With to_host=False rank `dst`
    keeps the parts as uint8 tensors on its device (no PCIe trip)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    parts = payload if isinstance(payload, (list, tuple)) else [payload]
    host, total = _stage(parts, device)
    n = torch.tensor([total], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(v) for v in torch.cat(sizes).cpu().tolist()]
    cap = max(max(sizes), 1)
    buf = torch.empty(cap, dtype=torch.uint8, device=device)
    if total:
        buf[:total].copy_(host[:total], non_blocking=True)
    if rank == dst:
        recv = torch.empty((world, cap), dtype=torch.uint8, device=device)
        dist.gather(buf, list(recv.unbind(0)), dst=dst, group=group)
        if not to_host:
            return [recv[i, :sizes[i]] for i in range(world)]
        hostr = recv.cpu()
        return [hostr[i, :sizes[i]].numpy().tobytes() for i in range(world)]
    dist.gather(buf, None, dst=dst, group=group)
    return None


# ------------------------------------------------------------------------------------------
# sharded codec
# ------------------------------------------------------------------------------------------
def compress_sharded(eng, x_local, err_threshold, *, mode="strict", codec="huff", group=None, dst=0):
    """Every rank passes its own rows (a time shard made with shard_time) as a CUDA tensor
    (T_local, H, W, 2); rank `dst` gets the slot list of the WHOLE field (others None), plus an info
    dict on every rank."""
    import torch.distributed as dist
    from .compress import compress_device
    dev = eng.dev
    T_local = int(x_local.shape[0])
    local = eng.stats_raw(x_local) if T_local else (0.0, 0.0, 0.0, 0.0)
    mean, std, max_abs = combine_stats([all_reduce_stats(local, group, dev)])
    if T_local:
        slots, info = compress_device(eng, x_local, err_threshold, mode=mode, codec=codec,
                                      inject={"mean": mean, "std": std, "max_abs": max_abs})
        mine = frame_slots(slots)
    else:
        slots, info, mine = None, {"n_out": 0}, b""
    parts = gather_bytes(mine, dst, group, dev)
    if dist.get_rank(group) != dst:
        return None, info
    rank_slots = [unframe_slots(p) for p in parts if p]
    return assemble_blob(rank_slots, eng.join_latents), info


def decompress_sharded(eng, slots_or_none, *, group=None, src=0):
    """Rank `src` holds the slot list; every rank decodes the autoencoder for its own time shard only
    and returns (rows_begin, rows_end, tensor (T_local, H, W)) of the reconstruction."""
    import torch
    import torch.distributed as dist
    from .decompress import decompress_device
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    obj = [slots_or_none]
    dist.broadcast_object_list(obj, src=src, group=group)
    slots = obj[0]
    T, H, W = (int(v) for v in slots[2][:3])
    t0, t1 = shard_time(T, world)[rank]
    if t1 <= t0:
        return t0, t1, torch.empty((0, H, W), dtype=torch.float32, device=eng.dev)
    out = decompress_device(eng, slots, rows=(t0, t1))
    return t0, t1, out
