"""Time-sharded compress / decompress with REAL torch.distributed ranks on one GPU box.

Two processes share cuda:0 for the kernels and talk through gloo (NCCL cannot put two ranks on one device; the
collectives' operands are staged through host memory for gloo by lossycomp.dist, everything else -- per-block
statistics, the strict-mode verdict, device-resident frames, the multi-part containers, the scatter of parts on
the decode side -- is exactly the code that runs over NCCL on 2/4/8 GPUs in bench.py --gpus N).
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, shape, abs_err, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lossycomp import _native as N, dist as D
        from lossycomp._engine import Engine
        from lossycomp.compress import compress_device
        from lossycomp.decompress import decompress_device
        torch.cuda.set_device(0)
        eng = Engine.get(None, "fp16", 0)
        T, H, W = shape
        full = torch.empty((T, H, W, 2), dtype=torch.float32, device=eng.dev)
        N.check(eng.lib.lc_synth_field(full.data_ptr(), T, H, W, 0, 0, 77, None), "synth")
        t0, t1 = D.shard_time(T, world)[rank]
        dst = world - 1
        slots, info = D.compress_sharded(eng, full[t0:t1].contiguous(), abs_err, t0=t0, T=T, dst=dst)
        # the same call without telling the ranks where they are (one more tiny all-gather), asynchronously
        pend, _ = D.compress_sharded(eng, full[t0:t1].contiguous(), abs_err, dst=dst, wait=False)
        slots2 = pend.result() if pend is not None else None
        # host-facing entry points: numpy rows in, pickled blob out on dst; the streamed form gives the same blobs
        import pickle
        xl = full[t0:t1].cpu().numpy()
        blob = D.compress_sharded_host(xl, abs_err, t0=t0, T=T, precision="fp16", dst=dst)
        blobs = list(D.compress_sharded_stream([xl, xl, xl], abs_err, t0=t0, T=T, precision="fp16", dst=dst))
        # two fields in flight (one process group per lane; gloo here -- see compress_sharded_stream on NCCL)
        blobs += list(D.compress_sharded_stream([xl, xl, xl], abs_err, t0=t0, T=T, precision="fp16", dst=dst, lanes=2))
        a3, b3, rows3 = D.decompress_sharded_host(blob, precision="fp16", src=dst)
        a, b, rows = D.decompress_sharded(eng, slots, src=dst)                       # aligned: parts scattered
        assert (a, b) == (t0, t1)
        err = float((full[a:b, :, :, 0].double() - rows.double()).abs().max()) if b > a else 0.0
        res = {"rank": rank, "err": err}
        if rank == dst:
            ref, rinfo = compress_device(eng, full, abs_err, return_parts=True)
            whole = decompress_device(eng, slots)
            res.update(
                stats_equal=bool(slots[4] == ref[4] and slots[5] == ref[5] and slots[8] == ref[8]),
                shapes_equal=bool(slots[1] == ref[1] and slots[2] == ref[2] and slots[3] == ref[3]),
                mask_equal=bool(torch.equal(eng.unpack_mask(slots[7]), rinfo["mask"])),
                codes_equal=bool(torch.equal(eng.unpack_codes(slots[6]), rinfo["q"])),
                xhat_equal=bool(torch.equal(whole, decompress_device(eng, ref))),
                rows_equal=bool(torch.equal(whole[a:b], rows)),
                same_without_hints=bool(all(bytes(slots[i]) == bytes(slots2[i]) for i in (0, 6, 7))),
                n_parts=len(D.unpack_multipart(slots[7])),
                host_blob_equal=bool(pickle.loads(blob) == [bytes(v) if isinstance(v, (bytes, memoryview)) else v
                                                            for v in slots]),
                stream_equal=bool(blobs == [blob] * 6),
                types=[type(s).__name__ for s in slots])
            unaligned = ref                      # a single-part blob on two ranks: the broadcast + slice path
        else:
            assert slots is None and slots2 is None and blob is None and blobs == [None] * 6
            unaligned = None
        res["host_rows_equal"] = bool((a3, b3) == (a, b) and np.array_equal(rows3[..., 0], rows.cpu().numpy()))
        a2, b2, rows2 = D.decompress_sharded(eng, unaligned, src=dst)
        res["unaligned_equal"] = bool(torch.equal(rows2, rows)) if b > a else True
        q.put(res)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape,world", [((48, 96, 144), 2), ((50, 100, 97), 2), ((24, 96, 96), 2)])
def test_sharded_ranks_reproduce_single_call(shape, world):
    """compress_sharded on `world` ranks == compress_device on one: statistics, mask, codes and reconstruction bit
    for bit, slots typed like the reference's; ragged time tails and an empty rank (24 rows on 2 ranks) included."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, shape, 0.1, q)) for r in range(world)]
    for p in procs:
        p.start()
    import queue as _queue
    res = []
    for _ in range(3000):                      # up to 5 minutes, but give up as soon as a rank has died
        try:
            res.append(q.get(timeout=0.1))
        except _queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
        if len(res) == len(procs):
            break
    for p in procs:
        p.join(timeout=120)
        if p.is_alive():
            p.kill()
        assert p.exitcode == 0
    assert len(res) == len(procs)
    res.sort(key=lambda r: r["rank"])
    assert all(r["err"] <= 0.1 and r["unaligned_equal"] and r["host_rows_equal"] for r in res)
    d = res[-1]
    assert d["stats_equal"] and d["shapes_equal"] and d["mask_equal"] and d["codes_equal"] and d["xhat_equal"]
    assert d["rows_equal"] and d["same_without_hints"] and d["host_blob_equal"] and d["stream_equal"]
    assert d["types"] == ["bytes", "tuple", "tuple", "tuple", "float32", "float32", "bytes", "bytes", "float"]
    n_nonempty = sum(1 for a, b in __import__("lossycomp.dist", fromlist=["x"]).shard_time(shape[0], world) if b > a)
    assert d["n_parts"] == n_nonempty
