"""BASELINE.json configs[2] and configs[4]: one synthetic ERA5 month, 720 h x L levels x 720 x 1440 fp32 (+ land-sea
channel), compressed level by level with the time axis sharded over the ranks (dist.compress_sharded, blobs gathered on
the last rank), then decompressed from those blobs (dist.decompress_sharded) with the exhaustive per-element bound
check against the regenerated input.

    torchrun --nproc-per-node N tools/month_run.py [--levels 37] [--hours 720] [--abs-err 0.1]

Prints one JSON line (rank 0) with compress / decompress GB/s (device time, max over ranks, all levels), compression
factor, violations, and the per-level spread.  Levels are generated on the device (lc_synth_field) one at a time, so a
rank holds one level's shard (0.8 GB at 8 ranks) plus the batch workspace.
"""
import argparse, json, os, pickle, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from lossycomp import _native as N, dist as D
from lossycomp._engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--levels", type=int, default=37)
ap.add_argument("--hours", type=int, default=720)
ap.add_argument("--abs-err", type=float, default=0.1)
ap.add_argument("--precision", default="fp16")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = Engine.get(None, a.precision, local)
T, H, W = a.hours, 720, 1440
t0, t1 = D.shard_time(T, world)[rank]
dst = world - 1
x = torch.empty((t1 - t0, H, W, 2), dtype=torch.float32, device=eng.dev)


def synth(level):
    N.check(eng.lib.lc_synth_field(x.data_ptr(), t1 - t0, H, W, t0, level, 1234, None), "synth")


def sync():
    dist.barrier()
    torch.cuda.synchronize()


def gmax(v):
    t = torch.tensor([v], dtype=torch.float64, device=eng.dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# warm-up (allocations, NCCL channels)
synth(0)
D.compress_sharded(eng, x, a.abs_err, t0=t0, T=T, dst=dst)
# compress loop: level k's blob reaches the host (copy stream + host threads on rank dst) behind level k+1's kernels
blobs, pend = [], []
sync()
wall0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for lev in range(a.levels):
    synth(lev)                                     # device-side generation of the level stands in for its H2D
    out, info = D.compress_sharded(eng, x, a.abs_err, t0=t0, T=T, dst=dst, wait=False, pickled=True)
    pend.append(out)
    if len(pend) > 1 and rank == dst:
        blobs.append(pend[-2].result())
if rank == dst:
    blobs.append(pend[-1].result())
e1.record()
sync()
c_ms = gmax(e0.elapsed_time(e1))
wall_c = time.perf_counter() - wall0
# decompress-only pass over the stored blobs + exhaustive bound check against the regenerated input
d_ms, viol, worst = 0.0, 0, 0.0
wall0 = time.perf_counter()
for lev in range(a.levels):
    slots = pickle.loads(blobs[lev]) if rank == dst else None
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    b0, b1, rows = D.decompress_sharded(eng, slots, src=dst)
    e1.record()
    sync()
    d_ms += gmax(e0.elapsed_time(e1))
    synth(lev)
    v, mx = eng.verify(x, rows, a.abs_err)
    vt = torch.tensor([float(v)], dtype=torch.float64, device=eng.dev)
    dist.all_reduce(vt)
    viol += int(vt.item())
    worst = max(worst, gmax(mx))
wall_d = time.perf_counter() - wall0
raw = T * H * W * 4 * a.levels
tot = [sum(len(b) for b in blobs)]
dist.broadcast_object_list(tot, src=dst)
if rank == 0:
    print(json.dumps({
        "workload": f"synthetic ERA5 month {T} h x {a.levels} levels x {H} x {W} fp32 + lsm ({raw / 1e9:.1f} GB of channel 0), "
                    f"abs_err {a.abs_err}, time-sharded over {world} ranks ({(t1 - t0)} h on rank 0), level by level",
        "n_gpus": world, "levels": a.levels, "chunks_per_level": eng.n_chunks(T, H, W),
        "compress_GBps": raw / (c_ms * 1e-3) / 1e9, "compress_ms_per_level": c_ms / a.levels,
        "decompress_GBps": raw / (d_ms * 1e-3) / 1e9, "decompress_ms_per_level": d_ms / a.levels,
        "compression_factor": raw / tot[0], "blob_bytes": tot[0], "violations": viol, "max_abs_err": worst,
        "wall_s": {"compress_loop_incl_synthesis": wall_c, "decompress_verify_loop": wall_d},
    }))
dist.destroy_process_group()
