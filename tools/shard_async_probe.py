"""torchrun entry: per-step wall time of dist.compress_sharded in its synchronous and asynchronous forms."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from lossycomp import _native as N, dist as D
from lossycomp._engine import Engine

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
hours = int(sys.argv[1]) if len(sys.argv) > 1 else 24
T, H, W = hours * world, 720, 1440
eng = Engine.get(None, "fp16", local)
t0, t1 = D.shard_time(T, world)[rank]
x = torch.empty((t1 - t0, H, W, 2), dtype=torch.float32, device=eng.dev)
N.check(eng.lib.lc_synth_field(x.data_ptr(), t1 - t0, H, W, t0, 0, 1234, None), "synth")
dst = world - 1
K = 6
for mode in ("sync", "async_end", "async_each", "async_end"):
    for _ in range(2):
        out, _ = D.compress_sharded(eng, x, 0.1, t0=t0, T=T, dst=dst)
    dist.barrier(); torch.cuda.synchronize()
    marks = [time.perf_counter()]
    pend = []
    for i in range(K):
        out, _ = D.compress_sharded(eng, x, 0.1, t0=t0, T=T, dst=dst, wait=(mode == "sync"))
        if mode == "async_each" and out is not None:
            out = out.result()
        pend.append(out)
        marks.append(time.perf_counter())
    for p in pend:
        if hasattr(p, "result"):
            p.result()
    marks.append(time.perf_counter())
    torch.cuda.synchronize()
    end = time.perf_counter()
    print(f"rank {rank} {mode:10s} total {1e3 * (end - marks[0]) / K:.2f} ms/step; steps " +
          " ".join(f"{1e3 * (b - a):.1f}" for a, b in zip(marks, marks[1:])), flush=True)
dist.destroy_process_group()
