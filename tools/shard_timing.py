"""torchrun entry: where the time of dist.compress_sharded goes on every rank (stage marks with device syncs)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT)
os.environ["LOSSYCOMP_SHARD_TIMING"] = "1"
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
for i in range(6):
    dist.barrier(); torch.cuda.synchronize()
    slots, info = D.compress_sharded(eng, x, 0.1, t0=t0, T=T, dst=dst)
    if i >= 4:
        print(f"rank {rank} rows {t1 - t0}: " + " ".join(f"{k}={v:.2f}" for k, v in info["shard_timing_ms"].items()), flush=True)
dist.destroy_process_group()
