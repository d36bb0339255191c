"""Where does the per-step time go at N > 1?  Run under torchrun; every rank times the stages of the
bench step (compress_device / frame / stage+sizes / H2D+gather) with a device sync after each and
rank 0 prints the per-rank means.  Diagnostic only (the syncs perturb the pipeline)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200"))

import torch
import torch.distributed as dist

from lossycomp import _native as N
from lossycomp import dist as LD
from lossycomp._engine import Engine
from lossycomp.compress import compress_device


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine.get(None, "fp16", local)
    T, H, W = 24, 721, 1440
    x = torch.empty((T, H, W, 2), dtype=torch.float32, device=eng.dev)
    N.check(eng.lib.lc_synth_field(x.data_ptr(), T, H, W, 24 * rank, 0, 1234, None), "lc_synth_field")
    torch.cuda.synchronize()
    steps = 8
    acc = {}

    def tick(name, t0):
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        acc[name] = acc.get(name, 0.0) + (t1 - t0) * 1e3
        return t1

    for it in range(3 + steps):
        if it == 3:
            acc.clear()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        slots, info = compress_device(eng, x, 0.1, mode="strict", codec="huff")
        t = tick("compress", t)
        if world > 1:
            fr = LD.frame_slots(slots)
            t = tick("frame", t)
            dist.barrier()
            t = tick("skew(barrier)", t)
            LD.gather_bytes(fr, 0, None, eng.dev, to_host=False)
            t = tick("gather_bytes", t)
    rows = torch.tensor([acc.get(k, 0.0) / steps for k in ("compress", "frame", "skew(barrier)", "gather_bytes")],
                        dtype=torch.float64, device=eng.dev)
    if world > 1:
        allr = [torch.zeros_like(rows) for _ in range(world)]
        dist.all_gather(allr, rows)
    else:
        allr = [rows]
    if rank == 0:
        print("rank  compress  frame  skew  gather_bytes   (ms, mean of %d steps)" % steps)
        for r, v in enumerate(allr):
            print(r, " ".join(f"{float(a):8.3f}" for a in v.cpu()))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
