"""torchrun entry: time-sharded compress/decompress of one field over all ranks (NCCL), checked against
the single-GPU result on rank 0.   torchrun --nproc-per-node N tools/run_sharded.py [T H W]"""
import os, sys, pickle
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from lossycomp import _native as N, dist as D
from lossycomp._engine import Engine
from lossycomp.compress import compress_device
from lossycomp.decompress import decompress_device

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
T, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (80, 240, 480)
eng = Engine.get(None, os.environ.get("LOSSYCOMP_PRECISION", "fp16"), local)
full = torch.empty((T, H, W, 2), dtype=torch.float32, device=eng.dev)
N.check(eng.lib.lc_synth_field(full.data_ptr(), T, H, W, 0, 0, 1234, None), "synth")
t0, t1 = D.shard_time(T, world)[rank]
slots, info = D.compress_sharded(eng, full[t0:t1].contiguous(), 0.1)
a, b, rows = D.decompress_sharded(eng, slots)
err = (full[a:b, :, :, 0].double() - rows.double()).abs().max().item() if b > a else 0.0
errs = [None] * world
dist.all_gather_object(errs, err)
if rank == 0:
    ref, rinfo = compress_device(eng, full, 0.1, return_parts=True)
    same_mask = torch.equal(eng.unpack_mask(slots[7]), rinfo["mask"])
    same_q = torch.equal(eng.unpack_codes(slots[6]), rinfo["q"])
    whole = decompress_device(eng, slots)
    same_x = torch.equal(whole, decompress_device(eng, ref))
    blob = pickle.dumps(slots)
    print(f"SHARDED world={world} T={T}: mask_equal={same_mask} codes_equal={same_q} xhat_equal={same_x} "
          f"max_err={max(errs):.6f} CF={T * H * W * 4 / len(blob):.2f} (single-call CF {T * H * W * 4 / len(pickle.dumps(ref)):.2f})")
    assert same_mask and same_q and same_x and max(errs) <= 0.1
dist.destroy_process_group()
