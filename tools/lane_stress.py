"""Stress of two lanes on one GPU: N compress + N decompress jobs of the benchmark field, two in flight.
   python tools/lane_stress.py [N] [mode]   mode: both | comp | dec"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT)
import torch
from lossycomp import _native as N
from lossycomp.compress import compress_device
from lossycomp.decompress import decompress_device
from lossycomp.lanes import LaneSet
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
mode = sys.argv[2] if len(sys.argv) > 2 else "both"
ls = LaneSet.get(None, "fp16", 0, int(os.environ.get("LANES", "2")))
eng = ls.lanes[0].eng
T, H, W = 24, 720, 1440
x = torch.empty((T, H, W, 2), dtype=torch.float32, device="cuda")
N.check(eng.lib.lc_synth_field(x.data_ptr(), T, H, W, 0, 0, 1234, None), "synth")
torch.cuda.synchronize()
slots, rinfo = compress_device(eng, x, 0.1, return_parts=True)
rparts = {k: rinfo[k].clone() for k in ("latent", "recon_std", "mask", "q")}
ref = decompress_device(eng, slots).clone()
torch.cuda.synchronize()


def cjob(lane):
    s, info = compress_device(lane.eng, x, 0.1, return_parts=True)
    bad = [k for k in ("latent", "recon_std", "mask", "q")
           if info[k].shape != rparts[k].shape or not bool(torch.equal(info[k], rparts[k]))]
    if not bad and not (s[6] == slots[6] and s[7] == slots[7] and s[0] == slots[0]):
        bad = ["slots:" + ",".join(str(i) for i in (0, 6, 7) if s[i] != slots[i])]
    if bad:
        k = bad[0]
        if k in rparts and info[k].shape == rparts[k].shape:
            d = (info[k].reshape(-1) != rparts[k].reshape(-1)).nonzero().reshape(-1)
            print("MISMATCH", bad, "count", int(d.numel()), "first", d[:4].tolist(), "last", d[-4:].tolist(), flush=True)
            a = info[k].reshape(-1)[d[:12]]
            b_ = rparts[k].reshape(-1)[d[:12]]
            print("  got", a.tolist(), "\n  ref", b_.tolist(), flush=True)
            if a.dtype == torch.float32:
                print("  int diff", (a.view(torch.int32) - b_.view(torch.int32)).tolist(), flush=True)
        else:
            print("MISMATCH", bad, flush=True)
    return ("c", not bad)


def djob(lane):
    out = decompress_device(lane.eng, slots)
    ok = bool(torch.equal(out, ref))
    torch.cuda.current_stream(lane.eng.dev).synchronize()
    return ("d", ok)


t0 = time.time()
jobs = []
for i in range(n):
    if mode in ("both", "comp"):
        jobs.append(ls.submit(cjob))
    if mode in ("both", "dec"):
        jobs.append(ls.submit(djob))
    while len(jobs) > len(ls):
        kind, ok = jobs.pop(0).result()
        assert ok, (kind, i)
for j in jobs:
    kind, ok = j.result()
    assert ok, kind
torch.cuda.synchronize()
print("STRESS OK", n, mode, "%.1f s" % (time.time() - t0), flush=True)
