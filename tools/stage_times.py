"""Print the per-stage wall time of compress_device on the benchmark field (diagnostics)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT)
import torch
from lossycomp import _native as N
from lossycomp._engine import Engine
from lossycomp.compress import compress_device
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
eng = Engine.get(None, prec, 0)
T, H, W = 24, 720, 1440
x = torch.empty((T, H, W, 2), dtype=torch.float32, device="cuda")
N.check(eng.lib.lc_synth_field(x.data_ptr(), T, H, W, 0, 0, 1234, None), "synth")
for i in range(3):
    slots, info = compress_device(eng, x, 0.1, timing=True)
print({k: round(v, 2) for k, v in info["timing_ms"].items()}, "total", round(sum(info["timing_ms"].values()), 2))
import cProfile, pstats
for i in range(3):
    compress_device(eng, x, 0.1)
pr = cProfile.Profile(); pr.enable()
for i in range(5):
    slots, info = compress_device(eng, x, 0.1)
torch.cuda.synchronize(); pr.disable()
print("per-call ms = listed seconds * 200")
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
