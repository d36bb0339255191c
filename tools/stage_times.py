"""Per-stage wall time of compress_device (device sync at every mark) on the benchmark field."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT)
import torch
from lossycomp import _native as N
from lossycomp._engine import Engine
from lossycomp.compress import compress_device
eng = Engine.get(None, "fp16", 0)
T, H, W = 24, 720, 1440
x = torch.empty((T, H, W, 2), dtype=torch.float32, device="cuda")
N.check(eng.lib.lc_synth_field(x.data_ptr(), T, H, W, 0, 0, 1234, None), "synth")
for i in range(3):
    compress_device(eng, x, 0.1)
for i in range(3):
    slots, info = compress_device(eng, x, 0.1, timing=True)
    print({k: round(v, 3) for k, v in info["timing_ms"].items()}, "sum", round(sum(info["timing_ms"].values()), 3))
torch.cuda.synchronize()
for k in range(3):
    t0 = time.perf_counter()
    for i in range(5):
        compress_device(eng, x, 0.1)
    torch.cuda.synchronize()
    print("untimed loop ms/step", round((time.perf_counter() - t0) / 5 * 1e3, 3))
eng.profile(True)
t0 = time.perf_counter()
for i in range(5):
    compress_device(eng, x, 0.1)
torch.cuda.synchronize()
print("with layer profiling ms/step", round((time.perf_counter() - t0) / 5 * 1e3, 3))
