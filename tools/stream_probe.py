"""Where does compress_stream() lose time against the device-resident loop?  (diagnostics)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT)
import numpy as np
import torch
from lossycomp import _native as N
from lossycomp._engine import Engine
from lossycomp.compress import compress, compress_device, compress_stream, dumps_slots

eng = Engine.get(None, "fp16", 0)
T, H, W = 24, 720, 1440
x = torch.empty((T, H, W, 2), dtype=torch.float32, device="cuda")
N.check(eng.lib.lc_synth_field(x.data_ptr(), T, H, W, 0, 0, 1234, None), "synth")
host = torch.empty((T, H, W, 2), dtype=torch.float32, pin_memory=True)
host.copy_(x)
xh = host.numpy()
K = 10


def timeit(name, fn):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    print(f"{name:46s} {(time.perf_counter() - t0) / K * 1e3:7.2f} ms/step", flush=True)


timeit("compress_device loop (device resident)", lambda: [compress_device(eng, x, 0.1) for _ in range(K)])
timeit("  + dumps_slots", lambda: [dumps_slots(compress_device(eng, x, 0.1)[0]) for _ in range(K)])
copy = torch.cuda.Stream()
buf = torch.empty_like(x)


def with_background_h2d():
    for _ in range(K):
        with torch.cuda.stream(copy):
            buf.copy_(host, non_blocking=True)
        dumps_slots(compress_device(eng, x, 0.1)[0])


timeit("  + an unrelated 199 MB H2D per step (copy stream)", with_background_h2d)
timeit("compress_stream (pinned input)", lambda: list(compress_stream([xh] * K, 0.1)))
timeit("compress (pinned input, one call per field)", lambda: [compress(xh, 0.1) for _ in range(K)])
t0 = time.perf_counter()
for _ in range(K):
    buf.copy_(host, non_blocking=True)
torch.cuda.synchronize()
print(f"H2D of one field alone                         {(time.perf_counter() - t0) / K * 1e3:7.2f} ms")
