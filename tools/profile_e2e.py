"""cProfile of the numpy-facing compress() / decompress() on the benchmark field (pinned input)."""
import os, sys, cProfile, pstats, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT)
import torch, numpy as np
from lossycomp import _native as N
from lossycomp._engine import Engine
from lossycomp.compress import compress
from lossycomp.decompress import decompress
eng = Engine.get(None, "fp16", 0)
T, H, W = 24, 720, 1440
x = torch.empty((T, H, W, 2), dtype=torch.float32, device="cuda")
N.check(eng.lib.lc_synth_field(x.data_ptr(), T, H, W, 0, 0, 1234, None), "synth")
host = torch.empty((T, H, W, 2), dtype=torch.float32, pin_memory=True); host.copy_(x); xh = host.numpy()
for i in range(3): b = compress(xh, 0.1)
ts = []
for i in range(5):
    t0 = time.perf_counter(); b = compress(xh, 0.1); ts.append((time.perf_counter() - t0) * 1e3)
print("compress e2e ms", [round(t, 2) for t in ts])
pr = cProfile.Profile(); pr.enable()
for i in range(5): b = compress(xh, 0.1)
pr.disable(); print("COMPRESS (seconds*200 = ms per call)"); pstats.Stats(pr).sort_stats("tottime").print_stats(12)
for i in range(3): y = decompress(b)
ts = []
for i in range(5):
    t0 = time.perf_counter(); y = decompress(b); ts.append((time.perf_counter() - t0) * 1e3)
print("decompress e2e ms", [round(t, 2) for t in ts])
pr = cProfile.Profile(); pr.enable()
for i in range(5): y = decompress(b)
pr.disable(); print("DECOMPRESS"); pstats.Stats(pr).sort_stats("tottime").print_stats(14)
