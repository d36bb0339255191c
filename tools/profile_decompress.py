import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT)
import torch, cProfile, pstats
from lossycomp import _native as N
from lossycomp._engine import Engine
from lossycomp.compress import compress_device
from lossycomp.decompress import decompress_device
eng = Engine.get(None, "fp16", 0)
T, H, W = 24, 720, 1440
x = torch.empty((T, H, W, 2), dtype=torch.float32, device="cuda")
N.check(eng.lib.lc_synth_field(x.data_ptr(), T, H, W, 0, 0, 1234, None), "synth")
slots, info = compress_device(eng, x, 0.1)
for i in range(3): out = decompress_device(eng, slots)
torch.cuda.synchronize()
ts = []
for i in range(8):
    t0 = time.perf_counter(); out = decompress_device(eng, slots); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("decompress ms:", [round(t, 2) for t in ts])
pr = cProfile.Profile(); pr.enable()
out = decompress_device(eng, slots); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
ts = []
for i in range(6):
    t0 = time.perf_counter(); s2, _ = compress_device(eng, x, 0.1); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("compress ms:", [round(t, 2) for t in ts])
pr = cProfile.Profile(); pr.enable()
s2, _ = compress_device(eng, x, 0.1); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(16)
