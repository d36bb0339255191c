import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT)
import torch, numpy as np, cProfile, pstats
from lossycomp import _native as N, huffman as HF
from lossycomp._engine import Engine
from lossycomp.compress import compress_device
from lossycomp.decompress import decompress_device
eng = Engine.get(None, "fp16", 0)
T, H, W = 24, 720, 1440
x = torch.empty((T, H, W, 2), dtype=torch.float32, device="cuda")
N.check(eng.lib.lc_synth_field(x.data_ptr(), T, H, W, 0, 0, 1234, None), "synth")
slots, info = compress_device(eng, x, 0.1, return_parts=True)
mask, q = info["mask"], info["q"]
for i in range(3): eng.pack_streams(mask, q)
torch.cuda.synchronize()
ts = []
for i in range(8):
    t0 = time.perf_counter(); eng.pack_streams(mask, q); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("pack_streams ms", [round(t, 2) for t in ts])
pr = cProfile.Profile(); pr.enable()
for i in range(5): eng.pack_streams(mask, q)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
for i in range(3): decompress_device(eng, slots)
torch.cuda.synchronize()
ts = []
for i in range(10):
    t0 = time.perf_counter(); decompress_device(eng, slots); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("decompress ms", [round(t, 2) for t in ts])
# interleave like bench: compress x5 then decompress x5 with events
for rep in range(3):
    for i in range(5): slots, _ = compress_device(eng, x, 0.1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5): decompress_device(eng, slots)
    e1.record(); torch.cuda.synchronize()
    print("after compress loop: decompress ms/step", round(e0.elapsed_time(e1) / 5, 2))
