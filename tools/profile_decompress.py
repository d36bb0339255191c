"""Device-resident decompress timing on the benchmark field (CUDA events, 10 repetitions after warm-up)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT)
import torch
from lossycomp import _native as N
from lossycomp._engine import Engine
from lossycomp.compress import compress_device
from lossycomp.decompress import decompress_device
eng = Engine.get(None, "fp16", 0)
T, H, W = 24, 720, 1440
x = torch.empty((T, H, W, 2), dtype=torch.float32, device="cuda")
N.check(eng.lib.lc_synth_field(x.data_ptr(), T, H, W, 0, 0, 1234, None), "synth")
slots, _ = compress_device(eng, x, 0.1)
for i in range(3): decompress_device(eng, slots)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10): y = decompress_device(eng, slots)
e1.record(); torch.cuda.synchronize()
print("decompress_device ms/step", round(e0.elapsed_time(e1) / 10, 3))
