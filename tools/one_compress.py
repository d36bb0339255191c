"""One warm compress_device of the benchmark field (for ncu captures):  python tools/one_compress.py [T H W]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT)
import torch
from lossycomp import _native as N
from lossycomp._engine import Engine
from lossycomp.compress import compress_device
from lossycomp.decompress import decompress_device
T, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (24, 720, 1440)
eng = Engine.get(None, os.environ.get("LOSSYCOMP_PRECISION", "fp16"), 0)
x = torch.empty((T, H, W, 2), dtype=torch.float32, device="cuda")
N.check(eng.lib.lc_synth_field(x.data_ptr(), T, H, W, 0, 0, 1234, None), "synth")
for _ in range(int(os.environ.get("REPS", "2"))):
    slots, info = compress_device(eng, x, 0.1)
    xhat = decompress_device(eng, slots)
torch.cuda.synchronize()
print("violations", info.get("violations"), "n_out", info["n_out"])
