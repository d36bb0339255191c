"""One timing line of the autoencoder layers under the current environment (stderr passes through)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT)
import torch
from lossycomp import _native as N
from lossycomp._engine import Engine
eng = Engine.get(None, "fp16", 0)
T, H, W = 24, 720, 1440
x = torch.empty((T, H, W, 2), dtype=torch.float32, device="cuda")
N.check(eng.lib.lc_synth_field(x.data_ptr(), T, H, W, 0, 0, 1234, None), "synth")
mean, std, _ = eng.stats(x)
for i in range(2):
    lat = eng.encode(x, mean, std); rec = eng.decode(lat, T, H, W)
eng.profile(True); eng.profile_read()
for i in range(3):
    lat = eng.encode(x, mean, std); rec = eng.decode(lat, T, H, W)
ms, cnt = eng.profile_read()
print(sys.argv[1] if len(sys.argv) > 1 else "", "RESULT", " ".join("%.3f" % (m / 3) for m in ms), "| sum %.3f" % (sum(ms) / 3), flush=True)
