"""Determinism check of the tensor-core decoder/encoder under different batch splits (diagnostics)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import golden_case
from test_gpu_tc import make_engine
g = golden_case("odd_all_axes")
x = torch.from_numpy(g["x"]).cuda()
T, H, W, _ = g["x"].shape
lat = torch.from_numpy(g["latent"].reshape(g["latent"].shape[0], -1)).cuda()
eng = make_engine("fp16")
outs = [eng.decode(lat, T, H, W, batch=b).cpu().numpy() for b in (27, 5, 1, 27, 9)]
encs = [eng.encode(x, g["mean"], g["std"], batch=b).cpu().numpy() for b in (27, 5, 1, 27)]
ref = outs[2]
print("decode diffs vs batch=1:", [float(np.abs(o - ref).max()) for o in outs], "mismatching voxels", [int((o != ref).sum()) for o in outs])
print("encode diffs vs batch=1:", [float(np.abs(o - encs[2]).max()) for o in encs])
bad = np.argwhere(outs[1] != ref)
print("mismatch coords (t,h,w) batch=5:", bad[:40].tolist())
bad = np.argwhere(outs[4] != ref)
print("mismatch coords (t,h,w) batch=9:", bad[:40].tolist())
