"""Debug helper: enable tensor-core modes one at a time and compare against the all-CUDA-core 16-bit path."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import golden_case
from test_gpu_tc import make_engine
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
case = sys.argv[2] if len(sys.argv) > 2 else "one_chunk"
g = golden_case(case)
x = torch.from_numpy(g["x"]).cuda()
T, H, W, _ = g["x"].shape
ref_lat = g["latent"].reshape(g["latent"].shape[0], -1)
base = make_engine(prec, True)
lat0 = base.encode(x, g["mean"], g["std"]).cpu().numpy()
rec0 = base.decode(torch.from_numpy(ref_lat).cuda(), T, H, W).cpu().numpy()
print("direct16 vs oracle: latent", np.abs(lat0 - ref_lat).max(), "recon", np.abs(rec0 - g["recon_std"][..., 0]).max())
for modes in ["1", "2", "3", "4", "5", "12345"]:
    os.environ["LOSSYCOMP_TC_MODES"] = modes
    try:
        eng = make_engine(prec)
        lat = eng.encode(x, g["mean"], g["std"]).cpu().numpy()
        rec = eng.decode(torch.from_numpy(ref_lat).cuda(), T, H, W).cpu().numpy()
        torch.cuda.synchronize()
        print(f"modes {modes}: latent max|d| vs direct16 {np.abs(lat - lat0).max():.4g}  recon max|d| {np.abs(rec - rec0).max():.4g}"
              f" mean|d| {np.abs(rec - rec0).mean():.3g}  nan {np.isnan(rec).sum() + np.isnan(lat).sum()}")
    except Exception as e:
        print("modes", modes, "FAILED", repr(e)[:300])
    finally:
        os.environ.pop("LOSSYCOMP_TC_MODES", None)
