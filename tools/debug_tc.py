"""Debug helper: compare the bf16 CUDA-core back end and the tcgen05 back end against the fp32 oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import golden_case
from test_gpu_tc import make_engine
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
g = golden_case("one_chunk")
x = torch.from_numpy(g["x"]).cuda()
T, H, W, _ = g["x"].shape
ref_lat = g["latent"].reshape(1, -1)
for name, eng in (("direct16", make_engine(prec, True)), ("tc", make_engine(prec))):
    lat = eng.encode(x, g["mean"], g["std"]).cpu().numpy()
    rec = eng.decode(torch.from_numpy(ref_lat).cuda(), T, H, W).cpu().numpy()
    print(name, "latent max|d|", np.abs(lat - ref_lat).max(), "ref max", np.abs(ref_lat).max(),
          "| recon max|d|", np.abs(rec - g["recon_std"][..., 0]).max(), "mean|d|", np.abs(rec - g["recon_std"][..., 0]).mean())
