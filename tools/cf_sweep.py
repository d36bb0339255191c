"""Compression factor across abs_error: the reference algorithm (oracle: fp32 AE + bz2 streams + bz2 latent) against
this package's default codec on the same synthetic field.  Prints one line per threshold with the slot sizes."""
import os, pickle, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT)
import numpy as np
from oracle import lossycomp_oracle as O
from lossycomp.compress import compress
from lossycomp.decompress import decompress
from lossycomp.models import load_model

T, H, W = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (24, 240, 480)))
x = O.synthetic_field(T, H, W, seed=1234)
model = load_model(None)
raw = x[..., 0].nbytes
print(f"field {T}x{H}x{W}, {raw / 1e6:.1f} MB of channel 0")
print("abs_err |  reference (oracle+bz2): CF  latent/codes/mask B |  ours (fp16 AE, GPU coder): CF  latent/codes/mask B | ratio | max err")
for thr in (1e-3, 1e-2, 0.1, 0.3, 0.5, 1.0):
    rb = O.compress(x, thr, model)
    rs = pickle.loads(rb)
    ob = compress(x, thr)
    os_ = pickle.loads(ob)
    err = np.abs(x[..., 0].astype(np.float64) - decompress(ob)[..., 0]).max()
    print(f"{thr:7g} | {raw / len(rb):8.3f}  {len(rs[0])}/{len(rs[6])}/{len(rs[7])} | {raw / len(ob):8.3f}  "
          f"{len(os_[0])}/{len(os_[6])}/{len(os_[7])} | {len(rb) / len(ob):.3f} | {err:.6g}")
