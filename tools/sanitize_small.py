"""Small end-to-end case for compute-sanitizer (one tool per run)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT)
import numpy as np
from lossycomp.compress import compress
from lossycomp.decompress import decompress
from oracle import lossycomp_oracle as O
for shape, prec in (((17, 49, 50), "fp16"), ((16, 48, 48), "fp32"), ((33, 100, 97), "bf16")):
    x = O.synthetic_field(*shape, seed=5, h0=100, w0=200)
    blob = compress(x, 0.1, precision=prec)
    xhat = decompress(blob)
    print(shape, prec, "max err", float(np.abs(x[..., 0] - xhat[..., 0]).max()), "CF", x[..., 0].nbytes / len(blob))
