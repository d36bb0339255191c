"""Small end-to-end cases for compute-sanitizer (one tool per run): every tensor-core mode (first conv, both 3x3x3
kernels -- register fold with residual input and TMEM accumulation --, stride-2, transposed, last conv + merge, and the
stride-2 first layer of the models without ResBlock), the CUDA-core back end, the scans and the entropy stage.
    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tools/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT)
import numpy as np
from lossycomp.compress import compress
from lossycomp.decompress import decompress
from oracle import lossycomp_oracle as O
W = os.path.join(ROOT, "lossy-ml_b200", "weights")
cases = [((17, 49, 50), "fp16", None), ((16, 48, 48), "fp32", None), ((16, 48, 96), "bf16", None),
         ((16, 48, 48), "fp16", os.path.join(W, "landsea_model.lcw"))]
if os.environ.get("SANITIZE_CASES"):
    cases = [cases[int(i)] for i in os.environ["SANITIZE_CASES"].split(",")]
for shape, prec, model in cases:
    x = O.synthetic_field(*shape, seed=5, h0=100, w0=200)
    blob = compress(x, 0.1, precision=prec, model=model)
    xhat = decompress(blob)
    print(shape, prec, os.path.basename(model or "model_2"), "max err", float(np.abs(x[..., 0] - xhat[..., 0]).max()),
          "CF", x[..., 0].nbytes / len(blob), flush=True)
