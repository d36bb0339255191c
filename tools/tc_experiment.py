"""Timing experiments on the K3 tensor-core kernel with parts of the pipeline disabled (diagnostics)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys
sys.path.insert(0, os.path.join(%r, "lossy-ml_b200")); sys.path.insert(0, %r)
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
print("RESULT", " ".join("%%.3f" %% (m / 3) for m in ms))
''' % (ROOT, ROOT)
for dbg, label in [(0, "full"), (16, "single issuer"), (1, "no MMA"), (8, "no copies"), (9, "no MMA no copies"),
                   (103, "K4_OUT1 at 1 CTA/SM"), (105, "UP at 1 CTA/SM")]:
    env = dict(os.environ, LOSSYCOMP_TC_DEBUG=str(dbg if dbg < 100 else 0))
    if dbg >= 100:      # 10m: force mode m to one CTA per SM
        env["LOSSYCOMP_TC_ONE_PER_SM"] = str(dbg - 100)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
    print(f"debug={dbg:2d} {label:28s}", line[0] if line else out.stderr[-300:])
