"""Timing experiments on the tensor-core kernels with parts of the pipeline disabled (diagnostics).
   python tools/tc_experiment.py [label=ENV1=v1,ENV2=v2 ...]   (default: the standard ablation list)"""
import os, sys, subprocess
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
print("RESULT", " ".join("%%.3f" %% (m / 3) for m in ms), "| sum %%.3f" %% (sum(ms) / 3))
''' % (ROOT, ROOT)
configs = [("full", {}), ("no MMA", {"LOSSYCOMP_TC_DEBUG": "1"}), ("no copies", {"LOSSYCOMP_TC_DEBUG": "8"}),
           ("no MMA no copies", {"LOSSYCOMP_TC_DEBUG": "9"}), ("old K3 kernel", {"LOSSYCOMP_NO_ACC": "1"})]
if len(sys.argv) > 1:
    configs = []
    for a in sys.argv[1:]:
        label, _, envs = a.partition("=")
        configs.append((label, dict(e.split("=", 1) for e in envs.split(",") if e)))
for label, envs in configs:
    env = dict(os.environ, **envs)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
    print(f"{label:28s}", line[0] if line else out.stderr[-300:], flush=True)
