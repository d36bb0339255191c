"""CUDA-event timings of the non-convolution kernels on the benchmark field (diagnostics)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT)
import torch
from lossycomp import _native as N
from lossycomp._engine import Engine
from lossycomp.compress import compress_device

eng = Engine.get(None, "fp16", 0)
T, H, W = 24, 720, 1440
x = torch.empty((T, H, W, 2), dtype=torch.float32, device="cuda")
N.check(eng.lib.lc_synth_field(x.data_ptr(), T, H, W, 0, 0, 1234, None), "synth")
slots, info = compress_device(eng, x, 0.1, return_parts=True)
mean, std, mx = eng.stats(x)
recon, mask, q, latent = info["recon_std"], info["mask"], info["q"], info["latent"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(name, fn, reps=5):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"{name:34s} median {ts[len(ts)//2]:.3f} ms  min {ts[0]:.3f}")


timed("stats", lambda: eng.stats(x))
timed("encode (gather + 7 layers)", lambda: eng.encode(x, mean, std))
eng.profile(True); eng.profile_read(); eng.encode(x, mean, std); ms, _ = eng.profile_read(); eng.profile(False)
print("   of which conv layers", round(sum(ms[:7]), 3), "ms -> gather ~ the rest")
timed("quantize", lambda: eng.quantize(x, recon, mean, std, 0.0999))
timed("quantize + fused check", lambda: eng.quantize(x, recon, mean, std, 0.0999, check_abs_err=0.1))
timed("reconstruct", lambda: eng.reconstruct(recon, mask, q, mean, std, 0.1998))
timed("pack_streams", lambda: eng.pack_streams(mask, q))

import ctypes as C
t = torch
n_m, n_q = mask.numel(), q.numel()
sym = t.empty((n_m + 4) // 5, dtype=t.uint8, device="cuda")
h = t.empty(768, dtype=t.int32, device="cuda")
s0 = C.c_void_p(t.cuda.current_stream().cuda_stream)
P = lambda v: C.c_void_p(v.data_ptr())
timed("pack_trits_hist", lambda: eng.lib.lc_pack_trits_hist(P(mask), n_m, P(sym), P(h), s0))
timed("hist_codes", lambda: eng.lib.lc_hist_codes(P(q), n_q, 255, 0, 127, C.c_void_p(h.data_ptr() + 1024), s0))
symq = t.empty(n_q, dtype=t.uint8, device="cuda"); esc = t.empty(1024, dtype=t.int32, device="cuda")
nesc = t.zeros(1, dtype=t.int64, device="cuda")
ws = t.empty(eng.lib.lc_scan_workspace_bytes(n_q), dtype=t.uint8, device="cuda")
timed("symbolize (plain)", lambda: eng.lib.lc_symbolize_i32(P(q), n_q, 0, 0, 255, P(symq), P(esc), P(nesc), P(ws), s0))
q2 = t.empty_like(q)
timed("unsymbolize (plain)", lambda: eng.lib.lc_unsymbolize_i32(P(symq), n_q, 0, 0, 255, P(esc), P(q2), P(ws), s0))
assert t.equal(q2, q)
timed("decode (7 layers + merge)", lambda: eng.decode(latent, T, H, W))
from lossycomp.decompress import decompress_device
timed("decompress_device", lambda: decompress_device(eng, slots))
