import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lossy-ml_b200")); sys.path.insert(0, ROOT)
import torch, numpy as np
from lossycomp import _native as N, huffman as HF
from lossycomp._engine import Engine
from lossycomp.compress import compress_device
eng = Engine.get(None, "fp16", 0)
T, H, W = 24, 720, 1440
x = torch.empty((T, H, W, 2), dtype=torch.float32, device="cuda")
N.check(eng.lib.lc_synth_field(x.data_ptr(), T, H, W, 0, 0, 1234, None), "synth")
slots, info = compress_device(eng, x, 0.1, return_parts=True)
mask, q = info["mask"], info["q"]
def timed(label, fn, n=5):
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{label:32s} {min(ts):7.3f} ms (min of {n})"); return r
timed("pack_mask total", lambda: eng.pack_mask(mask))
timed("pack_codes total", lambda: eng.pack_codes(q))
t = torch
n = mask.numel()
sym = t.empty((n + 4) // 5, dtype=t.uint8, device=eng.dev)
timed("  pack_trits kernel", lambda: N.check(eng.lib.lc_pack_trits(mask.data_ptr(), n, sym.data_ptr(), None), "x"))
hist = timed("  hist (kernel + D2H)", lambda: eng.hist(sym))
table = timed("  build_table", lambda: HF.build_table(hist))
bits = table.encoded_bits(hist)
res = timed("  huff_pack (alloc+kernel+D2H)", lambda: eng.huff_pack(sym, table, bits))
payload, lens, tb = res
timed("  table.to_bytes + join", lambda: b"".join([b"LCH1", table.to_bytes(), lens.tobytes(), payload]))
out = t.zeros(((bits + 31) // 32) * 4 + 16, dtype=t.uint8, device=eng.dev)
timed("    torch.zeros(out)", lambda: t.zeros(((bits + 31) // 32) * 4 + 16, dtype=t.uint8, device=eng.dev))
timed("    out.cpu()", lambda: out.cpu())
timed("    out.cpu().numpy().tobytes()", lambda: out.cpu().numpy().tobytes())
pin = t.empty(out.numel(), dtype=t.uint8, pin_memory=True)
timed("    pinned copy_", lambda: pin.copy_(out, non_blocking=False))
timed("    bytes(pinned)", lambda: bytes(memoryview(pin.numpy())))
import pickle
timed("pickle.dumps(slots)", lambda: pickle.dumps(slots))
