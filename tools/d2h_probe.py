"""How should decompress() hand 100 MB back to the host?  Times device->host variants (diagnostic)."""
import time
import numpy as np
import torch

n = 24 * 720 * 1440
d = torch.rand(n, device="cuda")
pin = torch.empty(n, dtype=torch.float32, pin_memory=True)
warm = np.empty(n, dtype=np.float32); warm[:] = 0


def t(f, reps=4):
    out = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize()
        out.append((time.perf_counter() - t0) * 1e3)
    return " ".join(f"{v:.2f}" for v in out)


def fresh_direct():
    h = np.empty(n, dtype=np.float32); torch.from_numpy(h).copy_(d); return h


def pinned_then_copy():
    pin.copy_(d, non_blocking=True); torch.cuda.synchronize()
    h = np.empty(n, dtype=np.float32); np.copyto(h, pin.numpy()); return h


def pinned_chunked_copy(parts=8):
    h = np.empty(n, dtype=np.float32); hp = pin.numpy()
    step = (n + parts - 1) // parts
    evs = []
    for i in range(parts):
        pin[i * step:(i + 1) * step].copy_(d[i * step:(i + 1) * step], non_blocking=True)
        e = torch.cuda.Event(); e.record(); evs.append(e)
    for i, e in enumerate(evs):
        e.synchronize(); np.copyto(h[i * step:(i + 1) * step], hp[i * step:(i + 1) * step])
    return h


print("fresh np.empty + direct D2H :", t(fresh_direct))
print("pinned D2H + copy to fresh  :", t(pinned_then_copy))
print("pinned chunked + overlap    :", t(pinned_chunked_copy))
print("direct D2H into warm array  :", t(lambda: torch.from_numpy(warm).copy_(d)))
print("pinned D2H only             :", t(lambda: pin.copy_(d, non_blocking=True)))
print("tensor.cpu()                :", t(lambda: d.cpu()))
