"""Host-side cost of materialising an 8 MB payload as Python bytes (join + pickle.dumps), with and
without glibc malloc tuning (LOSSYCOMP-style mallopt).  Diagnostic for the entropy-stage assembly."""
import ctypes, pickle, sys, time
import numpy as np
import torch

tune = len(sys.argv) > 1 and sys.argv[1] == "tune"
if tune:
    libc = ctypes.CDLL("libc.so.6")
    print("mallopt mmap_threshold", libc.mallopt(-3, 32 << 20), "trim", libc.mallopt(-1, 1 << 30))
pin = torch.empty(9_000_000, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
pin.random_(0, 255)
prev = None
for i in range(8):
    n = 7_400_000 + 10_000 * i
    t0 = time.perf_counter()
    a = b"".join([b"head", memoryview(pin.numpy())[:n], b"tail"])
    t1 = time.perf_counter()
    m = b"".join([b"head", memoryview(pin.numpy())[:800_000 + 1000 * i]])
    t2 = time.perf_counter()
    blob = pickle.dumps([a, (1, 2), m, 0.2])
    t3 = time.perf_counter()
    print(f"join7.4MB {1e3*(t1-t0):.3f} ms  join0.8MB {1e3*(t2-t1):.3f} ms  pickle {1e3*(t3-t2):.3f} ms")
    prev = (a, m, blob)
print(open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
