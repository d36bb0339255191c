import torch, time
x = torch.empty(1600 << 20, dtype=torch.uint8, device="cuda")
y = torch.empty(1600 << 20, dtype=torch.uint8, device="cuda")
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: x.zero_()); print(f"memset 1.68 GB: {ms:.3f} ms -> {x.numel()/ms/1e9:.2f} TB/s write")
ms = t(lambda: y.copy_(x)); print(f"copy 1.68 GB: {ms:.3f} ms -> {2*x.numel()/ms/1e9:.2f} TB/s read+write")
ms = t(lambda: x.sum(dtype=torch.int64)); print(f"read (sum u8) 1.68 GB: {ms:.3f} ms -> {x.numel()/ms/1e9:.2f} TB/s read")
xf = x.view(torch.float32)
ms = t(lambda: xf.sum()); print(f"read (sum f32) 1.68 GB: {ms:.3f} ms -> {x.numel()/ms/1e9:.2f} TB/s read")
