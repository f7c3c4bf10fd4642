"""Independent read-bandwidth reference on this box: torch (CUB) reductions over HBM-resident tensors >> L2."""
import torch
x = torch.ones(210_000_000, dtype=torch.int64, device="cuda")   # 1.68 GB, the bytes Q6 streams
for name, fn in (("int64 sum", lambda: x.sum()), ("int64 max", lambda: x.max())):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name}: {ms:.4f} ms  {x.numel() * 8 / ms / 1e6:.0f} GB/s")
y = torch.empty_like(x)
for _ in range(3):
    y.copy_(x)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    y.copy_(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"copy: {ms:.4f} ms  {2 * x.numel() * 8 / ms / 1e6:.0f} GB/s (read + write)")
