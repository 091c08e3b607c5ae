"""Raw pinned-host -> device copy bandwidth on this box (the ceiling of the e2e figure)."""
import torch
for mb in (24, 96, 192):
    h = torch.empty(mb * 1024 * 1024, dtype=torch.uint8).pin_memory()
    d = torch.empty_like(h, device="cuda")
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        d.copy_(h, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(f"H2D {mb} MiB pinned: {ms:.3f} ms  {mb * 1.048576 / ms:.1f} GB/s")
h2 = torch.empty(8 * 1024 * 1024, dtype=torch.uint8).pin_memory()
d2 = torch.empty(8 * 1024 * 1024, dtype=torch.uint8, device="cuda")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    h2.copy_(d2, non_blocking=True)
b.record()
torch.cuda.synchronize()
print(f"D2H 8 MiB: {a.elapsed_time(b) / 20:.3f} ms")
