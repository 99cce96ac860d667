"""Raw pinned host->device bandwidth of the box (the ceiling of the host-fed path)."""
import time, torch
for mb in (16, 256, 1024):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    n = max(4, 4096 // mb)
    t0 = time.perf_counter()
    for _ in range(n):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"pinned H2D {mb:5d} MB x{n}: {mb*n/1024/dt:6.1f} GB/s")
# 4 streams concurrently
hs = [torch.empty(64 << 20, dtype=torch.uint8).pin_memory() for _ in range(4)]
ds = [torch.empty(64 << 20, dtype=torch.uint8, device="cuda") for _ in range(4)]
ss = [torch.cuda.Stream() for _ in range(4)]
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(16):
    for h, d, s in zip(hs, ds, ss):
        with torch.cuda.stream(s):
            d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"pinned H2D 4 streams x 64 MB x16: {4*64*16/1024/dt:6.1f} GB/s")
