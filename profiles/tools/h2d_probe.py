import torch, time
n = 403 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(3): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print(f"H2D pinned 403 MiB: {dt*1e3:.3f} ms = {n/dt/1e9:.1f} GB/s")
