import torch, time
x = torch.empty(151_552_000 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(3):
    d.copy_(x, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print(f"H2D pinned 151.6 MB: {dt*1e3:.2f} ms  {151.552/dt/1e3:.1f} GB/s")
h = torch.empty_like(x).pin_memory()
t0 = time.perf_counter()
for _ in range(10):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print(f"D2H pinned: {dt*1e3:.2f} ms  {151.552/dt/1e3:.1f} GB/s")
