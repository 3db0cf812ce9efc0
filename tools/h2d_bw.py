import torch, time
x = torch.empty(1 << 28, dtype=torch.float32, pin_memory=True)   # 1 GiB
y = torch.empty_like(x, device="cuda")
for _ in range(2): y.copy_(x, non_blocking=True)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5): y.copy_(x, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t
print("pinned H2D 1 GiB x5: %.1f GB/s" % (5 * x.numel() * 4 / dt / 1e9))
