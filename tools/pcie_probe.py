import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, src, dst in (("H2D", h, d), ("D2H", d, h)):
    for _ in range(2): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(8): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 8
    print(name, "1 GiB pinned: %.1f GB/s" % (n / dt / 1e9))
# chunked 256 MB like the library
c = 256 << 20
torch.cuda.synchronize(); t0 = time.perf_counter()
for k in range(16):
    h[(k % 4) * c:(k % 4 + 1) * c].copy_(d[(k % 4) * c:(k % 4 + 1) * c], non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("D2H 16 x 256 MiB: %.1f GB/s" % (16 * c / dt / 1e9))
