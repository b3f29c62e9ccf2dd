"""Host->device copy rate of this box, alone and under a running GEMM load (context for bench.py's e2e)."""
import time
import torch

x = torch.empty((4096, 4096), dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
s = torch.cuda.Stream()
for load in (False, True):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    for i in range(n):
        if load:
            for _ in range(3):
                a @ a
        with torch.cuda.stream(s):
            d.copy_(x, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("H2D 64 MiB x %d %s: %.2f ms per copy-slot (%.1f GB/s if copy-bound)" % (
        n, "under GEMM load" if load else "alone", 1e3 * dt / n, n * x.numel() * 4 / dt / 1e9))
