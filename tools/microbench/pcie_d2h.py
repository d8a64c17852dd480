import torch, time
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
for sz in (n, n // 8):
    for _ in range(2):
        torch.cuda.synchronize(); t = time.perf_counter()
        for o in range(0, n, sz):
            h[o:o + sz].copy_(d[o:o + sz], non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("D2H", sz >> 20, "MiB pieces:", round(n / dt / 1e9, 1), "GB/s")
