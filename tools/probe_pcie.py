import torch, time
x = torch.empty(1<<30, dtype=torch.uint8, device='cuda')
h = torch.empty(1<<30, dtype=torch.uint8, pin_memory=True)
for n in (1<<30, 256<<20, 50<<20):
    torch.cuda.synchronize()
    reps = (1<<31)//n
    t0=time.perf_counter()
    for i in range(reps): h[:n].copy_(x[:n], non_blocking=True)
    torch.cuda.synchronize()
    dt=time.perf_counter()-t0
    print(f"D2H {n>>20} MiB x{reps}: {reps*n/dt/1e9:.1f} GB/s")
