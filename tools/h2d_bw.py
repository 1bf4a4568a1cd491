import time, torch
torch.cuda.init()
for gb in (1, 4):
    n = gb * (1 << 28)
    h = torch.empty(n, dtype=torch.float32, pin_memory=True); h.fill_(1.0)
    d = torch.empty(n, dtype=torch.float32, device="cuda")
    for rep in range(3):
        torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize()
        print(f"H2D {gb} GiB pinned: {gb*1.0737/(time.perf_counter()-t):.1f} GB/s")
    torch.cuda.synchronize(); t = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize()
    print(f"D2H {gb} GiB pinned: {gb*1.0737/(time.perf_counter()-t):.1f} GB/s")
    # two streams concurrently
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    half = n // 2
    torch.cuda.synchronize(); t = time.perf_counter()
    with torch.cuda.stream(s1): d[:half].copy_(h[:half], non_blocking=True)
    with torch.cuda.stream(s2): d[half:].copy_(h[half:], non_blocking=True)
    torch.cuda.synchronize()
    print(f"H2D {gb} GiB pinned, 2 streams: {gb*1.0737/(time.perf_counter()-t):.1f} GB/s")
t=time.perf_counter(); x = torch.empty(3 << 30, dtype=torch.float32, pin_memory=True); print("pin 12 GiB alloc s:", time.perf_counter()-t)
