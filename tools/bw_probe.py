import torch
def t(f, n=10):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
n = 1 << 30
x = torch.rand(n, device="cuda"); y = torch.rand(n, device="cuda"); z = torch.rand(n, device="cuda")
o = torch.empty(n, device="cuda")
ms = t(lambda: x.sum()); print(f"read-only sum 4.3GB: {ms:.3f} ms  {4*n/ms/1e6:.0f} GB/s")
ms = t(lambda: torch.add(x, y, out=o)); print(f"add 2r+1w: {ms:.3f} ms {12*n/ms/1e6:.0f} GB/s")
ms = t(lambda: o.copy_(x)); print(f"copy 1r+1w: {ms:.3f} ms {8*n/ms/1e6:.0f} GB/s")
ms = t(lambda: o.zero_()); print(f"memset 1w: {ms:.3f} ms {4*n/ms/1e6:.0f} GB/s")
ms = t(lambda: torch.maximum(torch.maximum(x, y), z).sum()); print(f"3 streams read (approx, with temp writes): {ms:.3f} ms")
