"""cuFFT timings that decide whether the interlaced pair should be ONE complex transform (tools, not product)."""
import torch
def t(f, n=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for N in (512, 1024):
    x = torch.randn(N, N, N, device="cuda")
    print(N, "rfftn out-of-place (one real mesh) ms:", t(lambda: torch.fft.rfftn(x)))
    del x
    z = torch.randn(N, N, N, 2, device="cuda")
    zc = torch.view_as_complex(z)
    print(N, "fftn c2c out-of-place (two real meshes as one complex) ms:", t(lambda: torch.fft.fftn(zc)))
    del z, zc
    torch.cuda.empty_cache()
