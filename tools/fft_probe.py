import torch, time
x = torch.randn(1024, 1024, 1024, device="cuda")
def t(f, n=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("torch rfftn 1024^3 out-of-place ms:", t(lambda: torch.fft.rfftn(x)))
print("rfft over z only:", t(lambda: torch.fft.rfft(x, dim=2)))
y = torch.fft.rfft(x, dim=2)
print("fft over y (strided):", t(lambda: torch.fft.fft(y, dim=1)))
print("fft over x (strided):", t(lambda: torch.fft.fft(y, dim=0)))
print("fft2 over (y,z) r2c:", t(lambda: torch.fft.rfft2(x, dim=(1, 2))))
x5 = torch.randn(512, 512, 512, device="cuda")
print("torch rfftn 512^3 ms:", t(lambda: torch.fft.rfftn(x5), 20))
