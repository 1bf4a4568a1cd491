"""Times the (k, mu) binning kernel (row N4) on a 512^3 interlaced auto spectrum (tools, not product)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import astrild_b200 as ab
from astrild_b200 import synthetic
N, L = 512, 1000.0
dev = torch.device("cuda", 0)
pos = synthetic.zeldovich_particles(N, L, 31337, dev)
eng = ab.get_engine(N, L, dev)
m0, m1 = eng.deposit_pair(pos, None, "tsc", 1.0, "sorted")
c0, c1 = eng.r2c(m0), eng.r2c(m1)
for Nmu, poles in ((5, (0, 2, 4)), (1, ())):
    kb = eng.kmu_binning(2 * np.pi / L, None, None, Nmu, poles, (0.0, 0.0, 1.0), ("tsc", True), True)
    eng.bin_kmu(kb, c0, c1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        r = eng.bin_kmu(kb, c0, c1)
    e1.record(); torch.cuda.synchronize()
    print(f"bin_kmu 512^3 interlaced+compensated Nmu={Nmu} poles={poles}: {e0.elapsed_time(e1) / 3:.3f} ms (incl. D2H of the sums); "
          f"modes {int(r['modes'].sum())}")
b1 = eng.binning(kmin=2 * np.pi / L, compensation=("tsc", True), interlaced=True)
r1 = eng.bin_power(b1, c0, c1)
print("1-D modes", int(r1["modes"].sum()), "monopole vs 1-D max rel", float(np.nanmax(np.abs(r["poles"]["power_0"].real / r1["power"].real - 1))))
