"""One warm-up + one timed P(k) step of a bench workload, for ncu captures (tools, not product)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import astrild_b200 as ab
from astrild_b200 import synthetic
from bench import WORKLOADS

wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n, N, L = wl["n"], wl["mesh"], wl["box"]
dev = torch.device("cuda", 0)
pos = synthetic.zeldovich_particles(n, L, wl["seed"], dev) if wl["kind"] == "zeldovich" else synthetic.uniform_particles(n, wl["seed"], dev)
torch.cuda.empty_cache()
eng = ab.get_engine(N, L, dev)
comp = (wl["resampler"], wl["interlaced"]) if wl["compensated"] else None
binning = eng.binning(kmin=2 * np.pi / L, compensation=comp, interlaced=wl["interlaced"])
mesh1 = eng.new_mesh()
mesh2 = eng.new_mesh() if wl["interlaced"] else None
eng.ensure_workspace(n ** 3, False, wl["interlaced"])
for _ in range(steps):
    if mesh2 is not None:
        eng.deposit_pair(pos, None, wl["resampler"], 1.0, "sorted", out=(mesh1, mesh2))
    else:
        eng.deposit(pos, None, wl["resampler"], 0.0, 1.0, "sorted", out=mesh1)
    c1 = eng.r2c(mesh1)
    c1s = eng.r2c(mesh2) if mesh2 is not None else None
    res = eng.bin_power(binning, c1, c1s, scale=1.0)
torch.cuda.synchronize()
print("ok", res["modes"][:3], res["power"].real[:3])
