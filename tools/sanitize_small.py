"""Small run of every hand-written kernel family for compute-sanitizer memcheck (tools, not product)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import astrild_b200 as ab
from astrild_b200 import ingest

N, L = 48, 300.0
rng = np.random.default_rng(3)
pos = (rng.random((120001, 3)) * L).astype(np.float32)           # odd count: tail tiles
mass = (rng.random(120001) + 0.5).astype(np.float32)
for kw in (dict(resampler="tsc", interlaced=True, compensated=True), dict(resampler="cic", interlaced=False, compensated=False)):
    for w in (None, mass):
        for method in ("sorted", "atomic"):
            m = ab.CatalogMesh(pos, L, N, weight=w, normalize=True, method=method, **kw)
            r = ab.FFTPower(m, mode="1d", kmin=2 * np.pi / L)
m = ab.CatalogMesh(pos, L, N, resampler="tsc", interlaced=True, compensated=True, normalize=True, fold=2)
r = ab.FFTPower(m, mode="2d", Nmu=4, poles=[0, 2, 4], kmin=2 * np.pi / (L / 4), los=[0.0, 0.6, 0.8])
vm = rng.normal(1.0, 0.3, (N, N, N))
ps = ab.FFTPower(ab.ArrayMesh(vm, BoxSize=L), mode="1d", second=ab.ArrayMesh(vm.astype(np.float32) * 2, BoxSize=L))
g = ingest.assign_grid(rng.random(5000), rng.random(5000), rng.random(5000), rng.random(5000), 16)
b = ab.PkBatch(N, L, chunk_rows=30000).run([(1, tuple(np.ascontiguousarray(pos[:, d]) for d in range(3))), (2, pos, mass)])
# slab routing / ghost / transpose kernels with ranks emulated in one process (sequentially is enough for memcheck)
from astrild_b200 import distributed
be = distributed.CudaSlabBackend(N, L, 12, 12, 4, "cuda:0")
sp, sm, counts = be.route(torch.from_numpy(pos).cuda(), torch.from_numpy(mass).cuda(), 1.0 / L)
meshes = be.deposit_pair(torch.from_numpy(pos).cuda(), None, "tsc", 1.0 / L)
torch.cuda.synchronize()
print("sanitize_small ok", int(sum(counts)), float(meshes[0].sum()))
