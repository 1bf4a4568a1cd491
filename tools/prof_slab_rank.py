"""Compute stages of ONE rank of the slab-decomposed path on one GPU (tools, not product).

usage: prof_slab_rank.py [workload=c3] [P=8] [rank=0] [steps=3]
The particles that the neighbours would send are produced here with torch (untimed); the exchanges themselves
(all-to-all-v, ghost ring, transpose, all-reduce) are not run -- this isolates the per-rank kernel time.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from astrild_b200 import distributed, synthetic
from bench import WORKLOADS

wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
P = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
n, N, L = wl["n"], wl["mesh"], wl["box"]
dev = torch.device("cuda", 0)
n0 = N // P
x0 = rank * n0


def planes(r):
    return (r * n) // P, ((r + 1) * n) // P


def gen(r):
    a, b = planes(r % P)
    if wl["kind"] == "zeldovich":
        return synthetic.zeldovich_particles(n, L, wl["seed"], dev, x_planes=(a, b))
    return synthetic.sine_displaced_particles(n, wl["seed"], dev, x_planes=(a, b))


own = gen(rank)
recv = []
for r in (rank - 1, rank + 1):
    q = gen(r)
    cell = torch.floor(q[0].double() * N).long() % N
    m = (cell >= x0) & (cell < x0 + n0)
    recv.append(torch.stack([c[m] for c in q], dim=1))
    del q, cell, m
recv = torch.cat(recv).contiguous()
torch.cuda.empty_cache()
print(f"rank {rank}/{P}: own {own[0].numel()} particles, received {recv.shape[0]}")

be = distributed.CudaSlabBackend(N, L, x0, n0, P, dev)
comp = (wl["resampler"], wl["interlaced"]) if wl["compensated"] else None
binning = be.make_binning(x0, n0, 2 * np.pi / L, None, None, comp, wl["interlaced"])
lo, hi = 1, 2


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


for it in range(steps):
    marks = [("start", ev())]
    sp, sm, counts = be.route(own, None, 1.0)
    marks.append(("route", ev()))
    if wl["interlaced"]:
        pair = be.deposit_pair(own, None, wl["resampler"], 1.0)
        marks.append(("deposit_own", ev()))
        pair = be.deposit_pair(recv, None, wl["resampler"], 1.0, out=pair)
        marks.append(("deposit_recv", ev()))
        owned = [m[lo: lo + n0] for m in pair]
    else:
        mesh = be.deposit(own, None, wl["resampler"], 0.0, 1.0)
        marks.append(("deposit_own", ev()))
        mesh = be.deposit(recv, None, wl["resampler"], 0.0, 1.0, out=mesh)
        marks.append(("deposit_recv", ev()))
        owned = [mesh[lo: lo + n0]]
    for o in owned:
        be.accumulate(o[n0 - lo:], o[:lo].clone())
    marks.append(("ghost_add", ev()))
    total = be.mesh_sum(owned[0])
    grids = [be.fft2d(o) for o in owned]
    marks.append(("fft2d", ev()))
    grids = [g.reshape(n0, P, n0, N // 2 + 1).permute(1, 0, 2, 3).contiguous().reshape(N, n0, N // 2 + 1) for g in grids]
    marks.append(("local_pack", ev()))
    grids = [be.fft1d(g, n0) for g in grids]
    marks.append(("fft1d", ev()))
    raw = be.bin(binning, grids[0], grids[1] if wl["interlaced"] else None)
    marks.append(("bin", ev()))
    red = be.to_reduce_tensor(raw, total).cpu()
    marks.append(("d2h", ev()))
    torch.cuda.synchronize()
    prof = {b[0]: round(a[1].elapsed_time(b[1]), 3) for a, b in zip(marks[:-1], marks[1:])}
    print(it, "leavers", int(sum(counts)), "total ms", round(marks[0][1].elapsed_time(marks[-1][1]), 3), prof)
