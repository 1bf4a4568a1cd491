"""CPU model of the tile kernel's shared-memory bank conflicts on the bench's particle set (tools, not product).

The tile kernel's time follows its ATOMS wavefronts = the largest number of lanes on one bank, per instruction
(DESIGN.md 4.1).  This script rebuilds what a warp of the kernel sees -- a Zel'dovich set in lattice order, filed into
bricks in the order the scatter's warp-runs arrive (random within the window of tiles in flight), 32 consecutive payload
entries per warp -- and evaluates walking orders of the 27 TSC updates by max lanes per bank, averaged over the steps.
Round 2 (n = 256, 12 x 6 x 29 bricks): plain walk 3.13 (ncu: 3.10 at 1024^3), x-planes rotated by the rank among lanes
starting on the same bank, x-planes 11 banks apart: 2.13 (ncu: 2.61), x and y by two digits of the rank: 2.19, greedy
choice among the 27 digit-wise rotations: 2.06, z-rotation only on the unskewed tile: 2.69.

usage: python tools/bank_model.py [n_side=256] [bx=12] [by=6]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from astrild_b200 import synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
BX = int(sys.argv[2]) if len(sys.argv) > 2 else 12
BY = int(sys.argv[3]) if len(sys.argv) > 3 else 6
ZC = 29                                                   # TSC, interlaced pair
pos = synthetic.zeldovich_particles(n, 1000.0 * n / 1024, 31337, "cpu")
g = [p.numpy().astype(np.float64) * n for p in pos]
h = [np.rint(x).astype(np.int64) % n for x in g]          # mesh-0 home cells
bx, by, bz = h[0] // BX, h[1] // BY, h[2] // ZC
nby, nbz = (n + BY - 1) // BY, (n + ZC - 1) // ZC
key = (bx * nby + by) * nbz + bz
hl = [h[0] - bx * BX, h[1] - by * BY, h[2] - bz * ZC]      # window origin in tile coordinates

# arrival order of the scatter: warp-slices of 32 consecutive particles, random within 1184 tiles x 32 slices in flight
rng = np.random.default_rng(1)
nslice = n ** 3 // 32
prio = (np.arange(nslice) // (1184 * 32)) + rng.random(nslice)
pidx = (np.argsort(prio, kind="stable")[:, None] * 32 + np.arange(32)[None]).ravel()
order = pidx[np.argsort(key[pidx], kind="stable")]
ks, H = key[order], [a[order] for a in hl]
starts = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]])
ends = np.r_[starts[1:], len(ks)]

A, B, C = (a.ravel()[None] for a in np.meshgrid(range(3), range(3), range(3), indexing="ij"))


def wavefronts(banks):
    return sum(np.bincount(banks[:, s] % 32, minlength=32).max() for s in range(banks.shape[1])) / banks.shape[1]


def walk(b0, rx, ry, rz, sx, sy):
    return b0[:, None] + sx * ((A + rx[:, None]) % 3) + sy * ((B + ry[:, None]) % 3) + (C + rz[:, None]) % 3


def rank_among_equal(b0):
    return np.array([np.sum(b0[:i] == b0[i]) for i in range(len(b0))])


res = {}
z = np.zeros(32, int)
for b in rng.choice(len(starts), min(300, len(starts)), replace=False):
    for w0 in range(starts[b], ends[b] - 31, 32):
        hx, hy, hz = (a[w0:w0 + 32] for a in H)
        res.setdefault("plain walk, bank = z", []).append(wavefronts(walk(hz, z, z, z, 0, 0)))
        b0 = (11 * hx + hz) % 32
        r = rank_among_equal(b0)
        res.setdefault("x-planes rotated by rank, bank = z + 11 x (built)", []).append(wavefronts(walk(b0, r % 3, z, z, 11, 0)))
        b9 = (9 * hx + 3 * hy + hz) % 32
        r9 = rank_among_equal(b9)
        res.setdefault("x and y rotated by two digits of the rank, bank = z + 3 y + 9 x", []).append(
            wavefronts(walk(b9, r9 % 3, (r9 // 3) % 3, z, 9, 3)))
        rz = rank_among_equal(hz)
        res.setdefault("z rotated by rank, bank = z", []).append(wavefronts(walk(hz, z, z, rz % 3, 0, 0)))
        load = np.zeros((27, 32), int)                    # greedy: every lane in turn takes the cheapest of the 27 rotations
        for i in range(32):
            best = None
            for rr in range(27):
                bk = (b9[i] + 9 * ((A[0] + rr // 9) % 3) + 3 * ((B[0] + (rr // 3) % 3) % 3) + (C[0] + rr % 3) % 3) % 32
                cost = load[np.arange(27), bk].sum()
                if best is None or cost < best[0]:
                    best = (cost, bk)
            load[np.arange(27), best[1]] += 1
        res.setdefault("greedy choice among the 27 digit-wise rotations", []).append(load.max(axis=1).mean())
print(f"n = {n}, bricks {BX} x {BY} x {ZC}: lanes on the busiest bank per ATOMS, averaged over the 27 steps and the warps")
for k, v in res.items():
    print(f"  {np.mean(v):.3f}  {k}")
