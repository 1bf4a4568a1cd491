"""Cell-occupancy statistics of a synthetic workload as the tile kernel sees them (CPU, tools -- not product).

usage: occupancy_stats.py [n_side=128] [seed=31337]
Prints, for the Zel'dovich particle set of bench.py at n_side^3: the fraction of empty cells, the mean number of
iterations of the per-cell particle loop per column of 32 lanes (= max count over the lanes), the lanes active in
iteration k, and what capping the loop at K iterations would leave over.  These numbers are why the per-cell loop is the
floor of brick_deposit_kernel.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from astrild_b200 import synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 31337
x, y, z = [c.numpy().astype(np.float64) for c in synthetic.zeldovich_particles(n, 1000.0, seed, "cpu")]
for shift in (0.0, 0.5):
    h = [np.floor(c * n + shift + 0.5).astype(np.int64) % n for c in (x, y, z)]       # TSC home cell
    cnt = np.zeros((n, n, n), np.int32)
    np.add.at(cnt, tuple(h), 1)
    cols = cnt.reshape(n, n, n // 32, 32)
    mx = cols.max(axis=3)
    print(f"shift {shift}: empty cells {np.mean(cnt == 0):.3f} (Poisson(1): 0.368); loop iterations per column "
          f"{mx.mean():.2f}; lane efficiency {cnt.sum() / (mx.sum() * 32):.3f}")
    print("  lanes active in iteration k:", " ".join(f"{(cols > k).sum(axis=3)[mx > k].mean():.1f}" for k in range(8)))
    for K in (2, 3, 4):
        print(f"  loop capped at {K}: {np.minimum(mx, K).sum() / mx.sum():.2f} of the iterations remain, "
              f"{np.maximum(cnt - K, 0).sum() / cnt.sum():.3f} of the particles overflow")
    pair = (cnt[0::2] + cnt[1::2]).reshape(n // 2, n, n // 32, 32).max(axis=3)
    print(f"  x-pairs of cells per lane: {pair.mean():.2f} iterations per 2 columns (vs {2 * mx.mean():.2f})")
