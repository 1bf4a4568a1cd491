"""Regenerates tests/golden/*.  Run from the repo root: python tests/golden/make_golden.py

The reference's own arithmetic (nbodykit/pmesh/pfft) cannot be imported or built in this
image and astrild holds no golden vectors for this path, so:
  * mode_counts.json holds the vectors pinned by the survey's independent probe
     -- these pin the ORACLE;
  * pk_small.npz holds oracle outputs on seeded inputs -- these pin the CUDA path on the
    GPU box, where the oracle also runs live beside it.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pk_oracle as o  # noqa: E402

MODE_COUNTS = {  # SURVEY.md Appendix B.2, kmin = dk = 2pi/L, L = 1000, float64 k tables
    "8": {"modes": [26, 66, 158], "underflow": 1, "overflow": 261},
    "16": {"modes": [26, 66, 158, 234, 410, 470, 738], "underflow": 1, "overflow": 1993},
    "32": {"modes": [26, 66, 158, 234, 410, 470, 738, 866, 1170, 1364, 1620, 1970, 2366, 2624, 2988],
           "underflow": 1, "overflow": 15697},
    "128": {"first_bin": 26, "visible_total": 1097910},
}


def particles(seed, Np, L):
    rng = np.random.default_rng(seed)
    pos = rng.random((Np, 3)) * L
    mass = np.exp(rng.normal(0.0, 1.0, Np))
    return pos.astype(np.float32), mass.astype(np.float32)


CASES = {
    "tsc_mass_asis": dict(resampler="tsc", use_mass=True),
    "cic_unit": dict(resampler="cic", use_mass=False),
    "tsc_interlaced_compensated_normalized": dict(resampler="tsc", use_mass=False, interlaced=True,
                                                  compensated=True, normalize=True),
    "cic_compensated_normalized": dict(resampler="cic", use_mass=True, compensated=True, normalize=True),
}
N, L, NP, SEED = 32, 1000.0, 20000, 20261018

if __name__ == "__main__":
    with open(os.path.join(HERE, "mode_counts.json"), "w") as f:
        json.dump(MODE_COUNTS, f, indent=1)
    pos, mass = particles(SEED, NP, L)
    out = {}
    for name, kw in CASES.items():
        kw = dict(kw)
        m = mass if kw.pop("use_mass") else None
        k, pk, modes = o.power_from_particles(pos, m, N, L, **kw)
        out[name + "/k"], out[name + "/pk"], out[name + "/modes"] = k, pk, modes
    pos2, mass2 = particles(SEED + 1, NP // 4, L)
    k, pk, modes = o.power_from_particles(pos, None, N, L, resampler="tsc", normalize=True,
                                          pos2=pos2, mass2=mass2)
    out["cross_tsc/k"], out["cross_tsc/pk"], out["cross_tsc/modes"] = k, pk, modes
    rng = np.random.default_rng(SEED + 2)
    vm = rng.normal(5.0, 1.0, (N, N, N))
    vm2 = vm * 0.5 + rng.normal(0.0, 1.0, (N, N, N))
    k, pk, modes = o.power_from_mesh(vm, None, L)
    out["mesh_auto/k"], out["mesh_auto/pk"], out["mesh_auto/modes"] = k, pk, modes
    k, pk, modes = o.power_from_mesh(vm, vm2, L)
    out["mesh_cross/k"], out["mesh_cross/pk"], out["mesh_cross/modes"] = k, pk, modes
    np.savez_compressed(os.path.join(HERE, "pk_small.npz"), **out)
    print("wrote", sorted(out))
