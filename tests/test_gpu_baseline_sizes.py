"""Parity at the BASELINE sizes ("pin parity at the BASELINE sizes"):

* config 2 (512^3 Zel'dovich particles, CIC, 512^3 mesh) and the config-3 semantics at 512^3 (TSC + interlacing +
  compensation) are compared here, on the GPU box, with the oracle run on the box's host cores on the SAME particles:
  mode counts equal, <k> to 1e-12, P(k) to 1e-4 per bin;
* the committed fixtures tests/golden/c{2,3s,3,4s}_pk.npz (oracle output at 512^3 / 1024^3, made once by
  tools/make_fixtures.py) are compared with the CUDA path through the public API.

Through the C ABI like every -m gpu test; the oracle is the checker only.
"""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
PK_RTOL = 1e-4


def _gpu_power(ab, wl, pos, halos=None):
    N, L = wl["mesh"], wl["box"]
    kw = dict(resampler=wl["resampler"], interlaced=wl["interlaced"], compensated=wl["compensated"], normalize=True,
              pos_scale=1.0, device="cuda:0")
    matter = ab.CatalogMesh(pos, L, N, method="sorted", **kw)
    if halos is None:
        return ab.FFTPower(matter, mode="1d", kmin=2 * np.pi / L).power
    hal = ab.CatalogMesh(halos[:3], L, N, weight=halos[3], **kw)
    return ab.FFTPower(hal, mode="1d", second=matter, kmin=2 * np.pi / L).power


def _compare(p, k, P, modes):
    ok = np.isfinite(P) & (modes > 0)
    np.testing.assert_array_equal(p["modes"], modes)
    np.testing.assert_allclose(p["k"][ok], k[ok], rtol=1e-12)
    np.testing.assert_allclose(p["power"].real[ok], P[ok], rtol=PK_RTOL)
    return float(np.abs(p["power"].real[ok] / P[ok] - 1).max())


@pytest.mark.parametrize("key", ["c2", "c3s", "c4s"])
def test_512_cube_against_oracle_on_host_cores(oracle_fast, key):
    """CUDA path vs the oracle on the same 512^3 particle set (c2: CIC; c3s: TSC + interlaced + compensated; c4s: the
    halo x matter cross spectrum of config 4 at 512^3)."""
    import astrild_b200 as ab
    from bench import WORKLOADS, make_particles
    wl = WORKLOADS[key]
    dev = torch.device("cuda", 0)
    pos, halos = make_particles(wl, dev)
    got = _gpu_power(ab, wl, pos, halos)
    host = tuple(c.cpu().numpy() for c in pos)
    hh = None if halos is None else tuple(c.cpu().numpy() for c in halos)
    del pos, halos
    ab.engine.clear_engines()
    torch.cuda.empty_cache()
    threads = os.cpu_count() or 1
    kw = dict(resampler=wl["resampler"], interlaced=wl["interlaced"], compensated=wl["compensated"], normalize=True,
              workers=threads, threads=threads, paint_L=1.0, lean=True)
    if hh is None:
        k, P, modes = oracle_fast.power_from_particles(host, None, wl["mesh"], wl["box"], **kw)
    else:
        k, P, modes = oracle_fast.power_from_particles(hh[:3], hh[3], wl["mesh"], wl["box"], pos2=host, mass2=None, **kw)
    worst = _compare(got, k, P, modes)
    print(f"{key}: max |P/P_oracle - 1| = {worst:.2e} over {len(k)} bins")


@pytest.mark.parametrize("key", ["c2", "c3s", "c3", "c4s", "c4"])
def test_committed_fixture(key):
    """CUDA path vs the committed oracle fixture of the workload (tools/make_fixtures.py)."""
    path = os.path.join(GOLD, f"{key}_pk.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated yet (tools/make_fixtures.py {key})")
    import astrild_b200 as ab
    from bench import WORKLOADS, make_particles
    wl = WORKLOADS[key]
    need = 45e9 if wl["n"] >= 1024 else 8e9
    if torch.cuda.mem_get_info(0)[0] < need:
        pytest.skip("not enough free device memory for this workload")
    pos, halos = make_particles(wl, torch.device("cuda", 0))
    got = _gpu_power(ab, wl, pos, halos)
    del pos, halos
    ab.engine.clear_engines()
    torch.cuda.empty_cache()
    g = np.load(path)
    worst = _compare(got, g["k"], g["power"], g["modes"])
    print(f"{key}: max |P/P_fixture - 1| = {worst:.2e}")
