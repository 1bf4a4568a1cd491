"""Pins the oracle: known-answer tests and the golden vectors.

The reference holds no test, fixture or golden vector for this path, and
its arithmetic lives in un-vendored nbodykit/pmesh, so these analytic answers plus the mode
counts pinned by the survey's independent probe are what the oracle is anchored on.
"""
import json
import os

import numpy as np
import pytest

from oracle import pk_oracle as o

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_deposit_weights_cic_tsc():
    N, L = 16, 1000.0
    p = np.array([[3.25, 7.0, 0.5]]) * L / N
    c = o.paint(p, 1.0, N, L, "cic")
    assert c[3, 7, 0] == pytest.approx(0.375) and c[4, 7, 0] == pytest.approx(0.125)
    assert c[3, 7, 1] == pytest.approx(0.375) and c[4, 7, 1] == pytest.approx(0.125)
    assert c[3, 8, 0] == 0.0 and c.sum() == pytest.approx(1.0, abs=1e-15)
    t = o.paint(p, 1.0, N, L, "tsc")
    line = t.sum(axis=(1, 2))
    np.testing.assert_allclose(line[2:5], [0.03125, 0.6875, 0.28125], rtol=1e-14)
    assert t.sum() == pytest.approx(1.0, abs=1e-15)


def test_deposit_wraps_and_conserves_mass():
    N, L = 8, 100.0
    p = np.array([[(N - 0.25) * L / N, 0.0, 0.0], [-0.3 * L / N, L * 2.5, 1e-9]])
    m = np.array([2.0, 3.0])
    for rs in ("cic", "tsc", "nearest"):
        c = o.paint(p, m, N, L, rs)
        assert c.sum() == pytest.approx(5.0, rel=1e-14)
    c = o.paint(p[:1], 1.0, N, L, "cic")
    assert c[0, 0, 0] == pytest.approx(0.75) and c[N - 1, 0, 0] == pytest.approx(0.25)


def test_c_twin_matches_numpy(oracle_fast):
    rng = np.random.default_rng(3)
    N, L, Np = 24, 250.0, 5000
    pos = rng.random((Np, 3)) * L * 1.2 - 0.1 * L      # some outside the box
    mass = rng.random(Np) + 0.5
    for rs in ("nearest", "cic", "tsc"):
        for sh in (0.0, 0.5):
            a = o.paint(pos, mass, N, L, rs, sh)
            b = oracle_fast.paint(pos, mass, N, L, rs, sh)
            np.testing.assert_allclose(b, a, rtol=0, atol=1e-13)
    kw = dict(resampler="tsc", interlaced=True, compensated=True, normalize=True)
    r1 = o.power_from_particles(pos, mass, N, L, **kw)
    r2 = oracle_fast.power_from_particles(pos, mass, N, L, threads=2, **kw)
    np.testing.assert_array_equal(r1[2], r2[2])
    np.testing.assert_allclose(r2[0], r1[0], rtol=1e-14)
    np.testing.assert_allclose(r2[1], r1[1], rtol=1e-12)


@pytest.mark.parametrize("resampler", ["cic", "tsc"])
@pytest.mark.parametrize("threads", [2, 5, 16])
def test_threaded_paint_equals_scalar_loop(oracle_fast, resampler, threads):
    """The CPU arm's threaded deposit (y-row blocks, two phases) against the scalar loop: same canvas up to the
    order of float64 additions, mass conserved, also with out-of-box positions, a shift and N not divisible by the
    number of blocks."""
    rng = np.random.default_rng(11)
    L = 250.0
    pos = (rng.random((60000, 3)) * 1.3 * L - 0.15 * L).astype(np.float32)
    mass = rng.random(60000)
    for N in (37, 64):
        a = oracle_fast.paint(pos, mass, N, L, resampler, shift=0.5)
        b = oracle_fast.paint(pos, mass, N, L, resampler, shift=0.5, threads=threads)
        np.testing.assert_allclose(b, a, rtol=1e-12, atol=1e-12 * a.max())
        assert abs(b.sum() - mass.sum()) < 1e-9 * mass.sum()


@pytest.mark.parametrize("N", [8, 16, 32])
def test_mode_counts_pinned(N):
    pins = json.load(open(os.path.join(GOLD, "mode_counts.json")))[str(N)]
    L = 1000.0
    edges = o.k_edges(N, L, kmin=2 * np.pi / L)
    assert len(edges) == N // 2
    _, _, nsum = o.project_to_basis_1d(np.zeros((N, N, N // 2 + 1), complex), N, L, edges)
    assert nsum[1:-1].tolist() == pins["modes"]
    assert nsum[0] == pins["underflow"] and nsum[-1] == pins["overflow"]


@pytest.mark.parametrize("N", [8, 16, 32, 33])
def test_mode_counts_bruteforce_full_lattice(N):
    L = 1000.0
    edges = o.k_edges(N, L, kmin=2 * np.pi / L)
    _, _, nsum = o.project_to_basis_1d(np.zeros((N, N, N // 2 + 1), complex), N, L, edges)
    np.testing.assert_array_equal(nsum, o.mode_counts_bruteforce(N, L, edges))
    assert nsum.sum() == N ** 3


def test_mode_counts_128_sanity(oracle_fast):
    pins = json.load(open(os.path.join(GOLD, "mode_counts.json")))["128"]
    N, L = 128, 1000.0
    r = oracle_fast.fftpower_1d(np.zeros((N, N, N // 2 + 1), complex), None, N, L, kmin=2 * np.pi / L)
    assert len(r["modes"]) == 63
    assert r["modes"][0] == pins["first_bin"] and r["modes"].sum() == pins["visible_total"]


def test_plane_wave():
    N, L, A = 32, 500.0, 0.3
    m = np.array([3, 0, 2])
    x = (np.arange(N) + 0.0) * L / N
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    delta = A * np.cos(2 * np.pi * (m[0] * X + m[1] * Y + m[2] * Z) / L)
    k, pk, modes = o.power_from_mesh(delta, None, L)
    kf = 2 * np.pi / L
    b = int(np.floor(np.sqrt((m ** 2).sum()))) - 1          # edges start at kf
    expect = 2 * (A ** 2 * L ** 3 / 4) / modes[b]
    assert pk[b] == pytest.approx(expect, rel=1e-12)
    mask = np.ones(len(pk), bool)
    mask[b] = False
    assert np.abs(pk[mask]).max() < 1e-20 * expect + 1e-12
    assert k[b] > b * kf and k[b] < (b + 2) * kf


def test_shot_noise_levels():
    rng = np.random.default_rng(11)
    N, L = 32, 1000.0
    Np = 4 * N ** 3
    pos = rng.random((Np, 3)) * L
    V = L ** 3
    k, pk, modes = o.power_from_particles(pos, None, N, L, "tsc", interlaced=True, compensated=True, normalize=True)
    ratio = np.average(pk / (V / Np), weights=modes)
    assert ratio == pytest.approx(1.0, abs=0.02)
    assert abs(pk[-4:].mean() / (V / Np) - 1) < 0.05           # flat up to Nyquist
    k, pk, modes = o.power_from_particles(pos, None, N, L, "cic", compensated=True, normalize=True)
    assert np.average(pk / (V / Np), weights=modes) == pytest.approx(1.0, abs=0.02)
    k, pk, modes = o.power_from_particles(pos, None, N, L, "tsc", normalize=True)
    assert pk[:3].mean() / (V / Np) == pytest.approx(1.0, abs=0.15)
    assert pk[-1] < 0.5 * V / Np                               # uncompensated window suppresses


def test_astrild_as_written_amplitude():
    """rho = paint/dx^3 is not normalised: P scales as rhobar^2."""
    rng = np.random.default_rng(5)
    N, L, Np = 16, 200.0, 3000
    pos = rng.random((Np, 3)) * L
    mass = rng.random(Np) + 1.0
    k1, p1, m1 = o.power_from_particles(pos, mass, N, L, "tsc")
    k2, p2, m2 = o.power_from_particles(pos, mass, N, L, "tsc", normalize=True)
    rhobar = mass.sum() / L ** 3
    np.testing.assert_allclose(p1, p2 * rhobar ** 2, rtol=1e-12)
    np.testing.assert_array_equal(m1, m2)


def test_golden_file_matches_oracle():
    import tests.golden.make_golden as g
    z = np.load(os.path.join(GOLD, "pk_small.npz"))
    pos, mass = g.particles(g.SEED, g.NP, g.L)
    for name, kw in g.CASES.items():
        kw = dict(kw)
        m = mass if kw.pop("use_mass") else None
        k, pk, modes = o.power_from_particles(pos, m, g.N, g.L, **kw)
        np.testing.assert_array_equal(modes, z[name + "/modes"])
        np.testing.assert_allclose(k, z[name + "/k"], rtol=1e-13)
        np.testing.assert_allclose(pk, z[name + "/pk"], rtol=1e-10)
