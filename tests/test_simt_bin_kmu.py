"""Row N4: the (k, mu) binning kernel's SOURCE on CPU fibers (tests/simt) against the oracle, and the oracle's
Hermitian half-space bookkeeping against a brute-force sum over all N^3 modes.

bin_kmu.cu's device code, unchanged except that its two inline-PTX RED helpers become plain adds: mode counts per
(k, mu) bin bit for bit, <k>, <mu>, P(k, mu) and the multipoles (even and odd ell), auto / cross, interlaced + compensated,
line of sight along z and oblique.
"""
import ctypes as ct
import os
import sys

import numpy as np
import pytest

from oracle import pk_oracle as o

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "simt"))


@pytest.fixture(scope="module")
def simt_kmu():
    import build_simt
    lib = ct.CDLL(build_simt.build_kmu())
    lib.simt_bin_kmu.restype = ct.c_int
    lib.simt_bin_kmu.argtypes = ([ct.c_void_p] * 4 + [ct.c_int] * 3 + [ct.c_void_p] * 5 + [ct.c_int, ct.c_int, ct.c_void_p, ct.c_int]
                                 + [ct.c_void_p] * 7 + [ct.c_int] * 3 + [ct.c_void_p] * 5)
    return lib


def bin_kmu(lib, N, L, c1, c1s=None, c2=None, c2s=None, kmin=0.0, Nmu=5, poles=(), los=(0.0, 0.0, 1.0), compensation=None):
    """-> dict like oracle.fftpower_2d, computed by the CPU-fiber kernels with the tables of the real library."""
    from astrild_b200 import tables
    kfull = tables.k_axis(N, L)
    edges = tables.k_edges(N, L, kmin)
    Nk = N // 2 + 1
    ka = kb = np.ascontiguousarray(kfull)
    kz = np.ascontiguousarray(kfull[:Nk])
    wz = np.ascontiguousarray(tables.hermitian_weights(N))
    comp, ph = [None] * 3, [None] * 3
    if compensation is not None:
        c = tables.compensation_axis(compensation[0], compensation[1], N)
        comp = [np.ascontiguousarray(c), np.ascontiguousarray(c), np.ascontiguousarray(c[:Nk])]
    if c1s is not None:
        p = tables.interlace_phase_axis(N, L)
        ph = [np.ascontiguousarray(p), np.ascontiguousarray(p), np.ascontiguousarray(p[:Nk])]
    ells = np.array(sorted(set([0] + [int(e) for e in poles])), dtype=np.int32)
    nb = (len(edges) + 1) * (Nmu + 2)
    xsum, musum = np.zeros(nb), np.zeros(nb)
    yre, yim = np.zeros((len(ells), nb)), np.zeros((len(ells), nb))
    nsum = np.zeros(nb, dtype=np.int64)
    losa = np.asarray(los, dtype=np.float64)

    def hp(a):
        return None if a is None else a.ctypes.data_as(ct.c_void_p)

    grids = [None if g is None else np.ascontiguousarray(g, dtype=np.complex64) for g in (c1, c1s, c2, c2s)]
    rc = lib.simt_bin_kmu(*[hp(g) for g in grids], N, N, Nk, hp(ka), hp(kb), hp(kz), hp(wz), hp(edges), len(edges), Nmu,
                          hp(ells), len(ells), hp(losa), hp(comp[0]), hp(comp[1]), hp(comp[2]), hp(ph[0]), hp(ph[1]), hp(ph[2]),
                          0, 0, 3, hp(xsum), hp(musum), hp(yre), hp(yim), hp(nsum))
    assert rc == 0
    shape = (len(edges) + 1, Nmu + 2)
    xsum, musum, nsum = xsum.reshape(shape), musum.reshape(shape), nsum.reshape(shape)
    ysum = (yre + 1j * yim).reshape((len(ells),) + shape) * L ** 3
    sl = slice(1, -1)
    with np.errstate(invalid="ignore", divide="ignore"):
        out = {"k": (xsum / nsum)[sl, sl], "mu": (musum / nsum)[sl, sl], "power": (ysum[0] / nsum)[sl, sl], "modes": nsum[sl, sl]}
        n1 = nsum[sl, sl].sum(axis=-1)
        out["poles"] = {"k": xsum[sl, sl].sum(axis=-1) / n1, "modes": n1}
        for e in poles:
            out["poles"]["power_%d" % e] = ysum[list(ells).index(int(e))][sl, sl].sum(axis=-1) / n1
    return out


def compare(got, want, poles, rtol=2e-5):
    np.testing.assert_array_equal(got["modes"], want["modes"])
    ok = want["modes"] > 0
    np.testing.assert_allclose(got["k"][ok], want["k"][ok], rtol=1e-12)
    np.testing.assert_allclose(got["mu"][ok], want["mu"][ok], rtol=1e-12, atol=1e-15)
    scale = np.abs(want["power"][ok]).max()
    np.testing.assert_allclose(got["power"][ok], want["power"][ok], rtol=rtol, atol=rtol * scale)
    np.testing.assert_array_equal(got["poles"]["modes"], want["poles"]["modes"])
    for e in poles:
        np.testing.assert_allclose(got["poles"]["power_%d" % e], want["poles"]["power_%d" % e], rtol=rtol, atol=rtol * scale)


@pytest.mark.parametrize("N", [8, 12, 15])
@pytest.mark.parametrize("los", [(0.0, 0.0, 1.0), (0.6, 0.0, 0.8)])
def test_oracle_half_space_rules_equal_the_sum_over_all_modes(N, los):
    """project_to_basis_2d's doubling rules (even ell: 2 Re, odd ell: 2i Im on non-singular modes) against every one of
    the N^3 modes of the full complex transform entering once: auto and cross, even and odd multipoles, N even and odd."""
    rng = np.random.default_rng(N)
    L, poles = 100.0, (0, 1, 2, 3, 4)
    d1 = rng.normal(size=(N, N, N))
    d2 = 0.5 * d1 + rng.normal(size=(N, N, N))
    for a, b in ((d1, None), (d1, d2)):
        r = o.fftpower_2d(o.r2c(a), None if b is None else o.r2c(b), N, L, Nmu=4, poles=poles, los=los, kmin=2 * np.pi / L)
        p2, n, pol = o.power2d_bruteforce(a, b, L, r["edges"], 4, poles=poles, los=los)
        np.testing.assert_array_equal(n, r["modes"])
        ok = n > 0
        scale = np.abs(p2[ok]).max()
        np.testing.assert_allclose(r["power"][ok].real, p2[ok].real, rtol=0, atol=1e-12 * scale)
        for e in poles:
            np.testing.assert_allclose(r["poles"]["power_%d" % e], pol[e], rtol=0, atol=1e-12 * scale)


@pytest.mark.parametrize("N,Nmu", [(16, 5), (20, 3), (33, 1), (24, 8)])
def test_kmu_mode_counts_and_power_match_the_oracle(simt_kmu, N, Nmu):
    rng = np.random.default_rng(100 + N)
    L, poles = 250.0, (0, 2, 4)
    c1 = o.r2c(rng.normal(1.0, 0.5, (N, N, N)))
    want = o.fftpower_2d(c1.copy(), None, N, L, Nmu=Nmu, poles=poles, kmin=2 * np.pi / L)
    got = bin_kmu(simt_kmu, N, L, c1, kmin=2 * np.pi / L, Nmu=Nmu, poles=poles)
    compare(got, want, poles)
    assert got["modes"].sum() == want["modes"].sum() > 0


def test_kmu_cross_oblique_line_of_sight_odd_poles(simt_kmu):
    rng = np.random.default_rng(7)
    N, L, poles, los = 18, 100.0, (0, 1, 2, 3), (0.36, 0.48, 0.8)
    a = rng.normal(size=(N, N, N))
    c1, c2 = o.r2c(a), o.r2c(0.3 * a + rng.normal(size=(N, N, N)))
    want = o.fftpower_2d(c1.copy(), c2.copy(), N, L, Nmu=6, poles=poles, los=los)
    got = bin_kmu(simt_kmu, N, L, c1, c2=c2, Nmu=6, poles=poles, los=los)
    compare(got, want, poles)


def test_kmu_interlaced_compensated(simt_kmu, oracle_fast):
    """Both interlaced meshes, TSC window deconvolution, on the combined field the oracle builds."""
    rng = np.random.default_rng(9)
    N, L, poles = 16, 64.0, (0, 2)
    pos = rng.random((4000, 3)) * L
    r0 = o.paint(pos, 1.0, N, L, "tsc")
    r1 = o.paint(pos, 1.0, N, L, "tsc", shift=0.5)
    s = N ** 3 / r0.sum()
    c0, c1 = o.r2c(r0) * s, o.r2c(r1) * s
    comb = o.compensate(o.interlace_combine(c0.copy(), c1.copy(), N, L), "tsc", True, N)
    want = o.fftpower_2d(comb, None, N, L, Nmu=4, poles=poles, kmin=2 * np.pi / L)
    got = bin_kmu(simt_kmu, N, L, c0, c1s=c1, kmin=2 * np.pi / L, Nmu=4, poles=poles, compensation=("tsc", True))
    compare(got, want, poles, rtol=5e-5)
