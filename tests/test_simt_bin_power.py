"""The fused binning kernel's SOURCE on CPU fibers (tests/simt) against the oracle: mode counts bit for bit.

Same idea as tests/test_simt_deposit.py: bin_power.cu's device code, unchanged except that its two inline-PTX RED
helpers become plain adds, with the tables astrild_b200.tables builds for the real library.  Checks the
digitize fix-up, Hermitian weights, DC zeroing, the skip of lines beyond the last edge, the private shell windows
and their eviction, the interlacing combine, window compensation, cross spectra and the transposed-slab axes.
"""
import ctypes as ct
import json
import os
import sys

import numpy as np
import pytest

from oracle import pk_oracle as o

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "simt"))
GOLD = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def simt_bin():
    import build_simt
    lib = ct.CDLL(build_simt.build_bin())
    lib.simt_bin_power.restype = ct.c_int
    lib.simt_bin_power.argtypes = [ct.c_void_p] * 4 + [ct.c_int] * 3 + [ct.c_void_p] * 5 + [ct.c_int] + [ct.c_void_p] * 6 + [
        ct.c_int] * 3 + [ct.c_void_p] * 4
    return lib


def bin_power(lib, N, L, c1, c1s=None, c2=None, c2s=None, kmin=0.0, compensation=None, ia=None, ib=None, ctas=3):
    """Mirrors PkEngine.binning + bin_power_raw + finish (scale = L^3) with the CPU-fiber kernels."""
    from astrild_b200 import tables
    kfull = tables.k_axis(N, L)
    edges = tables.k_edges(N, L, kmin)
    Nk = N // 2 + 1
    ia = np.arange(N) if ia is None else ia
    ib = np.arange(N) if ib is None else ib
    ka, kb, kz = (np.ascontiguousarray(kfull[ia]), np.ascontiguousarray(kfull[ib]), np.ascontiguousarray(kfull[:Nk]))
    wz = np.ascontiguousarray(tables.hermitian_weights(N))
    dc_a = int(np.flatnonzero(ia == 0)[0]) if (ia == 0).any() else -1
    dc_b = int(np.flatnonzero(ib == 0)[0]) if (ib == 0).any() else -1
    comp = [None] * 3
    if compensation is not None:
        c = tables.compensation_axis(compensation[0], compensation[1], N)
        comp = [np.ascontiguousarray(c[ia]), np.ascontiguousarray(c[ib]), np.ascontiguousarray(c[:Nk])]
    ph = [None] * 3
    if c1s is not None:
        p = tables.interlace_phase_axis(N, L)
        ph = [np.ascontiguousarray(p[ia]), np.ascontiguousarray(p[ib]), np.ascontiguousarray(p[:Nk])]
    nb1 = len(edges) + 1
    ksum, pre, pim = np.zeros(nb1), np.zeros(nb1), np.zeros(nb1)
    nmodes = np.zeros(nb1, dtype=np.int64)

    def hp(a):
        return None if a is None else a.ctypes.data_as(ct.c_void_p)

    grids = [None if g is None else np.ascontiguousarray(g, dtype=np.complex64) for g in (c1, c1s, c2, c2s)]
    rc = lib.simt_bin_power(hp(grids[0]), hp(grids[1]), hp(grids[2]), hp(grids[3]), len(ka), len(kb), Nk, hp(ka), hp(kb),
                            hp(kz), hp(wz), hp(edges), len(edges), hp(comp[0]), hp(comp[1]), hp(comp[2]), hp(ph[0]),
                            hp(ph[1]), hp(ph[2]), dc_a, dc_b, ctas, hp(ksum), hp(pre), hp(pim), hp(nmodes))
    assert rc == 0
    with np.errstate(invalid="ignore", divide="ignore"):
        return {"edges": edges, "Nsum": nmodes, "modes": nmodes[1:-1], "k": (ksum / nmodes)[1:-1],
                "power": ((pre + 1j * pim) * L ** 3 / nmodes)[1:-1], "raw": (ksum, pre, pim)}


@pytest.mark.parametrize("N", [8, 16, 33, 40])
def test_mode_counts_bit_exact_on_cpu_fibers(simt_bin, N):
    L = 1000.0
    rng = np.random.default_rng(N)
    shape = (N, N, N // 2 + 1)
    c = (rng.normal(size=shape) + 1j * rng.normal(size=shape)).astype(np.complex64)
    for kmin in (2 * np.pi / L, 0.0):
        want = o.fftpower_1d(c.astype(np.complex128), None, N, L, kmin=kmin)
        got = bin_power(simt_bin, N, L, c, kmin=kmin)
        np.testing.assert_array_equal(got["edges"], want["edges"])
        np.testing.assert_array_equal(got["Nsum"], want["Nsum"])
        assert got["Nsum"].sum() == N ** 3
        np.testing.assert_allclose(got["k"], want["k"], rtol=1e-12)
        np.testing.assert_allclose(got["power"].real, want["power"].real, rtol=1e-6)


def test_pinned_mode_counts_on_cpu_fibers(simt_bin):
    pins = json.load(open(os.path.join(GOLD, "mode_counts.json")))
    L = 1000.0
    for N in (8, 16, 32):
        got = bin_power(simt_bin, N, L, np.zeros((N, N, N // 2 + 1), np.complex64), kmin=2 * np.pi / L)
        assert got["modes"].tolist() == pins[str(N)]["modes"]
        assert got["Nsum"][0] == pins[str(N)]["underflow"] and got["Nsum"][-1] == pins[str(N)]["overflow"]


def test_cross_interlaced_compensated_on_cpu_fibers(simt_bin):
    N, L = 24, 300.0
    rng = np.random.default_rng(1)
    shape = (N, N, N // 2 + 1)
    cs = [(rng.normal(size=shape) + 1j * rng.normal(size=shape)).astype(np.complex64) for _ in range(4)]
    c1, c1s, c2, c2s = [x.astype(np.complex128) for x in cs]
    a = o.compensate(o.interlace_combine(c1, c1s, N, L), "tsc", True, N)
    b = o.compensate(o.interlace_combine(c2, c2s, N, L), "tsc", True, N)
    want = o.fftpower_1d(a, b, N, L, kmin=2 * np.pi / L)
    got = bin_power(simt_bin, N, L, cs[0], cs[1], cs[2], cs[3], kmin=2 * np.pi / L, compensation=("tsc", True))
    np.testing.assert_array_equal(got["modes"], want["modes"])
    scale = np.abs(want["power"]).max()
    np.testing.assert_allclose(got["power"].real, want["power"].real, rtol=0, atol=2e-5 * scale)
    np.testing.assert_allclose(got["power"].imag, want["power"].imag, rtol=0, atol=2e-5 * scale)


def test_transposed_slabs_add_up_on_cpu_fibers(simt_bin):
    """The slab path bins [x][y_local][z] blocks with per-axis tables: the blocks' sums are the full grid's."""
    N, L, P = 16, 1000.0, 4
    rng = np.random.default_rng(9)
    shape = (N, N, N // 2 + 1)
    c = (rng.normal(size=shape) + 1j * rng.normal(size=shape)).astype(np.complex64)
    want = o.fftpower_1d(c.astype(np.complex128), None, N, L, kmin=2 * np.pi / L)
    nsum = 0
    pre = 0.0
    for r in range(P):
        ys = np.arange(r * N // P, (r + 1) * N // P)
        block = np.ascontiguousarray(c[:, ys, :])               # [x][y_local][z]: a = x, b = local y
        got = bin_power(simt_bin, N, L, block, kmin=2 * np.pi / L, ia=np.arange(N), ib=ys)
        nsum = nsum + got["Nsum"]
        pre = pre + got["raw"][1]
    np.testing.assert_array_equal(nsum, want["Nsum"])
    np.testing.assert_allclose((pre * L ** 3 / nsum)[1:-1], want["power"].real, rtol=1e-6)


def test_table_driven_variant_matches_the_one_pass_kernel(simt_bin):
    """Experimental APK_BIN_TABLE=1 path: a geometry pass stores the shell of every mode (and yields the mode counts
    and sum(w k)), the data pass reads that table instead of doing float64 wavenumber arithmetic.  Mode counts and
    sum(w P) must be bit-identical to the one-pass kernel, sum(w k) equal up to summation order."""
    rng = np.random.default_rng(5)
    for N, inter, cross, comp in [(24, True, False, True), (17, False, False, False), (20, True, True, True), (32, False, True, False)]:
        L = 300.0
        shape = (N, N, N // 2 + 1)
        grids = [(rng.normal(size=shape) + 1j * rng.normal(size=shape)).astype(np.complex64) for _ in range(4)]
        args = (grids[0], grids[1] if inter else None, grids[2] if cross else None, grids[3] if (cross and inter) else None)
        kw = dict(kmin=2 * np.pi / L, compensation=("tsc", inter) if comp else None)
        a = bin_power(simt_bin, N, L, *args, ctas=3, **kw)
        b = bin_power(simt_bin, N, L, *args, ctas=3 | (1 << 16), **kw)
        np.testing.assert_array_equal(a["Nsum"], b["Nsum"])
        np.testing.assert_array_equal(a["raw"][1], b["raw"][1])
        np.testing.assert_array_equal(a["raw"][2], b["raw"][2])
        np.testing.assert_allclose(a["raw"][0], b["raw"][0], rtol=1e-13)
