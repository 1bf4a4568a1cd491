"""Parity of the CUDA path (through the C ABI) against the oracle on a real GPU.

Bars (BASELINE.json north_star): mode counts and bin edges bit-exact; <k> to 1e-12; P(k)
within 1e-4 relative per bin (fp32 mesh + fp32 FFT vs the reference's float64).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import pk_oracle as o  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
PK_RTOL = 1e-4


@pytest.fixture(scope="module")
def ab():
    import astrild_b200
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return astrild_b200


def _particles(seed, Np, L, lo=0.0, hi=1.0):
    rng = np.random.default_rng(seed)
    pos = (lo + (hi - lo) * rng.random((Np, 3))) * L
    mass = np.exp(rng.normal(0.0, 1.0, Np))
    return pos.astype(np.float32), mass.astype(np.float32)


# ------------------------------------------------------------------------------ deposit
@pytest.mark.parametrize("method", ["atomic", "sorted"])
@pytest.mark.parametrize("resampler", ["cic", "tsc"])
@pytest.mark.parametrize("N", [32, 45, 64])
def test_deposit_matches_oracle(ab, oracle_fast, method, resampler, N):
    L = 1000.0
    pos, mass = _particles(7 + N, 50000, L, -0.2, 1.3)      # some particles outside the box: wrapped
    pm = ab.ParticleMesh(Nmesh=[N] * 3, BoxSize=L)
    for shift in (0.0, 0.5):
        for m in (None, mass):
            got = pm.paint(pos, mass=1.0 if m is None else m, resampler=resampler, shift=shift, method=method).value
            want = oracle_fast.paint(pos, m, N, L, resampler, shift)
            assert got.shape == (N, N, N)
            scale = np.abs(want).max()
            # float32 mesh + the sorted path's fixed-point tile: one deposit is rounded to 2^-18 .. 2^-21 of the chunk's
            # largest mass (2^-18 = 3.8e-6 for a full chunk of 8191 CIC particles, the case of the 32^3 mesh here)
            np.testing.assert_allclose(got, want, rtol=0, atol=5e-6 * scale)
            assert got.sum() == pytest.approx(want.sum(), rel=1e-6)


@pytest.mark.parametrize("method", ["atomic", "sorted"])
def test_deposit_layouts_and_dtypes_agree(ab, oracle_fast, method):
    N, L = 48, 250.0
    pos, mass = _particles(99, 30000, L)
    pm = ab.ParticleMesh(Nmesh=[N] * 3, BoxSize=L)
    want = oracle_fast.paint(pos, mass, N, L, "tsc")
    variants = {
        "aos_f32": (pos, mass),
        "aos_f64": (pos.astype(np.float64), mass.astype(np.float64)),
        "soa_f32": (tuple(np.ascontiguousarray(c) for c in pos.T), mass),
        "soa_f64": (tuple(np.ascontiguousarray(c.astype(np.float64)) for c in pos.T), mass),
        "torch_cuda": (torch.from_numpy(pos).cuda(), torch.from_numpy(mass).cuda()),
    }
    for name, (p, m) in variants.items():
        got = pm.paint(p, mass=m, resampler="tsc", method=method).value
        np.testing.assert_allclose(got, want, rtol=0, atol=5e-6 * want.max(), err_msg=name)


def test_deposit_known_weights(ab):
    N, L = 16, 1000.0
    p = np.array([[3.25, 7.0, 0.5], [N - 0.25, 0.0, 0.0]]) * L / N
    pm = ab.ParticleMesh(Nmesh=[N] * 3, BoxSize=L)
    for method in ("atomic", "sorted"):
        c = pm.paint(p[:1], resampler="cic", method=method).value
        assert c[3, 7, 0] == pytest.approx(0.375) and c[4, 7, 1] == pytest.approx(0.125)
        t = pm.paint(p[:1], resampler="tsc", method=method).value
        np.testing.assert_allclose(t.sum(axis=(1, 2))[2:5], [0.03125, 0.6875, 0.28125], rtol=1e-6)
        w = pm.paint(p[1:], resampler="cic", method=method).value
        assert w[0, 0, 0] == pytest.approx(0.75) and w[N - 1, 0, 0] == pytest.approx(0.25)


def test_deposit_empty_and_clustered(ab, oracle_fast):
    N, L = 32, 100.0
    pm = ab.ParticleMesh(Nmesh=[N] * 3, BoxSize=L)
    for method in ("atomic", "sorted"):
        z = pm.paint(np.zeros((0, 3), np.float32), resampler="tsc", method=method).value
        assert z.shape == (N, N, N) and not z.any()
    # everything in one cell plus a thin sheet: exercises multi-chunk bricks and empty bricks
    rng = np.random.default_rng(0)
    blob = (np.array([[10.3, 20.7, 5.1]]) + 0.4 * rng.random((40000, 3))) * L / N
    sheet = np.column_stack([rng.random(20000) * L, rng.random(20000) * L, np.full(20000, 17.49 * L / N)])
    pos = np.concatenate([blob, sheet]).astype(np.float32)
    for rs in ("cic", "tsc"):
        want = oracle_fast.paint(pos, None, N, L, rs)
        for method in ("atomic", "sorted"):
            got = pm.paint(pos, resampler=rs, method=method).value
            np.testing.assert_allclose(got, want, rtol=0, atol=2e-5 * want.max())  # 40000 fp32 adds into one cell


# ------------------------------------------------------------------------------ binning
@pytest.mark.parametrize("N", [8, 16, 32, 33, 64])
def test_bin_power_mode_counts_bit_exact(ab, N):
    L = 1000.0
    rng = np.random.default_rng(N)
    eng = ab.get_engine(N, L)
    c = (rng.normal(size=(N, N, N // 2 + 1)) + 1j * rng.normal(size=(N, N, N // 2 + 1))).astype(np.complex64)
    for kmin in (2 * np.pi / L, 0.0):
        want = o.fftpower_1d(c.astype(np.complex128), None, N, L, kmin=kmin)
        b = eng.binning(kmin=kmin)
        np.testing.assert_array_equal(b.edges, want["edges"])
        got = eng.bin_power(b, torch.from_numpy(c).cuda(), scale=L ** 3)
        np.testing.assert_array_equal(got["modes"], want["modes"])
        np.testing.assert_array_equal(got["Nsum"], want["Nsum"])
        assert got["Nsum"].sum() == N ** 3
        np.testing.assert_allclose(got["k"], want["k"], rtol=1e-12)
        np.testing.assert_allclose(got["power"].real, want["power"].real, rtol=1e-6)


def test_bin_power_pinned_mode_counts(ab):
    import json
    pins = json.load(open(os.path.join(GOLD, "mode_counts.json")))
    L = 1000.0
    for N in (8, 16, 32):
        eng = ab.get_engine(N, L)
        got = eng.bin_power(eng.binning(kmin=2 * np.pi / L),
                            torch.zeros((N, N, N // 2 + 1), dtype=torch.complex64, device="cuda"))
        assert got["modes"].tolist() == pins[str(N)]["modes"]
        assert got["Nsum"][0] == pins[str(N)]["underflow"] and got["Nsum"][-1] == pins[str(N)]["overflow"]
    N = 128
    eng = ab.get_engine(N, L)
    got = eng.bin_power(eng.binning(kmin=2 * np.pi / L),
                        torch.zeros((N, N, N // 2 + 1), dtype=torch.complex64, device="cuda"))
    assert got["modes"][0] == pins["128"]["first_bin"] and got["modes"].sum() == pins["128"]["visible_total"]


def test_bin_power_cross_interlaced_compensated(ab):
    N, L = 24, 300.0
    rng = np.random.default_rng(1)
    shape = (N, N, N // 2 + 1)
    cs = [(rng.normal(size=shape) + 1j * rng.normal(size=shape)).astype(np.complex64) for _ in range(4)]
    c1, c1s, c2, c2s = [x.astype(np.complex128) for x in cs]
    a = o.compensate(o.interlace_combine(c1, c1s, N, L), "tsc", True, N)
    b = o.compensate(o.interlace_combine(c2, c2s, N, L), "tsc", True, N)
    want = o.fftpower_1d(a, b, N, L, kmin=2 * np.pi / L)
    eng = ab.get_engine(N, L)
    bn = eng.binning(kmin=2 * np.pi / L, compensation=("tsc", True), interlaced=True)
    d = [torch.from_numpy(x).cuda() for x in cs]
    got = eng.bin_power(bn, d[0], d[1], d[2], d[3], scale=L ** 3)
    np.testing.assert_array_equal(got["modes"], want["modes"])
    scale = np.abs(want["power"]).max()
    np.testing.assert_allclose(got["power"].real, want["power"].real, rtol=0, atol=2e-5 * scale)
    np.testing.assert_allclose(got["power"].imag, want["power"].imag, rtol=0, atol=2e-5 * scale)


# ------------------------------------------------------------------------------ end to end
def test_golden_particle_cases(ab):
    import tests.golden.make_golden as g
    z = np.load(os.path.join(GOLD, "pk_small.npz"))
    pos, mass = g.particles(g.SEED, g.NP, g.L)
    for name, kw in g.CASES.items():
        kw = dict(kw)
        m = mass if kw.pop("use_mass") else None
        mesh = ab.CatalogMesh(pos, g.L, g.N, weight=m, resampler=kw["resampler"],
                              interlaced=kw.get("interlaced", False), compensated=kw.get("compensated", False),
                              normalize=kw.get("normalize", False))
        r = ab.FFTPower(mesh, mode="1d", kmin=2 * np.pi / g.L)
        np.testing.assert_array_equal(r.power["modes"], z[name + "/modes"], err_msg=name)
        np.testing.assert_allclose(r.power["k"], z[name + "/k"], rtol=1e-12, err_msg=name)
        np.testing.assert_allclose(r.power["power"].real, z[name + "/pk"], rtol=PK_RTOL, err_msg=name)
    pos2, mass2 = g.particles(g.SEED + 1, g.NP // 4, g.L)
    m1 = ab.CatalogMesh(pos, g.L, g.N, resampler="tsc", normalize=True)
    m2 = ab.CatalogMesh(pos2, g.L, g.N, weight=mass2, resampler="tsc", normalize=True)
    r = ab.FFTPower(m1, mode="1d", second=m2, kmin=2 * np.pi / g.L)
    want = z["cross_tsc/pk"]
    np.testing.assert_array_equal(r.power["modes"], z["cross_tsc/modes"])
    # a cross spectrum of independent catalogues scatters around 0: compare against the bin-to-bin scale
    np.testing.assert_allclose(r.power["power"].real, want, rtol=PK_RTOL, atol=PK_RTOL * np.abs(want).max())


def test_golden_mesh_cases_power_spectrum_3d(ab):
    import tests.golden.make_golden as g
    z = np.load(os.path.join(GOLD, "pk_small.npz"))
    rng = np.random.default_rng(g.SEED + 2)
    vm = rng.normal(5.0, 1.0, (g.N, g.N, g.N))
    vm2 = vm * 0.5 + rng.normal(0.0, 1.0, (g.N, g.N, g.N))

    class Sim:
        boxsize, domain_level, npar = g.L, g.N, g.N

    ps = ab.PowerSpectrum3D("particles", Sim())
    k, pk, modes = ps._power_spectrum_3d(vm, return_modes=True)
    np.testing.assert_array_equal(modes, z["mesh_auto/modes"])
    np.testing.assert_allclose(k, z["mesh_auto/k"], rtol=1e-12)
    np.testing.assert_allclose(pk, z["mesh_auto/pk"], rtol=PK_RTOL)
    k, pk = ps._power_spectrum_3d(vm, vm2)
    np.testing.assert_allclose(pk, z["mesh_cross/pk"], rtol=PK_RTOL)
    # float32 input (DTFE .npy maps are f4 in the reference)
    k, pk = ps._power_spectrum_3d(vm.astype(np.float32))
    np.testing.assert_allclose(pk, z["mesh_auto/pk"], rtol=PK_RTOL)


def test_config1_128_uniform_cic(ab, oracle_fast):
    """BASELINE config 1: 128^3 uniform-random particles, CIC on a 128^3 mesh."""
    N, L = 128, 1000.0
    rng = np.random.default_rng(12345)
    pos = (rng.random((N ** 3, 3)) * L).astype(np.float32)
    want = oracle_fast.power_from_particles(pos, None, N, L, resampler="cic", normalize=True, threads=4, workers=4)
    for method in ("sorted", "atomic"):
        mesh = ab.CatalogMesh(pos, L, N, resampler="cic", normalize=True, method=method)
        r = ab.FFTPower(mesh, mode="1d", kmin=2 * np.pi / L)
        np.testing.assert_array_equal(r.power["modes"], want[2])
        np.testing.assert_allclose(r.power["k"], want[0], rtol=1e-12)
        np.testing.assert_allclose(r.power["power"].real, want[1], rtol=PK_RTOL)


def test_subfind_power_spectrum_dropin(ab, oracle_fast):
    """SubFind.power_spectrum with the reference's keywords on a fake snapshot."""
    nbins, boxsize, h = 64, 500.0, 0.7
    rng = np.random.default_rng(2)
    n = 200000

    class Header:
        hubble = h
        boxsize = 500.0e3

    class Snap:
        header = Header()
        cat = {"SubhaloPos": (rng.random((n, 3)) * boxsize * 1e3 / h).astype(np.float32),
               "SubhaloMass": np.exp(rng.normal(2.0, 1.0, n)).astype(np.float32)}

    k, pk, modes = ab.SubFind.power_spectrum(Snap(), objects="subhalo", nbins=nbins, boxsize=boxsize, return_modes=True)
    pos = Snap.cat["SubhaloPos"][:] * h / 1e3
    mass = Snap.cat["SubhaloMass"][:] * h / 1e10
    wk, wpk, wmodes = oracle_fast.power_from_particles(pos, mass, nbins, boxsize, resampler="tsc")
    np.testing.assert_array_equal(modes, wmodes)
    np.testing.assert_allclose(k, wk, rtol=1e-12)
    np.testing.assert_allclose(pk, wpk, rtol=PK_RTOL)
    assert len(k) == nbins // 2 - 1


def test_arraymesh_kwargs_are_inert(ab):
    """compensated / interlaced / window on ArrayMesh change nothing."""
    N, L = 32, 100.0
    vm = np.random.default_rng(4).normal(size=(N, N, N))
    a = ab.FFTPower(ab.ArrayMesh(vm, BoxSize=L, Nmesh=N, compensated=False), mode="1d", kmin=2 * np.pi / L)
    b = ab.FFTPower(ab.ArrayMesh(vm, BoxSize=L, Nmesh=N, compensated=True, interlaced=True, window="TSC"),
                    mode="1d", kmin=2 * np.pi / L)
    np.testing.assert_array_equal(a.power["power"], b.power["power"])
    assert a.power.attrs["shotnoise"] == 0.0


def test_plane_wave_and_linearity(ab):
    N, L, A = 64, 500.0, 0.3
    m = np.array([3, 0, 2])
    x = np.arange(N) * L / N
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    delta = A * np.cos(2 * np.pi * (m[0] * X + m[1] * Y + m[2] * Z) / L)
    r = ab.FFTPower(ab.ArrayMesh(delta, BoxSize=L), mode="1d", kmin=2 * np.pi / L)
    b = int(np.floor(np.sqrt((m ** 2).sum()))) - 1
    expect = 2 * (A ** 2 * L ** 3 / 4) / r.power["modes"][b]
    assert r.power["power"].real[b] == pytest.approx(expect, rel=1e-5)
    others = np.delete(r.power["power"].real, b)
    assert np.abs(others).max() < 1e-9 * expect
    r4 = ab.FFTPower(ab.ArrayMesh(2 * delta, BoxSize=L), mode="1d", kmin=2 * np.pi / L)   # P scales as amplitude^2
    assert r4.power["power"].real[b] == pytest.approx(4 * expect, rel=1e-5)


def test_error_paths(ab):
    from astrild_b200._lib import AstrildPkError
    with pytest.raises(AstrildPkError):
        ab.ArrayMesh(np.zeros((8, 8, 4)), BoxSize=10.0)
    with pytest.raises(AstrildPkError):
        ab.ParticleMesh(Nmesh=[16] * 3, BoxSize=10.0).paint(np.zeros((5, 2)))
    with pytest.raises(AstrildPkError):
        ab.FFTPower(ab.ArrayMesh(np.zeros((8, 8, 8)), BoxSize=10.0), mode="3d")


def test_streamed_host_deposit_matches_device_deposit(ab, oracle_fast):
    """Chunked upload + deposit (host inputs) gives the same meshes as one deposit of device arrays."""
    N, L = 64, 1000.0
    pos, mass = _particles(5, 300000, L)
    eng = ab.get_engine(N, L)
    want = [oracle_fast.paint(pos, mass, N, L, "tsc", sh) for sh in (0.0, 0.5)]
    variants = {
        "aos_pageable": (pos, mass),
        "soa_pinned": (tuple(torch.from_numpy(np.ascontiguousarray(c)).pin_memory() for c in pos.T),
                       torch.from_numpy(mass).pin_memory()),
        "device": (torch.from_numpy(pos).cuda(), torch.from_numpy(mass).cuda()),
    }
    for name, (p, m) in variants.items():
        meshes = eng.deposit_many(p, m, "tsc", (0.0, 0.5), method="sorted", chunk_rows=70000)
        for mesh, w in zip(meshes, want):
            got = eng.store_mesh(mesh).cpu().numpy()
            np.testing.assert_allclose(got, w, rtol=0, atol=3e-6 * w.max(), err_msg=name)
    # unit masses through the public CatalogMesh path with a forced small chunk
    r1 = ab.FFTPower(ab.CatalogMesh(pos, L, N, resampler="cic", normalize=True), mode="1d", kmin=2 * np.pi / L)
    r2 = ab.FFTPower(ab.CatalogMesh(torch.from_numpy(pos).cuda(), L, N, resampler="cic", normalize=True), mode="1d",
                     kmin=2 * np.pi / L)
    np.testing.assert_allclose(r1.power["power"].real, r2.power["power"].real, rtol=1e-6)


@pytest.mark.parametrize("resampler", ["cic", "tsc"])
@pytest.mark.parametrize("method", ["sorted", "atomic"])
def test_interlaced_pair_deposit_matches_two_deposits(ab, oracle_fast, resampler, method):
    """apk_deposit_interlaced (one shared partition) == two independent deposits == oracle."""
    N, L = 50, 777.0
    pos, mass = _particles(21, 400000, L, -0.1, 1.1)
    eng = ab.get_engine(N, L)
    for m in (None, mass):
        pair = eng.deposit_pair(pos, m, resampler, method=method)
        for mesh, sh in zip(pair, (0.0, 0.5)):
            want = oracle_fast.paint(pos, m, N, L, resampler, sh)
            got = eng.store_mesh(mesh).cpu().numpy()
            np.testing.assert_allclose(got, want, rtol=0, atol=3e-6 * want.max())
            assert got.sum() == pytest.approx(want.sum(), rel=1e-6)


def test_2048_mesh_64bit_indexing(ab):
    """Size-independent properties at BASELINE's largest mesh: 2048^3 cells exceed 2^32 flat indices."""
    free, _ = torch.cuda.mem_get_info()
    if free < 90e9:
        pytest.skip("needs ~80 GB of device memory")
    from astrild_b200 import engine as _engine
    N, L = 2048, 1000.0
    eng = ab.PkEngine(N, L)
    try:
        # particles in the far corner: flat cell index ~ N^3 - 1 > 2^32
        g = np.array([[N - 1.25, N - 1.5, N - 1.75], [0.25, 0.5, 0.75], [N / 2 + 0.5, N - 0.75, 3.25]])
        pos = (g * L / N).astype(np.float64)
        mass = np.array([2.0, 3.0, 5.0])
        for method in ("atomic", "sorted"):
            mesh = eng.deposit(pos, mass, "tsc", method=method)
            assert eng.mesh_sum(mesh) == pytest.approx(10.0, rel=1e-6)
            m = mesh.view(N, N, eng.ldz)
            # TSC: the home cell of particle 0 is (N-1, N-1 or N-2.., ...): check one known weight product
            w = lambda d: 0.75 - d * d                       # |d| <= 0.5
            # particle 0: g = (N-1.25, N-1.5, N-1.75): home cells (N-1, N-1 [ties round up: floor(g+.5)=N-1], N-2)
            hx, hy, hz = N - 1, N - 1, N - 2
            dx, dy, dz = (N - 1.25) - hx, (N - 1.5) - hy, (N - 1.75) - hz
            want = 2.0 * w(dx) * w(dy) * w(dz)
            assert float(m[hx, hy, hz]) == pytest.approx(want, rel=1e-5)
            del mesh, m
        # mode counts of the whole 2048^3 lattice: every mode lands in exactly one bin
        grid = torch.zeros((N, N, N // 2 + 1), dtype=torch.complex64, device="cuda")
        res = eng.bin_power(eng.binning(kmin=2 * np.pi / L), grid)
        assert int(res["Nsum"].sum()) == N ** 3
        assert res["modes"][0] == 26 and len(res["modes"]) == N // 2 - 1
        # r2c of a single plane wave at 2048^3 (cuFFT 64-bit plan): power only in its shell
        del grid
        mesh = eng.new_mesh()
        x = torch.arange(N, device="cuda", dtype=torch.float32)
        mesh.zero_()
        mesh.view(N, N, eng.ldz)[:, :, :N] += torch.cos(2 * np.pi * 5 * x / N)[None, None, :]
        c = eng.r2c(mesh)
        res = eng.bin_power(eng.binning(kmin=2 * np.pi / L), c, scale=L ** 3 / float(N) ** 6)
        b = 5 - 1
        expect = 2 * (L ** 3 / 4) / res["modes"][b]
        assert res["power"].real[b] == pytest.approx(expect, rel=1e-4)
        assert np.abs(np.delete(res["power"].real, b)).max() < 1e-6 * expect
    finally:
        eng.close()
        torch.cuda.empty_cache()


@pytest.mark.parametrize("cross", [False, True])
def test_fftpower_2d_wedges_and_multipoles(ab, oracle_fast, cross):
    """Row N4: FFTPower(mode='2d', Nmu=, poles=, los=) through the C ABI against the oracle's project_to_basis: mode
    counts per (k, mu) bin equal, <k>, <mu>, P(k, mu) and P_0, P_2, P_4 within the fp32-mesh tolerance."""
    from oracle import pk_oracle as o
    N, L = 48, 300.0
    rng = np.random.default_rng(77)
    pos = (rng.random((150000, 3)) * L).astype(np.float32)
    pos2 = np.mod(pos[:60000] + rng.normal(0, 2.0, (60000, 3)).astype(np.float32), L).astype(np.float32) if cross else None
    kw = dict(resampler="tsc", interlaced=True, compensated=True, normalize=True)
    m1 = ab.CatalogMesh(pos, L, N, **kw)
    m2 = ab.CatalogMesh(pos2, L, N, **kw) if cross else None
    r = ab.FFTPower(m1, mode="2d", Nmu=5, poles=[0, 2, 4], second=m2, kmin=2 * np.pi / L, los=[0, 0, 1])

    def field(p):
        real, real2 = o.paint(p, 1.0, N, L, "tsc"), o.paint(p, 1.0, N, L, "tsc", shift=0.5)
        s = N ** 3 / real.sum()
        return o.compensate(o.interlace_combine(o.r2c(real) * s, o.r2c(real2) * s, N, L), "tsc", True, N)

    want = o.fftpower_2d(field(pos), field(pos2) if cross else None, N, L, Nmu=5, poles=(0, 2, 4), kmin=2 * np.pi / L)
    assert r.power["modes"].shape == (len(want["edges"]) - 1, 5)
    np.testing.assert_array_equal(r.power["modes"], want["modes"])
    ok = want["modes"] > 0
    np.testing.assert_allclose(r.power["k"][ok], want["k"][ok], rtol=1e-12)
    np.testing.assert_allclose(r.power["mu"][ok], want["mu"][ok], rtol=1e-12, atol=1e-15)
    scale = np.abs(want["power"][ok]).max()
    np.testing.assert_allclose(r.power["power"][ok].real, want["power"][ok].real, rtol=PK_RTOL, atol=PK_RTOL * scale)
    np.testing.assert_array_equal(r.poles["modes"], want["poles"]["modes"])
    for ell in (0, 2, 4):
        np.testing.assert_allclose(r.poles["power_%d" % ell].real, want["poles"]["power_%d" % ell].real, rtol=PK_RTOL,
                                   atol=PK_RTOL * scale)
    # the monopole of the wedges is the 1-D spectrum
    r1 = ab.FFTPower(m1, mode="1d", second=m2, kmin=2 * np.pi / L)
    np.testing.assert_array_equal(r1.power["modes"], r.poles["modes"])
    np.testing.assert_allclose(r.poles["power_0"].real, r1.power["power"].real, rtol=1e-6, atol=1e-6 * scale)
