"""Row N2: catalog driver and table writer (astrild_b200/catalog.py) against the reference's layout
(/root/reference/src/astrild/particles/halo.py:157-207, 499-539; power_spectra/power_spectrum_3d.py:228-249)."""
import numpy as np
import pytest


def test_table_round_trip_keeps_the_reference_layout(tmp_path):
    """index = bins of the first snapshot, one column per snap_<n>; .npz stand-in when pytables is missing."""
    from astrild_b200 import catalog
    k = np.linspace(0.01, 1.0, 17)
    cols = {"snap_3": np.arange(17.0), "snap_11": np.arange(17.0) ** 2}
    path = catalog.write_table(str(tmp_path / "pk_rho.h5"), k, cols)
    assert path.endswith((".h5", ".npz"))
    idx, got = catalog.read_table(str(tmp_path / "pk_rho.h5"))
    np.testing.assert_array_equal(idx, k)
    assert list(got) == ["snap_3", "snap_11"]
    for c in cols:
        np.testing.assert_array_equal(got[c], cols[c])
    with pytest.raises(Exception):
        catalog.write_table(str(tmp_path / "bad.h5"), k, {"snap_1": np.zeros(3)})


def test_save_power_spectra_names_the_file_like_the_reference(tmp_path):
    from astrild_b200 import catalog
    pk = {"k": {"snap_1": np.array([1.0, 2.0]), "snap_2": np.array([1.0, 2.0])},
          "P": {"snap_1": np.array([5.0, 6.0]), "snap_2": np.array([7.0, 8.0])}}
    path = catalog.save_power_spectra(str(tmp_path) + "/", ["rho", "phi"], pk)
    assert path.rsplit(".", 1)[0].endswith("pk_rho_phi")
    idx, got = catalog.read_table(path)
    np.testing.assert_array_equal(idx, [1.0, 2.0])
    np.testing.assert_array_equal(got["snap_2"], [7.0, 8.0])


def test_subfind_stats_dispatches_by_name(tmp_path):
    """The statistic is looked up BY NAME and called with the YAML's args, per snapshot (halo.py:178,195-197)."""
    from astrild_b200 import catalog

    class FakeStats:
        calls = []

        def power_spectrum(snapshot, nbins=4, boxsize=1.0):
            FakeStats.calls.append((snapshot, nbins, boxsize))
            return np.arange(nbins, dtype=float), np.full(nbins, float(snapshot))

        def nothing(snapshot):
            return None, None

    stats = {"power_spectrum": {"args": {"nbins": 3, "boxsize": 2.0}, "resolution": 0}, "nothing": {"args": {}}}
    out = catalog.subfind_stats({5: 50, 7: 70, 9: None}, stats, stats_class=FakeStats, dir_out=str(tmp_path))
    assert FakeStats.calls == [(50, 3, 2.0), (70, 3, 2.0)]
    assert list(out["power_spectrum"]["results"]["values"]) == ["snap_5", "snap_7"]
    assert out["nothing"]["results"]["bins"] == {}
    idx, got = catalog.read_table(str(tmp_path / "subfind_power_spectrum_00.h5"))
    np.testing.assert_array_equal(idx, [0.0, 1.0, 2.0])
    np.testing.assert_array_equal(got["snap_7"], [70.0] * 3)


@pytest.mark.gpu
def test_batch_equals_one_run_per_snapshot(oracle_fast):
    """Three snapshots through one plan with deferred results == CatalogMesh -> FFTPower one at a time, and the oracle.
    Host columns small enough to force the chunked upload path (chunk_rows < Np), a device-resident set, a weighted set."""
    import torch
    import astrild_b200 as ab
    N, L = 48, 300.0
    rng = np.random.default_rng(11)
    sets = []
    for i in range(3):
        pos = (rng.random((60000 + 5000 * i, 3)) * L).astype(np.float32)
        sets.append(tuple(np.ascontiguousarray(pos[:, d]) for d in range(3)))
    w = np.exp(rng.normal(0, 1, sets[1][0].shape[0])).astype(np.float32) * 1e12
    batch = ab.PkBatch(N, L, resampler="tsc", interlaced=True, compensated=True, normalize=True, chunk_rows=16384, depth=2)
    dev_cols = tuple(torch.from_numpy(c).cuda() for c in sets[2])
    batch.submit(4, sets[0])
    batch.submit(9, lambda: (sets[1], w))
    batch.submit(12, dev_cols)
    batch.submit(13, sets[0], 2.5)                       # a scalar weight changes nothing once normalised
    got = batch.collect()
    assert list(got["k"]) == ["snap_4", "snap_9", "snap_12", "snap_13"]
    for key, cols, weight in (("snap_4", sets[0], None), ("snap_9", sets[1], w), ("snap_12", sets[2], None)):
        mesh = ab.CatalogMesh(cols, L, N, weight=weight, resampler="tsc", interlaced=True, compensated=True, normalize=True)
        r = ab.FFTPower(mesh, mode="1d", kmin=2 * np.pi / L)
        np.testing.assert_array_equal(got["modes"][key], r.power["modes"])
        np.testing.assert_allclose(got["k"][key], r.power["k"], rtol=1e-14)
        np.testing.assert_allclose(got["P"][key], r.power["power"].real - r.attrs["shotnoise"], rtol=1e-5, atol=1e-6 * abs(r.power["power"].real).max())
        want = oracle_fast.power_from_particles(np.stack(cols, axis=1), weight, N, L, resampler="tsc", interlaced=True,
                                                compensated=True, normalize=True)
        np.testing.assert_array_equal(got["modes"][key], want[2])
        np.testing.assert_allclose(got["P"][key] + got["shotnoise"][key], want[1], rtol=1e-4)
    np.testing.assert_allclose(got["P"]["snap_13"], got["P"]["snap_4"], rtol=1e-6, atol=1e-6 * abs(got["P"]["snap_4"]).max())
    assert batch.collect()["k"] == {}


@pytest.mark.gpu
def test_batch_unnormalised_density_matches_subfind_path(oracle_fast):
    """normalize=False, no interlacing / compensation: rho = mass / dx^3, what SubFind.power_spectrum feeds FFTPower."""
    import astrild_b200 as ab
    N, L = 32, 64.0
    rng = np.random.default_rng(5)
    pos = (rng.random((20000, 3)) * L).astype(np.float32)
    mass = (rng.random(20000) * 3 + 0.5).astype(np.float32)
    got = ab.PkBatch(N, L, resampler="tsc", interlaced=False, compensated=False, normalize=False).run([(1, pos, mass)])
    want = oracle_fast.power_from_particles(pos, mass, N, L, resampler="tsc", interlaced=False, compensated=False, normalize=False)
    np.testing.assert_array_equal(got["modes"]["snap_1"], want[2])
    np.testing.assert_allclose(got["P"]["snap_1"] + got["shotnoise"]["snap_1"], want[1], rtol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("normalize", [True, False])
def test_folded_box_matches_the_oracle_on_folded_positions(oracle_fast, normalize):
    """Row N3: fold = 2 puts x -> 4x mod L on a mesh that covers L / 4; equal to the oracle run on the explicitly folded
    positions in a box of L / 4, with the amplitude referred to the full volume (x 8^f) and, un-normalised, the density
    of the full box (x 8^-f on the field)."""
    import astrild_b200 as ab
    N, L, f = 32, 200.0, 2
    rng = np.random.default_rng(8)
    pos = (rng.random((40000, 3)) * L).astype(np.float32)
    mass = (rng.random(40000) + 0.5).astype(np.float32)
    mesh = ab.CatalogMesh(pos, L, N, weight=mass, resampler="tsc", interlaced=True, compensated=True, normalize=normalize, fold=f)
    r = ab.FFTPower(mesh, mode="1d", kmin=2 * np.pi / (L / 2 ** f))
    Lf = L / 2 ** f
    folded = np.mod(pos.astype(np.float64), Lf)
    want = oracle_fast.power_from_particles(folded, mass, N, Lf, resampler="tsc", interlaced=True, compensated=True,
                                            normalize=normalize)
    amp = 8.0 ** f if normalize else 8.0 ** f / 64.0 ** f
    np.testing.assert_array_equal(r.power["modes"], want[2])
    np.testing.assert_allclose(r.power["k"], want[0], rtol=1e-12)
    np.testing.assert_allclose(r.power["power"].real, want[1] * amp, rtol=2e-4)
    W, W2 = float(mass.astype(np.float64).sum()), float((mass.astype(np.float64) ** 2).sum())
    assert r.attrs["shotnoise"] == pytest.approx(L ** 3 * W2 / W ** 2 if normalize else W2 / L ** 3, rel=1e-12)


@pytest.mark.gpu
def test_power_spectrum_3d_compute_over_snapshots_and_saved_table(tmp_path, oracle_fast):
    """``PowerSpectrum3D.compute`` (power_spectrum_3d.py:33-81) over two snapshot files of gridded fields, saved as the
    reference's ``pk_<quantities>`` table and read back; values against the oracle's ArrayMesh -> FFTPower."""
    import astrild_b200 as ab
    from astrild_b200 import catalog
    N, L = 24, 150.0
    rng = np.random.default_rng(31)
    files = {}
    for nr in (3, 7):
        vm = rng.normal(2.0, 1.0, (N, N, N)).astype(np.float32)
        path = tmp_path / f"rho_{nr:03d}.npy"
        np.save(path, vm)
        files[nr] = (str(path), vm)

    class Sim:
        boxsize, domain_level, npar = L, N, N
        dir_nrs = [3, 7]
        dirs = {"out": str(tmp_path) + "/"}

        def get_file_nrs(self, dsc, where, kind):
            return [3, 7]

        def get_file_paths(self, dsc, where, kind):
            return [files[3][0], files[7][0]]

    ps = ab.PowerSpectrum3D("particles", Sim())
    pk = ps.compute(["rho"], [{"path": "x", "root": "rho", "extension": "npy"}], save=False)
    assert list(pk["P"]) == ["snap_3", "snap_7"]
    for nr in (3, 7):
        want = oracle_fast.power_from_mesh(files[nr][1].astype(np.float64), None, L)
        np.testing.assert_allclose(pk["k"]["snap_%d" % nr], want[0], rtol=1e-12)
        np.testing.assert_allclose(pk["P"]["snap_%d" % nr], want[1], rtol=1e-4)
    ps.compute(["rho"], [{"path": "x", "root": "rho", "extension": "npy"}], save=True)
    idx, cols = catalog.read_table(str(tmp_path / "pk_rho.h5"))
    np.testing.assert_array_equal(idx, pk["k"]["snap_3"])
    np.testing.assert_array_equal(cols["snap_7"], pk["P"]["snap_7"])
