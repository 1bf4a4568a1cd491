"""N1 ingest: the host-side record walk against the reference's unpack loop (CPU), and the device kernels against the
NumPy restatements (GPU)."""
import numpy as np
import pytest
import torch

from oracle import ingest_oracle as io


def _blocks(rng, ncpu, nboundary, levels, nfields, ndim=3):
    blocks = {}
    for lev in levels:
        for ib in range(1, nboundary + ncpu + 1):
            ncache = int(rng.integers(0, 40))
            if ncache and rng.random() < 0.8:
                blocks[(lev, ib)] = rng.random((2 ** ndim, nfields, ncache))
    return blocks


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_record_walk_matches_the_reference_loop(seed):
    """poisson_record_pieces (offsets only) reproduces ecosmog.py:184-230's unpack loop value for value."""
    from astrild_b200.ingest import poisson_record_pieces
    rng = np.random.default_rng(seed)
    nfields, levels = 4, (7, 8)
    img = io.write_poisson(_blocks(rng, 3, 2, levels, nfields), 3, 2, min(levels), max(levels))
    want = io.unpack_poisson(img, nfields, min(levels), max(levels))
    pieces, counts = poisson_record_pieces(img, nfields, min(levels), max(levels))
    raw = np.frombuffer(img, dtype=np.uint8)
    for j in range(nfields):
        assert counts[j] == len(want[j])
        got = np.empty(counts[j])
        for src, dst, cnt in pieces[j]:
            got[dst:dst + cnt] = raw[src:src + 8 * cnt].view(np.float64) if src % 8 == 0 else \
                np.frombuffer(raw[src:src + 8 * cnt].tobytes(), dtype=np.float64)
        np.testing.assert_array_equal(got, want[j])


def test_truncated_image_is_an_error():
    from astrild_b200.ingest import poisson_record_pieces
    from astrild_b200._lib import AstrildPkError
    rng = np.random.default_rng(5)
    img = io.write_poisson(_blocks(rng, 2, 1, (7,), 3), 2, 1, 7, 7)
    with pytest.raises(AstrildPkError):
        poisson_record_pieces(img[:-20], 3, 7, 7)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_assign_grid_matches_numpy_fancy_assignment(dtype):
    """apk_assign_grid == value_map[(x, y, z)] = values: truncation toward zero, negative wrap, last write wins."""
    from astrild_b200 import ingest
    rng = np.random.default_rng(11)
    N, n = 48, 200000                                        # ~2 samples per cell: plenty of repeated cells
    x, y, z = (rng.random(n).astype(dtype) for _ in range(3))
    x[:50] = -x[:50] * 0.5                                    # negative coordinates: index -k wraps to N - k
    x[50:60] = 0.0
    z[60:70] = np.nextafter(dtype(1.0), dtype(0.0))
    vals = rng.normal(size=n)
    want = io.read_data_assign(N, x, y, z, vals)
    got = ingest.assign_grid(x, y, z, vals, N, device="cuda:0").cpu().numpy()
    np.testing.assert_array_equal(got, want)
    with pytest.raises(IndexError):
        bad = x.copy(); bad[7] = 1.5
        ingest.assign_grid(bad, y, z, vals, N, device="cuda:0")


@pytest.mark.gpu
def test_poisson_reader_and_gridder_end_to_end():
    """Two cpu files of a snapshot -> device columns (apk_gather_records) -> gridded field (apk_assign_grid) -> P(k),
    against the reference's unpack loop + NumPy assignment + the oracle's power_from_mesh."""
    import astrild_b200 as ab
    from astrild_b200 import ingest
    from oracle import pk_oracle_fast as f
    rng = np.random.default_rng(3)
    N, fields, levels = 16, ["x", "y", "z", "phi"], (4,)
    cells = (np.stack(np.meshgrid(*[np.arange(N)] * 3, indexing="ij"), -1).reshape(-1, 3) + 0.5) / N
    phi = rng.normal(size=len(cells))
    rows = np.concatenate([cells, phi[:, None]], axis=1)[rng.permutation(len(cells))]
    images = []
    for half in np.array_split(rows, 2):
        per = np.array_split(half, 3)                          # 3 cpus x 8 octant blocks each
        blocks = {}
        for ib, chunk in enumerate(per, start=1):
            m = len(chunk) // 8
            blocks[(4, ib)] = chunk[:8 * m].reshape(8, m, 4).transpose(0, 2, 1)
        images.append(io.write_poisson(blocks, 3, 0, 4, 4))
    want_cols = [np.concatenate(c) for c in zip(*[io.unpack_poisson(img, 4, 4, 4) for img in images])]
    cols = ingest.read_poisson_output(images, fields, levels, device="cuda:0")
    for name, w in zip(fields, want_cols):
        np.testing.assert_array_equal(cols[name].cpu().numpy(), w)
    vm = ingest.assign_grid(cols["x"], cols["y"], cols["z"], cols["phi"], N, boxsize=100.0, device="cuda:0")
    want_map = io.read_data_assign(N, *want_cols)
    np.testing.assert_array_equal(vm.cpu().numpy(), want_map)
    r = ab.FFTPower(ab.ArrayMesh(vm, BoxSize=100.0), mode="1d", kmin=2 * np.pi / 100.0)
    k, pk, modes = f.power_from_mesh(want_map, None, 100.0)
    np.testing.assert_array_equal(r.power["modes"], modes)
    np.testing.assert_allclose(r.power["power"].real, pk, rtol=1e-4)
