"""The remaining hand-written kernels' SOURCE on CPU fibers (tests/simt): direct-atomic deposit, slab routing,
peer-store transpose, ghost-plane adds, gridded-field helpers -- against the oracle / NumPy.  No GPU needed; see
tests/test_simt_deposit.py for what this kind of test does and does not show."""
import ctypes as ct
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "simt"))
vp, i32, i64, f64 = ct.c_void_p, ct.c_int, ct.c_longlong, ct.c_double


@pytest.fixture(scope="module")
def misc():
    import build_simt
    lib = ct.CDLL(build_simt.build_misc())
    lib.simt_deposit_atomic.argtypes = [vp, vp, vp, i32, i32, vp, i32, i64, i32, f64, f64, i32, i32, i32, vp, i32]
    lib.simt_route.argtypes = [vp, vp, vp, vp, i64, i32, f64, i32, i32, vp, i64, vp, vp, i32]
    lib.simt_transpose_p2p.argtypes = [vp, vp, i32, i32, i32, i32, i32]
    lib.simt_accumulate.argtypes = [vp, vp, i64, i32]
    lib.simt_mesh_roundtrip.argtypes = [vp, i64, i32, vp, vp, vp, f64, vp, i32]
    lib.simt_assign_grid.argtypes = [vp, vp, vp, i32, vp, i32, i64, i32, vp, vp, vp, i32]
    lib.simt_gather_records.argtypes = [vp, vp, i64, vp]
    return lib


def ptr(a):
    return None if a is None else a.ctypes.data_as(vp)


@pytest.mark.parametrize("resampler,code", [("nearest", 1), ("cic", 2), ("tsc", 3)])
def test_atomic_deposit_on_cpu_fibers(misc, oracle_fast, resampler, code):
    rng = np.random.default_rng(21)
    N, L = 20, 100.0
    pos = (rng.random((5000, 3)) * 1.4 * L - 0.2 * L).astype(np.float32)
    mass = rng.random(5000).astype(np.float32)
    mesh = np.zeros((N, N, 2 * (N // 2 + 1)), np.float32)
    assert misc.simt_deposit_atomic(ptr(pos), None, None, 0, 0, ptr(mass), 0, len(pos), N, 1.0 / L, 0.5, code, 0, N, ptr(mesh), 2) == 0
    want = oracle_fast.paint(pos, mass, N, L, resampler, 0.5)
    np.testing.assert_allclose(mesh[:, :, :N], want, rtol=0, atol=4e-6 * want.max())


def test_route_extracts_exactly_the_leavers(misc):
    """Rank 1 of 4 on a 32^3 mesh: every particle whose floor(g_x) lies outside planes [8, 16) leaves, grouped by
    destination, nothing else is touched; positions outside the box wrap to their periodic owner."""
    rng = np.random.default_rng(22)
    N, P, rank, L = 32, 4, 1, 1.0
    n0 = N // P
    x = np.clip(rng.normal((rank + 0.5) / P, 0.12, 40000), -0.3, 1.3).astype(np.float32)
    y, z = rng.random(40000).astype(np.float32), rng.random(40000).astype(np.float32)
    mass = rng.random(40000).astype(np.float32)
    cell = np.floor(x.astype(np.float64) * N).astype(np.int64) % N
    dest = cell // n0
    leaving = dest != rank
    cap = int(leaving.sum()) + 100
    counts = np.zeros(2 * P, np.uint64)
    out_pos = np.full((cap, 3), np.nan, np.float32)
    out_mass = np.full(cap, np.nan, np.float32)
    assert misc.simt_route(ptr(x), ptr(y), ptr(z), ptr(mass), len(x), N, 1.0 / L, P, rank * n0, ptr(counts), cap,
                           ptr(out_pos), ptr(out_mass), 2) == 0
    want_counts = np.bincount(dest[leaving], minlength=P)
    np.testing.assert_array_equal(counts[:P].astype(np.int64), want_counts)
    assert counts[rank] == 0
    start = 0
    for d in range(P):                       # each destination's block holds exactly its particles (any order)
        n = int(want_counts[d])
        got = out_pos[start:start + n]
        sel = leaving & (dest == d)
        want = np.stack([x[sel], y[sel], z[sel]], axis=1)
        order_g, order_w = np.lexsort(got.T[::-1]), np.lexsort(want.T[::-1])
        np.testing.assert_array_equal(got[order_g], want[order_w])
        np.testing.assert_array_equal(np.sort(out_mass[start:start + n]), np.sort(mass[sel]))
        start += n


def test_route_with_too_small_a_staging_buffer_writes_nothing(misc):
    """More leavers than `capacity`: the counts are complete (the caller reads them and repeats the pass with a larger
    buffer) and the grouping pass must not touch out_pos at all -- its cursors come from the FULL counts, so grouping the
    staged part wrote past the end of out_pos with three or more ranks (found on 4 and 8 GPUs in round 2: the second
    P(k) on the same plan came out wrong)."""
    rng = np.random.default_rng(24)
    N, P, rank, L = 32, 4, 1, 1.0
    n0 = N // P
    x, y, z = (rng.random(20000).astype(np.float32) for _ in range(3))
    cell = np.floor(x.astype(np.float64) * N).astype(int) % N
    want_counts = np.bincount((cell // n0)[(cell // n0) != rank], minlength=P)
    cap = 1000                                     # << 15000 leavers
    counts = np.zeros(2 * P, np.uint64)
    guard = np.full((cap + 20000, 3), np.nan, np.float32)     # out_pos is the first `cap` rows; the rest is the guard
    assert misc.simt_route(ptr(x), ptr(y), ptr(z), None, len(x), N, 1.0 / L, P, rank * n0, ptr(counts), cap, ptr(guard),
                           None, 2) == 0
    np.testing.assert_array_equal(counts[:P].astype(np.int64), want_counts)
    assert int(counts[:P].sum()) > cap
    assert np.isnan(guard).all(), "the grouping pass wrote although the staging buffer had overflowed"


def test_peer_store_transpose_is_the_slab_transpose(misc):
    """Every rank stores its [n0][N][nz] x-slab into all receive buffers: rank s ends up with [N][ny][nz] = the
    full grid restricted to its y range."""
    rng = np.random.default_rng(23)
    N, P = 16, 4
    nz, n0 = N // 2 + 1, N // P
    full = (rng.normal(size=(N, N, nz)) + 1j * rng.normal(size=(N, N, nz))).astype(np.complex64)
    recv = [np.zeros((N, n0, nz), np.complex64) for _ in range(P)]
    table = (vp * P)(*[r.ctypes.data for r in recv])
    for r in range(P):
        slab = np.ascontiguousarray(full[r * n0:(r + 1) * n0])
        assert misc.simt_transpose_p2p(ptr(slab), table, r, P, N, nz, 1) == 0
    for s in range(P):
        np.testing.assert_array_equal(recv[s], full[:, s * n0:(s + 1) * n0, :])


def test_ghost_add_and_gridded_field_helpers(misc):
    rng = np.random.default_rng(24)
    a, b = rng.random(100003).astype(np.float32), rng.random(100003).astype(np.float32)
    want = a + b
    assert misc.simt_accumulate(ptr(a), ptr(b), len(a), 2) == 0
    np.testing.assert_array_equal(a, want)
    N = 12
    field = rng.normal(5.0, 1.0, (N, N, N))
    ldz = 2 * (N // 2 + 1)
    mesh = np.full((N, N, ldz), np.nan, np.float32)
    back = np.zeros((N, N, N))
    s_field, s_mesh = np.zeros(1), np.zeros(1)
    assert misc.simt_mesh_roundtrip(ptr(field), N * N, N, ptr(s_field), ptr(mesh), ptr(s_mesh), 2.0, ptr(back), 2) == 0
    assert s_field[0] == pytest.approx(field.sum(), rel=1e-13)
    np.testing.assert_allclose(mesh[:, :, :N], field - field.mean(), rtol=0, atol=1e-6)
    assert np.all(mesh[:, :, N:] == 0)                              # the r2c padding is cleared
    assert abs(s_mesh[0]) < 1e-3                                    # mean removed
    np.testing.assert_allclose(back, 2.0 * mesh[:, :, :N].astype(np.float64), rtol=0, atol=0)


@pytest.mark.parametrize("pos_dtype,val_dtype", [(np.float64, np.float64), (np.float32, np.float32), (np.float64, np.float32)])
def test_assign_grid_is_numpy_fancy_index_assignment(misc, pos_dtype, val_dtype):
    """ingest.cu's two passes == value_map[(N*x).astype(int), ...] = values: truncation toward zero, a negative index
    wraps once, the LAST of several samples of one cell wins, out-of-range samples are counted (row N1)."""
    from oracle import ingest_oracle as io
    rng = np.random.default_rng(5)
    N, n = 12, 6000                                    # 3.5 samples per cell: plenty of ties
    x, y, z = (rng.random(n).astype(pos_dtype) for _ in range(3))
    x[:50] = -rng.random(50).astype(pos_dtype) * 0.9   # negative coordinates: NumPy counts those indices from the end
    vals = rng.normal(size=n).astype(val_dtype)
    want = io.read_data_assign(N, x, y, z, vals)
    got = np.full((N, N, N), np.nan)
    winner = np.zeros(N ** 3, np.uint32)
    bad = np.zeros(1, np.uint64)
    assert misc.simt_assign_grid(ptr(x), ptr(y), ptr(z), int(pos_dtype == np.float64), ptr(vals), int(val_dtype == np.float64),
                                 n, N, ptr(got), ptr(winner), ptr(bad), 2) == 0
    assert bad[0] == 0
    np.testing.assert_array_equal(got, want)
    x[7] = 1.5                                         # index 18 on a 12-cell axis: NumPy raises, the kernel counts
    misc.simt_assign_grid(ptr(x), ptr(y), ptr(z), int(pos_dtype == np.float64), ptr(vals), int(val_dtype == np.float64),
                          n, N, ptr(got), ptr(winner), ptr(bad), 2)
    assert bad[0] == 1


def test_gather_records_picks_the_float64_blocks(misc):
    """The record gather against the reference's unpack loop on a synthetic output_poisson image (row N1)."""
    from astrild_b200.ingest import poisson_record_pieces
    from oracle import ingest_oracle as io
    rng = np.random.default_rng(6)
    nfields, levels = 3, (6, 7)
    blocks = {}
    for lev in levels:
        for ib in range(1, 2 + 3 + 1):
            ncache = int(rng.integers(0, 30))
            if ncache:
                blocks[(lev, ib)] = rng.random((8, nfields, ncache))
    img = io.write_poisson(blocks, 3, 2, min(levels), max(levels))
    want = io.unpack_poisson(img, nfields, min(levels), max(levels))
    pieces, counts = poisson_record_pieces(img, nfields, min(levels), max(levels))
    raw = np.frombuffer(img, dtype=np.uint8).copy()
    for j in range(nfields):
        out = np.full(counts[j], np.nan)
        pj = np.ascontiguousarray(pieces[j], dtype=np.int64)
        assert misc.simt_gather_records(ptr(raw), ptr(pj), len(pj), ptr(out)) == 0
        np.testing.assert_array_equal(out, np.asarray(want[j]))
