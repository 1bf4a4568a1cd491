"""The deposit kernels' SOURCE, executed on the CPU by tests/simt (fibers instead of GPU threads), against the oracle.

No GPU needed: this checks the kernels' logic -- brick keys, the shared partition of the interlaced twins, the
in-brick counting sort, moments, warp shuffles, window flush, slab ownership -- with a scheduler that resumes the
threads of a CTA in pseudo-random order, so a missing barrier gives a wrong mesh.  The product never runs this way
(tests/simt is test infrastructure; the CUDA library has no CPU path).  The GPU parity tests proper are
tests/test_gpu_parity.py.
"""
import ctypes as ct
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "simt"))


def load(path):
    lib = ct.CDLL(path)
    lib.simt_deposit_sorted.restype = ct.c_longlong
    lib.simt_deposit_sorted.argtypes = ([ct.c_void_p] * 3 + [ct.c_int, ct.c_int, ct.c_void_p, ct.c_int, ct.c_longlong, ct.c_int,
                                        ct.c_double, ct.c_double, ct.c_int, ct.c_int, ct.c_int, ct.c_void_p, ct.c_void_p, ct.c_int])
    return lib


# the tile kernel's build variants: 0 = the lanes walk their windows in the same order, 1 = lanes whose windows start on
# the same shared-memory bank rotate the order of their x-planes (deposit_sorted.cu, APK_TILE_ROT)
@pytest.fixture(scope="module", params=[0, 1], ids=["rot0", "rot1"])
def simt(request):
    import build_simt
    if os.environ.get("APK_SIMT_LIB"):               # another build of the same kernels (tools: A/B variants)
        return load(os.environ["APK_SIMT_LIB"])
    return load(build_simt.build_rot(request.param))


def deposit(lib, pos, mass, N, L, resampler, pair=False, shift=0.0, soa=False, x0=0, n0=None):
    """-> (mesh, twin or None) as float64 [planes][N][N]; planes = N, or 1 + n0 + 2 for a slab [x0, x0 + n0)."""
    n0 = N if n0 is None else n0
    planes = N if n0 == N else n0 + 3
    m0 = np.zeros((planes, N, 2 * (N // 2 + 1)), np.float32)
    m1 = np.zeros_like(m0) if pair else None
    if soa:
        cols = [np.ascontiguousarray(pos[:, d]) for d in range(3)]
        ptrs, f64 = [c.ctypes.data for c in cols], cols[0].dtype == np.float64
    else:
        p = np.ascontiguousarray(pos)
        ptrs, f64 = [p.ctypes.data, None, None], p.dtype == np.float64
    mp = None if mass is None else np.ascontiguousarray(mass)
    lib.simt_deposit_sorted(ptrs[0], ptrs[1], ptrs[2], int(soa), int(f64), None if mp is None else mp.ctypes.data,
                            int(mp is not None and mp.dtype == np.float64), len(pos), N, 1.0 / L, shift,
                            {"cic": 2, "tsc": 3}[resampler], x0, n0, m0.ctypes.data,
                            None if m1 is None else m1.ctypes.data, 2)
    return m0[:, :, :N].astype(np.float64), (None if m1 is None else m1[:, :, :N].astype(np.float64))


def close(got, want):
    np.testing.assert_allclose(got, want, rtol=0, atol=4e-6 * max(want.max(), 1.0))


@pytest.mark.parametrize("resampler", ["cic", "tsc"])
def test_interlaced_pair_on_cpu_fibers(simt, oracle_fast, resampler):
    rng = np.random.default_rng(3)
    N, L = 32, 1000.0
    pos = (rng.random((12000, 3)) * L).astype(np.float32)
    a, b = deposit(simt, pos, None, N, L, resampler, pair=True)
    close(a, oracle_fast.paint(pos, None, N, L, resampler, 0.0))
    close(b, oracle_fast.paint(pos, None, N, L, resampler, 0.5))
    assert a.sum() == pytest.approx(len(pos), rel=1e-6) and b.sum() == pytest.approx(len(pos), rel=1e-6)


def test_mass_soa_odd_mesh_out_of_box(simt, oracle_fast):
    """N = 45 (no brick edge divides it), SoA columns, masses, positions up to 0.3 L outside the box, shift 0.5."""
    rng = np.random.default_rng(4)
    N, L = 45, 250.0
    pos = (rng.random((9000, 3)) * 1.6 * L - 0.3 * L).astype(np.float32)
    mass = np.exp(rng.normal(0, 1, len(pos))).astype(np.float32)
    a, _ = deposit(simt, pos, mass, N, L, "tsc", shift=0.5, soa=True)
    close(a, oracle_fast.paint(pos, mass, N, L, "tsc", 0.5))


@pytest.mark.parametrize("soa", [True, False])
def test_odd_count_and_misaligned_columns(simt, oracle_fast, soa):
    """A particle count that is not a multiple of the partition's tile (tail threads) and arrays that start 4 bytes off
    a 16-byte boundary must give the same meshes (any vectorised load path has to fall back to scalar loads there)."""
    rng = np.random.default_rng(12)
    N, L, n = 32, 100.0, 10001
    raw = (rng.random((n + 1, 3)) * L).astype(np.float32)
    want = oracle_fast.paint(raw[1:], None, N, L, "tsc", 0.0)
    if soa:
        store = [np.ascontiguousarray(raw[:, d]) for d in range(3)]
        cols = [c[1:] for c in store]             # contiguous views, 4 bytes past a 16-byte boundary
        assert all(c.ctypes.data % 16 == 4 for c in cols)
        pos = np.stack(cols, axis=1)              # deposit(soa=True) re-splits into contiguous columns: rebuild views
        a, _ = deposit(simt, pos, None, N, L, "tsc", soa=True)
        close(a, want)
        # the misaligned pointers themselves
        m0 = np.zeros((N, N, 2 * (N // 2 + 1)), np.float32)
        simt.simt_deposit_sorted(cols[0].ctypes.data, cols[1].ctypes.data, cols[2].ctypes.data, 1, 0, None, 0, n, N, 1.0 / L,
                                 0.0, 3, 0, N, m0.ctypes.data, None, 2)
        close(m0[:, :, :N].astype(np.float64), want)
    else:
        flat = np.zeros(3 * n + 1, np.float32)
        flat[1:] = raw[1:].reshape(-1)
        aos = flat[1:].reshape(n, 3)
        assert aos.ctypes.data % 16 == 4
        m0 = np.zeros((N, N, 2 * (N // 2 + 1)), np.float32)
        simt.simt_deposit_sorted(aos.ctypes.data, None, None, 0, 0, None, 0, n, N, 1.0 / L, 0.0, 3, 0, N, m0.ctypes.data, None, 2)
        close(m0[:, :, :N].astype(np.float64), want)
        a, _ = deposit(simt, np.ascontiguousarray(raw[1:]), None, N, L, "tsc")
        close(a, want)


def test_float64_positions_cic(simt, oracle_fast):
    rng = np.random.default_rng(5)
    N, L = 24, 1.0
    pos = rng.random((6000, 3)) * L
    a, _ = deposit(simt, pos, None, N, L, "cic")
    close(a, oracle_fast.paint(pos, None, N, L, "cic", 0.0))


def test_clustered_cells_take_several_chunks(simt, oracle_fast):
    """11000 particles in one brick (more than the 8191-particle chunk between two flushes of the tile: the second chunk
    runs on a tile re-zeroed by the first flush), 6000 of them in one cell (ranks up to 31 in the rotation's vote)."""
    rng = np.random.default_rng(6)
    N, L = 32, 32.0
    pos = np.concatenate([rng.random((5000, 3)) * [10.0, 5.0, 20.0] + 1.0, rng.random((6000, 3)) * 0.9 + [3.0, 3.0, 3.0],
                          rng.random((500, 3)) * L]).astype(np.float32)
    a, b = deposit(simt, pos, None, N, L, "tsc", pair=True)
    close(a, oracle_fast.paint(pos, None, N, L, "tsc", 0.0))
    close(b, oracle_fast.paint(pos, None, N, L, "tsc", 0.5))


def test_slab_plan_ignores_foreign_particles(simt, oracle_fast):
    """Slab [8, 16) of a 32^3 mesh with ghost planes [7 | 8..15 | 16, 17]: only owned particles are deposited."""
    rng = np.random.default_rng(7)
    N, L, x0, n0 = 32, 1000.0, 8, 8
    pos = (rng.random((15000, 3)) * L).astype(np.float32)
    a, b = deposit(simt, pos, None, N, L, "tsc", pair=True, x0=x0, n0=n0)
    cell = np.floor(pos[:, 0].astype(np.float64) * N / L).astype(int) % N
    own = pos[(cell >= x0) & (cell < x0 + n0)]
    for got, shift in ((a, 0.0), (b, 0.5)):
        full = oracle_fast.paint(own, None, N, L, "tsc", shift)
        close(got, full[x0 - 1: x0 + n0 + 2])
        assert got.sum() == pytest.approx(len(own), rel=1e-6)


@pytest.mark.parametrize("resampler", ["cic", "tsc"])
@pytest.mark.parametrize("N,f64", [(6, True), (12, False), (30, False)])
def test_twin_wrapping_inside_one_brick(simt, oracle_fast, resampler, N, f64):
    """N <= 30: a single brick spans z (N <= 12 / 6: x / y too), so the tile's window wraps around the box onto
    itself, and the shifted twin of a particle in the last cell has its home cell one past the brick's last one.
    Positions include exact cell and half-cell faces."""
    rng = np.random.default_rng(40 + N)
    L = 7.3
    pos = np.concatenate([rng.random((2500, 3)) * 3 * L - L, rng.integers(0, 2 * N + 1, (800, 3)) * 0.5 * L / N])
    pos = pos.astype(np.float64 if f64 else np.float32)
    a, b = deposit(simt, pos, None, N, L, resampler, pair=True, soa=f64)
    close(a, oracle_fast.paint(pos, None, N, L, resampler, 0.0))
    close(b, oracle_fast.paint(pos, None, N, L, resampler, 0.5))


@pytest.mark.parametrize("resampler", ["cic", "tsc"])
@pytest.mark.parametrize("pair", [0, 1])
@pytest.mark.parametrize("N", [1024, 2048, 1000])
def test_brick_keys_at_benchmark_mesh_sizes(simt, resampler, pair, N):
    """The partition's index arithmetic stays in float32 registers (cells up to 2047, bricks up to 256, no integer
    division, no conversion): checked for every particle against integer arithmetic on the float64 grid coordinate.
    Away from cell faces the key is the brick of the mesh-0 home cell and the payload its coordinate in that brick;
    AT a face (within rounding) either neighbour is allowed, but key and payload must still name the same point."""
    rng = np.random.default_rng(N)
    n = 200000
    pos = rng.random((n, 3))
    pos[:20000] = rng.integers(0, 2 * N, (20000, 3)) * (0.5 / N) + rng.normal(0, 1e-5, (20000, 3))   # around cell faces
    pos[20000:22000] = rng.random((2000, 3)) * 3 - 1                                               # outside the box
    pos[22000:22100] = rng.integers(0, 2 * N + 1, (100, 3)) * (0.5 / N)                            # exactly on faces
    pos[22100:26000] = rng.random((3900, 3)) * 16 - 8                                              # folded boxes: many box lengths out
    pos = pos.astype(np.float32)
    simt.simt_brick_keys.restype = None
    simt.simt_brick_keys.argtypes = [ct.c_void_p, ct.c_longlong, ct.c_int, ct.c_double, ct.c_int, ct.c_int] + [ct.c_void_p] * 3
    key = np.zeros(n, np.uint32)
    l = np.zeros((n, 3), np.float32)
    grid = np.zeros(4, np.int32)
    simt.simt_brick_keys(pos.ctypes.data, n, N, 1.0, {"cic": 2, "tsc": 3}[resampler], pair, key.ctypes.data,
                         l.ctypes.data, grid.ctypes.data)
    zc = 32 - ({"cic": 1, "tsc": 2}[resampler]) - pair
    assert grid[3] == zc
    bxy = np.zeros(2, np.int32)
    simt.simt_brick_edges.restype = None
    simt.simt_brick_edges(bxy.ctypes.data_as(ct.c_void_p))
    edge = np.array([bxy[0], bxy[1], zc])
    g = pos.astype(np.float64) * N                               # pos_scale = 1: Ramses-style coordinates
    round_up = 0.5 if resampler == "tsc" else 0.0
    home = np.floor(g + round_up).astype(np.int64)
    cell = home % N
    brick = cell // edge
    want_key = (brick[:, 0] * grid[1] + brick[:, 1]) * grid[2] + brick[:, 2]
    want_l = cell - brick * edge + (g - home)                    # TSC: home + [-0.5, 0.5);  CIC: home + [0, 1)
    clear = np.abs(g + round_up - np.round(g + round_up)).min(axis=1) > 1e-3   # not within rounding of a face of the home cell
    np.testing.assert_array_equal(key[clear], want_key[clear])
    np.testing.assert_allclose(l[clear], want_l[clear], rtol=0, atol=3e-3)      # float32 positions at |g| up to 16000
    # every particle, faces included: brick origin + payload is the particle's grid coordinate (mod N)
    kb = np.stack([key // (grid[1] * grid[2]), (key // grid[2]) % grid[1], key % grid[2]], axis=1).astype(np.int64)
    back = kb * edge + l.astype(np.float64)
    diff = (back - g + N / 2) % N - N / 2
    assert np.abs(diff).max() < 3e-3
    lo, hi = (-0.5, edge - 0.5) if resampler == "tsc" else (0.0, edge)
    assert (l >= lo - 3e-4).all() and (l <= hi + 3e-4).all()
    assert (kb < grid[:3]).all()


def test_the_harness_sees_a_missing_barrier(oracle_fast, tmp_path):
    """Mutation check of the harness itself: without the barrier between the particle loop and the flush of the
    particle-parallel tile kernel, a warp that runs ahead sends (and clears) tile cells other warps are still adding to.
    The scheduler (random warp subsets and single-warp bursts) must turn that into a wrong mesh."""
    import build_simt
    src = build_simt.device_part(os.path.join(build_simt.CSRC, "deposit_sorted.cu"))
    barrier = "__syncthreads();   // every particle of the chunk is in the tile"
    assert src.count(barrier) == 1
    lib = load(build_simt.compile_kernels(src.replace(barrier, "(void)0;"), str(tmp_path)))
    rng = np.random.default_rng(3)
    N, L = 32, 32.0
    pos = np.concatenate([rng.random((9000, 3)) * L, rng.random((5000, 3)) * [10.0, 5.0, 20.0] + 1.0]).astype(np.float32)
    got, _ = deposit(lib, pos, None, N, L, "tsc")
    want = oracle_fast.paint(pos, None, N, L, "tsc", 0.0)
    assert not np.allclose(got, want, rtol=0, atol=1e-3 * want.max())


def test_rank_mod_three_expression_of_the_tile_kernel():
    """brick_tile_kernel takes the rank of a lane among the lanes that start on its bank (0 .. 31) modulo 3 as
    rank - 3 * ((rank * 11) >> 5) (three integer instructions instead of the compiler's ten for % 3): the expression
    itself, and that it is still the one in the source."""
    assert all(r - 3 * ((r * 11) >> 5) == r % 3 for r in range(32))
    src = open(os.path.join(HERE, "..", "astrild_b200", "csrc", "deposit_sorted.cu")).read()
    assert "rank - 3u * ((rank * 11u) >> 5)" in src
