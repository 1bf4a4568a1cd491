"""Interlaced pair deposit on meshes that ONE brick covers along an axis (N <= 30 along z, <= 12 / 6 along x / y).

There the shifted twin of a particle in the last cell wraps around the periodic boundary WITHOUT leaving the brick, so
the brick key alone does not say that it needs a copy of its own.  Found by the randomised CPU-fiber runs of
tests/simt (the GPU suite only had N >= 32 for the pair path); the same cases run on CPU in
tests/test_simt_deposit.py::test_twin_wrapping_inside_one_brick.  Kept in a file of its own, last in the GPU run.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("resampler", ["cic", "tsc"])
@pytest.mark.parametrize("N", [6, 12, 24, 30])
def test_interlaced_pair_on_single_brick_axes(oracle_fast, resampler, N):
    import astrild_b200 as ab
    assert torch.cuda.is_available()
    L = 7.3
    rng = np.random.default_rng(100 + N)
    pos = np.concatenate([rng.random((min(30000, 40 * N ** 3), 3)) * L,
                          rng.integers(0, 2 * N + 1, (2000, 3)) * 0.5 * L / N]).astype(np.float32)   # cell / half-cell faces
    eng = ab.get_engine(N, L)
    pair = eng.deposit_pair(pos, None, resampler, method="sorted")
    for mesh, sh in zip(pair, (0.0, 0.5)):
        want = oracle_fast.paint(pos, None, N, L, resampler, sh)
        got = eng.store_mesh(mesh).cpu().numpy()
        # the defect misplaced ~0.3 of a particle's mass; fp32 accumulation of ~50 particles per cell is ~1e-6
        np.testing.assert_allclose(got, want, rtol=0, atol=2e-5 * want.max())
        assert got.sum() == pytest.approx(len(pos), rel=1e-5)
