"""The slab-decomposed CUDA path on ONE GPU: P ranks emulated as P threads with an in-memory comm.

Every CUDA stage of astrild_b200.distributed (routing kernel, slab deposit with ghost planes,
ghost accumulation, batched 2-D r2c, batched 1-D c2c, binning on the transposed slab) runs for
real; only the collectives are replaced by thread-safe copies.  A 2-GPU NCCL run of the same
path is in test_gpu_slab_nccl (skipped with fewer than 2 devices).
"""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class ThreadComm:
    class Shared:
        def __init__(self, P):
            self.P, self.barrier, self.box = P, threading.Barrier(P), [None] * P

    def __init__(self, shared, rank):
        self.sh, self.P, self.rank = shared, shared.P, rank

    def _publish(self, item):
        self.sh.box[self.rank] = item
        self.sh.barrier.wait()

    def _done(self):
        self.sh.barrier.wait()

    def all_to_all_rows(self, send, counts, alloc):
        self._publish((send, counts))
        parts = []
        for q in range(self.P):
            s, c = self.sh.box[q]
            off = sum(c[: self.rank])
            parts.append(s[off: off + c[self.rank]])
        recv = torch.cat(parts, dim=0).contiguous()
        self._done()
        return recv

    def ring_exchange(self, to_prev, to_next):
        self._publish((to_prev, to_next))
        from_next = self.sh.box[(self.rank + 1) % self.P][0].clone()
        from_prev = self.sh.box[(self.rank - 1) % self.P][1].clone()
        self._done()
        return from_next, from_prev

    def all_to_all_blocks(self, send):
        self._publish(send)
        recv = torch.stack([self.sh.box[q][self.rank] for q in range(self.P)], dim=0).contiguous()
        self._done()
        return recv

    def all_reduce_sum(self, t):
        self._publish(t)
        total = torch.stack([self.sh.box[q] for q in range(self.P)], dim=0).sum(dim=0)
        self._done()
        return total


def _run_emulated(P, N, L, pos, mass, kw, power_kw=None):
    from astrild_b200 import distributed
    shared = ThreadComm.Shared(P)
    results, errors = [None] * P, []

    def work(rank):
        try:
            torch.cuda.set_device(0)
            runner = distributed.SlabPk(N, L, device="cuda:0", comm=ThreadComm(shared, rank),
                                        resampler=kw["resampler"], interlaced=kw["interlaced"], compensated=kw["compensated"])
            runner.backend.side_stream = None      # the in-memory comm has no cross-thread stream ordering
            # twice on the same runner: the second call sees whatever the first one left behind (a grouping pass that
            # wrote past its output when more particles left than the first staging buffer held went unnoticed once)
            for _ in range(2):
                results[rank] = runner.power(pos[rank::P], None if mass is None else mass[rank::P], kmin=2 * np.pi / L,
                                             normalize=kw["normalize"], **(power_kw or {}))
        except BaseException as e:  # noqa: BLE001
            errors.append(e)
            shared.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(P)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


@pytest.mark.parametrize("P", [2, 4])
@pytest.mark.parametrize("kw", [
    dict(resampler="cic", interlaced=False, compensated=False, normalize=True, mass=False),
    dict(resampler="tsc", interlaced=True, compensated=True, normalize=True, mass=True),
    dict(resampler="tsc", interlaced=False, compensated=False, normalize=False, mass=True),
])
def test_slab_path_matches_oracle_and_single_gpu(oracle_fast, P, kw):
    import astrild_b200 as ab
    N, L, Np = 64, 1000.0, 300000
    rng = np.random.default_rng(17)
    pos = (rng.random((Np, 3)) * L).astype(np.float32)
    mass = np.exp(rng.normal(0, 1, Np)).astype(np.float32) if kw["mass"] else None
    want = oracle_fast.power_from_particles(pos, mass, N, L, resampler=kw["resampler"], interlaced=kw["interlaced"],
                                            compensated=kw["compensated"], normalize=kw["normalize"])
    single = ab.FFTPower(ab.CatalogMesh(pos, L, N, weight=mass, resampler=kw["resampler"], interlaced=kw["interlaced"],
                                        compensated=kw["compensated"], normalize=kw["normalize"]), mode="1d", kmin=2 * np.pi / L)
    for res in _run_emulated(P, N, L, pos, mass, kw):
        np.testing.assert_array_equal(res["modes"], want[2])                       # bit-exact, any P
        np.testing.assert_array_equal(res["modes"], single.power["modes"])
        np.testing.assert_allclose(res["k"], want[0], rtol=1e-12)
        np.testing.assert_allclose(res["power"].real, want[1], rtol=1e-4)
        np.testing.assert_allclose(res["power"].real, single.power["power"].real, rtol=2e-5)


@pytest.mark.parametrize("P", [2, 4])
def test_slab_path_kmu_wedges_and_multipoles(P):
    """Row N4 on the slab path: every rank bins its y-range of the transposed grid into (k, mu) wedges, sums and
    integer mode counts are reduced -- equal to the single-GPU FFTPower(mode='2d') (mode counts bit for bit)."""
    import astrild_b200 as ab
    N, L, Np = 64, 400.0, 200000
    rng = np.random.default_rng(41)
    pos = (rng.random((Np, 3)) * L).astype(np.float32)
    kw = dict(resampler="tsc", interlaced=True, compensated=True, normalize=True, mass=False)
    los = (0.0, 0.6, 0.8)
    single = ab.FFTPower(ab.CatalogMesh(pos, L, N, resampler="tsc", interlaced=True, compensated=True, normalize=True),
                         mode="2d", Nmu=4, poles=[0, 2, 4], los=los, kmin=2 * np.pi / L)
    for res in _run_emulated(P, N, L, pos, None, kw, dict(mode="2d", Nmu=4, poles=(0, 2, 4), los=los)):
        np.testing.assert_array_equal(res["modes"], single.power["modes"])
        ok = res["modes"] > 0
        np.testing.assert_allclose(res["k"][ok], single.power["k"][ok], rtol=1e-12)
        np.testing.assert_allclose(res["mu"][ok], single.power["mu"][ok], rtol=1e-12, atol=1e-15)
        scale = np.abs(single.power["power"][ok]).max()
        np.testing.assert_allclose(res["power"][ok].real, single.power["power"][ok].real, rtol=2e-5, atol=2e-5 * scale)
        for ell in (0, 2, 4):
            np.testing.assert_allclose(res["poles"]["power_%d" % ell].real, single.poles["power_%d" % ell].real,
                                       rtol=2e-5, atol=2e-5 * scale)


def test_slab_path_large_sorted_deposit(oracle_fast):
    """Enough particles per rank for the sorted (brick) deposit on slab plans with ghost planes."""
    N, L, Np = 96, 500.0, 1200000
    rng = np.random.default_rng(3)
    pos = (rng.random((Np, 3)) * L).astype(np.float32)
    kw = dict(resampler="tsc", interlaced=True, compensated=True, normalize=True, mass=False)
    want = oracle_fast.power_from_particles(pos, None, N, L, resampler="tsc", interlaced=True, compensated=True,
                                            normalize=True, threads=4, workers=4)
    for res in _run_emulated(2, N, L, pos, None, kw):
        np.testing.assert_array_equal(res["modes"], want[2])
        np.testing.assert_allclose(res["power"].real, want[1], rtol=1e-4)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_gpu_slab_nccl():
    """The same path over NCCL on 2 GPUs (torchrun)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29571", os.path.join(ROOT, "tests", "run_slab_nccl.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "SLAB NCCL OK" in out.stdout
