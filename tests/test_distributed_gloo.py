"""world_size-2/4 gloo runs of astrild_b200.distributed's exchange logic on CPUs.

The compute stages are replaced by a NumPy backend built on the oracle (test infrastructure,
defined here, never in the product): what is under test is the slab routing, the ghost-plane
exchange, the x<->y transpose, the transposed binning tables and the final reduction -- the result
must equal the single-process oracle on the union of all ranks' particles.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pk_oracle as o


class NumpyBinning:
    def __init__(self, edges, ia, ib, comp, interlaced):
        self.edges, self.ia, self.ib, self.comp, self.interlaced = edges, ia, ib, comp, interlaced


class NumpySlabBackend:
    """CPU stand-in for CudaSlabBackend: same interface, float64 NumPy arithmetic from the oracle."""

    def __init__(self, N, L, x0, n0, nranks):
        self.N, self.L, self.x0, self.n0, self.nranks = N, L, x0, n0, nranks
        self.Nk = N // 2 + 1

    def _dest(self, pos, pos_scale):
        g = np.floor(np.asarray(pos, dtype=np.float64)[:, 0] * (pos_scale * self.N)).astype(np.int64) % self.N
        return g // (self.N // self.nranks)

    def route(self, pos, mass, pos_scale):
        p = np.asarray(pos, dtype=np.float64)
        dest = self._dest(p, pos_scale)
        me = self.x0 // self.n0
        leave = np.flatnonzero(dest != me)
        order = leave[np.argsort(dest[leave], kind="stable")]
        counts = np.bincount(dest[leave], minlength=self.nranks).tolist()
        sm = None if mass is None else torch.from_numpy(np.asarray(mass, dtype=np.float64)[order].copy())
        return torch.from_numpy(p[order].copy()), sm, counts

    def empty_like_rows(self, like, rows):
        return torch.empty((rows,) + tuple(like.shape[1:]), dtype=like.dtype)

    def deposit(self, pos, mass, resampler, shift, pos_scale, out=None):
        N = self.N
        pos = torch.as_tensor(np.asarray(pos, dtype=np.float64))
        mass = None if mass is None else torch.as_tensor(np.asarray(mass, dtype=np.float64))
        mine = self._dest(pos.numpy(), pos_scale) == self.x0 // self.n0        # a slab deposit ignores foreign particles
        pos = pos[torch.from_numpy(mine)]
        mass = None if mass is None else mass[torch.from_numpy(mine)]
        full = o.paint(pos.numpy() * (pos_scale * self.L), 1.0 if mass is None else mass.numpy(), N, self.L,
                       resampler, shift)
        if self.nranks == 1:
            return torch.from_numpy(full) if out is None else out.add_(torch.from_numpy(full))
        planes = (np.arange(self.x0 - 1, self.x0 + self.n0 + 2)) % N
        others = np.setdiff1d(np.arange(N), planes)
        assert not full[others].any(), "a routed particle touched a plane outside slab + ghosts"
        slab = torch.from_numpy(full[planes].copy())
        return slab if out is None else out.add_(slab)

    def accumulate(self, dst, src):
        dst += src

    def mesh_sum(self, owned):
        return owned.sum().reshape(1)

    def fft2d(self, owned):
        return torch.from_numpy(np.fft.rfft2(owned.numpy(), axes=(1, 2)))

    def fft1d(self, grid, ny):
        return torch.from_numpy(np.fft.fft(grid.numpy(), axis=0))

    def make_binning(self, y0, ny, kmin, dk, kmax, comp, interlaced):
        return NumpyBinning(o.k_edges(self.N, self.L, kmin, dk, kmax), np.arange(self.N), np.arange(y0, y0 + ny),
                            comp, interlaced)

    def bin(self, b, c1, c1s):
        N, L = self.N, self.L
        kx, ky, kz = o.k_tables(N, L)
        c = c1.numpy()
        if b.interlaced:
            ph = 0.5 * (kx[b.ia][:, None, None] + ky[b.ib][None, :, None] + kz[None, None, :]) * (L / N)
            c = 0.5 * c + 0.5 * c1s.numpy() * np.exp(1j * ph)
        if b.comp is not None:
            w = o.compensation_1d(b.comp[0], b.comp[1], N)
            c = c / (w[b.ia][:, None, None] * w[b.ib][None, :, None] * w[None, None, :self.Nk])
        k2 = (kx[b.ia][:, None, None] ** 2 + ky[b.ib][None, :, None] ** 2) + kz[None, None, :] ** 2
        dig = np.digitize(k2.ravel(), b.edges ** 2)
        w = np.broadcast_to(o.hermitian_weights(N)[None, None, :], k2.shape).ravel()
        p = (c * np.conj(c)).real
        dc = (b.ia[:, None, None] == 0) & (b.ib[None, :, None] == 0) & (np.arange(self.Nk)[None, None, :] == 0)
        p[dc] = 0.0
        nb1 = len(b.edges) + 1
        raw = np.zeros((4, nb1))
        raw[0] = np.bincount(dig, weights=w * np.sqrt(k2.ravel()), minlength=nb1)
        raw[1] = np.bincount(dig, weights=w * p.ravel(), minlength=nb1)
        raw[3] = np.bincount(dig, weights=w, minlength=nb1)
        return torch.from_numpy(raw)

    def to_reduce_tensors(self, raw, total):
        return (torch.cat([raw[:3].reshape(-1), total.reshape(1).to(torch.float64)]),
                torch.from_numpy(np.rint(raw[3].numpy()).astype(np.int64)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, cfg, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from astrild_b200 import distributed
        N, L = cfg["N"], cfg["L"]
        rng = np.random.default_rng(cfg["seed"])
        pos = rng.random((cfg["Np"], 3)) * L
        mass = rng.random(cfg["Np"]) + 0.5 if cfg["mass"] else None
        mine = slice(rank, None, world)                     # arbitrary split: routing must fix it
        n0 = N // world
        backend = NumpySlabBackend(N, L, rank * n0, n0, world)
        runner = distributed.SlabPk(N, L, resampler=cfg["resampler"], interlaced=cfg["interlaced"],
                                    compensated=cfg["compensated"], backend=backend)
        res = runner.power(pos[mine], None if mass is None else mass[mine], kmin=2 * np.pi / L,
                           normalize=cfg["normalize"])
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), k=res["k"], p=res["power"].real, modes=res["modes"])
    finally:
        dist.destroy_process_group()


CASES = [
    dict(N=16, L=100.0, Np=4000, seed=1, resampler="cic", interlaced=False, compensated=False, normalize=True, mass=False),
    dict(N=16, L=100.0, Np=4000, seed=2, resampler="tsc", interlaced=True, compensated=True, normalize=True, mass=True),
    dict(N=24, L=250.0, Np=3000, seed=3, resampler="tsc", interlaced=False, compensated=False, normalize=False, mass=True),
]


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("cfg", CASES)
def test_slab_exchange_logic_matches_single_rank_oracle(tmp_path, world, cfg):
    mp.spawn(_worker, args=(world, _free_port(), cfg, str(tmp_path)), nprocs=world, join=True)
    N, L = cfg["N"], cfg["L"]
    rng = np.random.default_rng(cfg["seed"])
    pos = rng.random((cfg["Np"], 3)) * L
    mass = rng.random(cfg["Np"]) + 0.5 if cfg["mass"] else None
    k, pk, modes = o.power_from_particles(pos, mass, N, L, resampler=cfg["resampler"], interlaced=cfg["interlaced"],
                                          compensated=cfg["compensated"], normalize=cfg["normalize"])
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), f"r{r}.npz"))
        np.testing.assert_array_equal(z["modes"], modes)
        np.testing.assert_allclose(z["k"], k, rtol=1e-12)
        np.testing.assert_allclose(z["p"], pk, rtol=1e-9)
