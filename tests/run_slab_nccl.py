"""torchrun entry: slab P(k) over NCCL vs the oracle (launched by tests/test_gpu_slab.py and by hand)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astrild_b200 import distributed  # noqa: E402
from oracle import pk_oracle_fast as oracle  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, L, Np = 128, 1000.0, 2000000
rng = np.random.default_rng(5)
pos = (rng.random((Np, 3)) * L).astype(np.float32)
mass = np.exp(rng.normal(0, 1, Np)).astype(np.float32)
for kw in (dict(resampler="cic", interlaced=False, compensated=False), dict(resampler="tsc", interlaced=True, compensated=True)):
    runner = distributed.SlabPk(N, L, device=f"cuda:{local}", **kw)
    for rep in range(2):                      # the second call runs on the state the first one left
        res = runner.power(pos[rank::world], mass[rank::world], kmin=2 * np.pi / L, normalize=True)
        if rank == 0:
            if rep == 0:
                want = oracle.power_from_particles(pos, mass, N, L, normalize=True, threads=4, workers=4, **kw)
            assert np.array_equal(res["modes"], want[2]), "mode counts differ"
            np.testing.assert_allclose(res["k"], want[0], rtol=1e-12)
            np.testing.assert_allclose(res["power"].real, want[1], rtol=1e-4)
# (k, mu) wedges and multipoles over NCCL (row N4): against the oracle's project_to_basis on the same combined field
from oracle import pk_oracle as slow  # noqa: E402
runner = distributed.SlabPk(N, L, device=f"cuda:{local}", resampler="tsc", interlaced=True, compensated=True)
res = runner.power(pos[rank::world], None, kmin=2 * np.pi / L, normalize=True, mode="2d", Nmu=4, poles=(0, 2), los=(0.0, 0.0, 1.0))
if rank == 0:
    r0, r1 = oracle.paint(pos, None, N, L, "tsc", 0.0, threads=4), oracle.paint(pos, None, N, L, "tsc", 0.5, threads=4)
    s = N ** 3 / r0.sum()
    field = slow.compensate(slow.interlace_combine(slow.r2c(r0) * s, slow.r2c(r1) * s, N, L), "tsc", True, N)
    want = slow.fftpower_2d(field, None, N, L, Nmu=4, poles=(0, 2), kmin=2 * np.pi / L)
    assert np.array_equal(res["modes"], want["modes"]), "(k, mu) mode counts differ"
    ok = want["modes"] > 0
    scale = np.abs(want["power"][ok]).max()
    np.testing.assert_allclose(res["power"][ok].real, want["power"][ok].real, rtol=1e-4, atol=1e-4 * scale)
    np.testing.assert_allclose(res["poles"]["power_2"].real, want["poles"]["power_2"].real, rtol=1e-4, atol=1e-4 * scale)
dist.barrier()
if rank == 0:
    print("SLAB NCCL OK", world, "ranks")
dist.destroy_process_group()
