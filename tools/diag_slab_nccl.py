"""torchrun diagnostic (tools, not product): the slab path over NCCL against the oracle with features switched off one
at a time -- APK_SLAB_P2P=0 (NCCL all-to-all transposes), DIAG_NO_SIDE=1 (no side stream: no overlap)."""
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
for kw in (dict(resampler="cic", interlaced=False, compensated=False), dict(resampler="tsc", interlaced=True, compensated=True)):
    runner = distributed.SlabPk(N, L, device=f"cuda:{local}", **kw)
    if os.environ.get("DIAG_NO_SIDE"):
        runner.backend.side_stream = None
    for rep in range(2):
        res = runner.power(pos[rank::world], None, kmin=2 * np.pi / L, normalize=True)
        if rank == 0:
            if rep == 0:
                want = oracle.power_from_particles(pos, None, N, L, normalize=True, threads=4, workers=4, **kw)
            rel = np.abs(res["power"].real / want[1] - 1)
            print(f"DIAG world={world} p2p={os.environ.get('APK_SLAB_P2P', '1')} noside={os.environ.get('DIAG_NO_SIDE', '0')} "
                  f"{kw['resampler']} rep={rep} transpose={runner.last_info.get('transpose')} modes_equal={np.array_equal(res['modes'], want[2])} "
                  f"max_rel_P={rel.max():.3e} first_bins={res['power'].real[:3]} want={want[1][:3]}", flush=True)
dist.barrier()
dist.destroy_process_group()
