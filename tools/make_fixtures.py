#!/usr/bin/env python
"""Golden P(k) vectors at the BASELINE sizes, from the ORACLE (oracle/pk_oracle_fast.py, float64, all host threads)
on exactly the particle sets bench.py times.  Run on a GPU box (the particle generator runs on the device; the
oracle runs on the box's host cores and needs ~60 GB of RAM at 1024^3):

    python tools/make_fixtures.py c2 c3 c4        ->  gpurun_out/golden_<workload>.npz

The files are then committed as tests/golden/<workload>_pk.npz; bench.py asserts against them after its timed
region at every GPU count and tests/test_gpu_baseline_sizes.py compares the CUDA path with them.
Test infrastructure: nothing in astrild_b200/ reads these.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import WORKLOADS, make_particles   # noqa: E402
from oracle import pk_oracle_fast as f        # noqa: E402


def main():
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    threads = os.cpu_count() or 1
    dev = torch.device("cuda", 0)
    for key in sys.argv[1:]:
        wl = WORKLOADS[key]
        t0 = time.time()
        pos, halos = make_particles(wl, dev)
        host = tuple(c.cpu().numpy() for c in pos)
        hh = None if halos is None else tuple(c.cpu().numpy() for c in halos)
        del pos, halos
        torch.cuda.empty_cache()
        t1 = time.time()
        N, L = wl["mesh"], wl["box"]
        tm = {}
        kw = dict(resampler=wl["resampler"], interlaced=wl["interlaced"], compensated=wl["compensated"], normalize=True,
                  workers=threads, threads=threads, timings=tm, paint_L=1.0, lean=True)
        if hh is None:
            k, P, modes = f.power_from_particles(host, None, N, L, **kw)
        else:                                  # cross spectrum: first = mass-weighted halos, second = matter
            k, P, modes = f.power_from_particles(hh[:3], hh[3], N, L, pos2=host, mass2=None, **kw)
        t2 = time.time()
        meta = {"workload": wl["name"], "threads": threads, "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0),
                "seconds_generate": round(t1 - t0, 1), "seconds_oracle": round(t2 - t1, 1), "oracle_stage_seconds": tm}
        np.savez(os.path.join(out_dir, f"golden_{key}.npz"), k=k, power=P, modes=modes, meta=json.dumps(meta))
        print(key, json.dumps(meta), "P[:3] =", P[:3], flush=True)


if __name__ == "__main__":
    main()
