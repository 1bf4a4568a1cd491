"""INTEGRATION.md section B, executed verbatim: the ctypes stub a maintainer would add to astrild binds the C ABI's
one-call entry points; its (k, Pk, modes) must match the oracle.  ("INTEGRATION.md's stub itself is never
executed by a test".)"""
import os
import pathlib
import re

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parent.parent


def _stub_namespace():
    text = (ROOT / "INTEGRATION.md").read_text()
    section = text[text.index("## B. Binding the C ABI directly"):text.index("## C. Multi-GPU")]
    blocks = re.findall(r"```python\n(.*?)```", section, flags=re.S)
    assert len(blocks) == 1
    ns = {"c_lib_path": ROOT / "astrild_b200" / "lib"}
    exec(compile(blocks[0], "INTEGRATION.md#B", "exec"), ns)
    return ns


def test_integration_stub_halo_power_spectrum(oracle_fast):
    ns = _stub_namespace()
    nbins, boxsize, n = 64, 500.0, 200000
    rng = np.random.default_rng(2)
    pos = (rng.random((n, 3)) * boxsize).astype(np.float32)
    mass = np.exp(rng.normal(2.0, 1.0, n)).astype(np.float32)
    keep = []

    def alloc(nbytes):
        t = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device="cuda:0")
        keep.append(t)
        return t.data_ptr()

    dp, dm = torch.from_numpy(pos).cuda(), torch.from_numpy(mass).cuda()
    k, pk, modes = ns["halo_power_spectrum"](dp.data_ptr(), dm.data_ptr(), n, nbins, boxsize, alloc)
    wk, wpk, wmodes = oracle_fast.power_from_particles(pos, mass, nbins, boxsize, resampler="tsc")
    np.testing.assert_array_equal(modes, wmodes)
    np.testing.assert_allclose(k, wk, rtol=1e-12)
    np.testing.assert_allclose(pk, wpk, rtol=1e-4)
    assert len(k) == nbins // 2 - 1


def test_integration_stub_mesh_power_spectrum(oracle_fast):
    ns = _stub_namespace()
    N, L = 48, 250.0
    rng = np.random.default_rng(5)
    vm = rng.normal(5.0, 1.0, (N, N, N))
    vm2 = vm * 0.5 + rng.normal(0.0, 1.0, (N, N, N))
    keep = []

    def alloc(nbytes):
        t = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device="cuda:0")
        keep.append(t)
        return t.data_ptr()

    d1, d2 = torch.from_numpy(vm).cuda(), torch.from_numpy(vm2).cuda()
    k, pk, modes = ns["mesh_power_spectrum"](d1.data_ptr(), None, N, L, alloc)
    wk, wpk, wmodes = oracle_fast.power_from_mesh(vm, None, L)
    np.testing.assert_array_equal(modes, wmodes)
    np.testing.assert_allclose(k, wk, rtol=1e-12)
    np.testing.assert_allclose(pk, wpk, rtol=1e-4)
    k, pk, modes = ns["mesh_power_spectrum"](d1.data_ptr(), d2.data_ptr(), N, L, alloc)
    wk, wpk, wmodes = oracle_fast.power_from_mesh(vm, vm2, L)
    np.testing.assert_allclose(pk, wpk, rtol=1e-4)
