"""The C-ABI library builds, loads without a GPU and exports every symbol the header declares."""
import ctypes as ct
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "astrild_pk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(apk_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("apk_plan_create", "apk_deposit", "apk_load_mesh", "apk_fft_r2c", "apk_bin_power",
                 "apk_binning_create", "apk_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = ct.CDLL(built_lib)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    lib.apk_version.restype = ct.c_int
    assert lib.apk_version() == 100


def test_binding_table_matches_header(built_lib):
    from astrild_b200 import _lib
    declared = set(declared_symbols())
    bound = set(_lib.SIGNATURES) | {"apk_version", "apk_last_error"}
    assert declared == bound


def test_no_torch_types_in_abi():
    text = open(os.path.join(ROOT, "include", "astrild_pk.h")).read()
    assert "torch" not in text.lower().replace("pytorch", "") and "at::" not in text


def test_library_is_sm100a_only(built_lib):
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", built_lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


# ---- the binning tables built inside the library (host code: runs without a GPU) ------------------------------------
import numpy as np
import pytest


def _tables_lib(built_lib):
    lib = ct.CDLL(built_lib)
    d, i, vp = ct.c_double, ct.c_int, ct.c_void_p
    lib.apk_tables_k_axis.argtypes = [i, d, i, vp]
    lib.apk_tables_k_edges.argtypes = [i, d, d, d, d, vp, i, ct.POINTER(i)]
    lib.apk_tables_hermitian_weights.argtypes = [i, vp]
    lib.apk_tables_compensation.argtypes = [i, i, i, vp]
    lib.apk_tables_interlace_phase.argtypes = [i, d, vp]
    return lib


@pytest.mark.parametrize("N", [8, 33, 128, 512, 1024])
@pytest.mark.parametrize("L", [1000.0, 500.0, 7.3])
def test_library_tables_bit_identical_to_the_numpy_tables(built_lib, N, L):
    """The bin-deciding tables (per-axis k, edges, Hermitian weights) made by apk_tables_* equal astrild_b200/tables.py --
    and therefore the oracle's -- bit for bit; the sin / pow based ones (compensation, phase) to 1e-15."""
    from astrild_b200 import tables
    lib = _tables_lib(built_lib)
    for code, dt in ((1, np.float64), (0, np.float32)):
        k = np.empty(N)
        assert lib.apk_tables_k_axis(N, L, code, k.ctypes.data) == 0
        np.testing.assert_array_equal(k, tables.k_axis(N, L, dt))
    for kmin, dk, kmax in ((2 * np.pi / L, 0.0, 0.0), (0.0, 0.0, 0.0), (0.013, 0.0071, 0.9 * np.pi * N / L)):
        want = tables.k_edges(N, L, kmin, dk or None, kmax or None)
        n = ct.c_int()
        assert lib.apk_tables_k_edges(N, L, kmin, dk, kmax, None, 0, ct.byref(n)) == 0
        assert n.value == len(want)
        e = np.empty(n.value)
        assert lib.apk_tables_k_edges(N, L, kmin, dk, kmax, e.ctypes.data, len(e), ct.byref(n)) == 0
        np.testing.assert_array_equal(e, want)
    w = np.empty(N // 2 + 1)
    assert lib.apk_tables_hermitian_weights(N, w.ctypes.data) == 0
    np.testing.assert_array_equal(w, tables.hermitian_weights(N))
    for rs, name in ((2, "cic"), (3, "tsc")):
        for inter in (0, 1):
            c = np.empty(N)
            assert lib.apk_tables_compensation(rs, inter, N, c.ctypes.data) == 0
            np.testing.assert_allclose(c, tables.compensation_axis(name, bool(inter), N), rtol=1e-15, atol=0)
    p = np.empty(N)
    assert lib.apk_tables_interlace_phase(N, L, p.ctypes.data) == 0
    np.testing.assert_array_equal(p, tables.interlace_phase_axis(N, L))
