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
