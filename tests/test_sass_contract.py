"""What the compiled sm_100a kernels must (not) contain, read from the library's SASS with cuobjdump (no GPU needed).

These are the machine-level facts DESIGN.md section 4 argues from: the tile kernel accumulates in registers (packed
FFMA2), spreads with shuffles, flushes with fire-and-forget float REDs and never uses a shared-memory float atomic
(a CAS loop on this architecture) or local memory; the binning kernel evicts with float64 / uint64 REDs.
"""
import re
import shutil
import subprocess

import pytest


@pytest.fixture(scope="module")
def sass(built_lib):
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    text = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True, check=True).stdout
    kernels, name = {}, None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
        if m and name:
            kernels[name].append(m.group(1))
    return kernels


def one(kernels, pattern):
    hits = [k for k in kernels if re.search(pattern, k)]
    assert len(hits) == 1, (pattern, hits)
    return kernels[hits[0]]


def count(ops, prefix):
    return sum(op == prefix or op.startswith(prefix + ".") for op in ops)


def test_tsc_tile_kernel_is_register_accumulation_plus_reds(sass):
    ops = one(sass, r"brick_deposit_kernelILi3ELb0ENS_2P3")
    assert count(ops, "FFMA2") >= 9 * 9              # 9 packed FMAs per particle body, 9 unrolled columns
    assert count(ops, "SHFL") >= 18 * 9              # z-spread: two shuffles per (a, b), per column
    assert count(ops, "REDG") == 25                  # one coalesced RED per column of the 5 x 5 window
    assert not any(op.startswith("ATOMS.CAS") for op in ops), "shared-memory float atomic (CAS loop) in the tile kernel"
    assert count(ops, "ATOMG") == 0                  # no returning global atomics: one CTA per brick, no work queue
    assert count(ops, "LDL") + count(ops, "STL") <= 8, "the TSC tile kernel spills"
    assert count(ops, "BAR") <= 10


def test_cic_tile_kernel_flushes_a_4x4_window(sass):
    ops = one(sass, r"brick_deposit_kernelILi2ELb0ENS_2P3")
    assert count(ops, "REDG") == 16
    assert not any(op.startswith("ATOMS.CAS") for op in ops)


def test_partition_counts_with_reds_and_scatters_with_returning_atomics(sass):
    cnt = one(sass, r"brick_count_kernelILi3EfLb1ELb1")          # TSC, float32, SoA, interlaced pair
    assert count(cnt, "REDG") >= 4 and count(cnt, "ATOMG") == 0
    sc = one(sass, r"brick_scatter_kernelILi3EfLb1ELb0ELb1")
    assert count(sc, "ATOMG") >= 4 and count(sc, "STG") >= 12


def test_binning_kernel_evicts_with_float64_reds(sass):
    ops = one(sass, r"bin_power_kernelILb1ELb0ELb1ELi0")        # interlaced auto spectrum, compensated, one-pass mode
    assert sum(op.startswith("RED") and "F64" in op for op in ops) >= 4
    assert count(ops, "ATOMG") == 0 and not any(op.startswith("ATOMS") for op in ops)
    assert count(ops, "LDL") + count(ops, "STL") == 0
