"""What the compiled sm_100a kernels must (not) contain, read from the library's SASS with cuobjdump (no GPU needed).

These are the machine-level facts DESIGN.md section 4 argues from: the tile kernel accumulates fixed-point integers in
shared memory with native ATOMS.ADD, flushes with fire-and-forget float REDs and never uses a shared-memory float
atomic (a CAS loop on this architecture) or local memory; the binning kernel evicts with float64 / uint64 REDs.
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


def no_spills_in_particle_loop(ops):
    """No local memory between the particle's loads and its last shared-memory atomic (the per-brick set-up and the flush,
    around the out-of-line wrap of far indices, may park a few registers)."""
    atoms = [i for i, op in enumerate(ops) if op.startswith("ATOMS.ADD")]
    spills = [i for i, op in enumerate(ops) if op.startswith("LDL") or op.startswith("STL")]
    return len(spills) <= 32 and not any(atoms[0] - 100 <= i <= atoms[-1] for i in spills)


def test_tsc_tile_kernel_is_integer_shared_atomics_plus_reds(sass):
    ops = one(sass, r"brick_tile_kernelILi3ELb0ELb1ELi1ENS_2P3")        # TSC, unit masses, interlaced pair
    assert count(ops, "ATOMS.ADD") == 27             # one native integer shared-memory atomic per window cell
    assert not any(op.startswith("ATOMS.CAS") for op in ops), "shared-memory CAS loop (float atomic) in the tile kernel"
    assert count(ops, "F2I") <= 8                    # the weight's subnormal bit pattern IS the fixed-point value: no F2I,
    assert 36 <= count(ops, "FMUL") <= 80            # ... one FMUL per cell (27 + 9 pair products + the axis weights; the
                                                     #     flush's two unrolled versions scale one value per y-row each)
    assert not any(op.startswith("FMUL") and ".FTZ" in op for op in ops), "flush-to-zero would zero every weight"
    assert count(ops, "REDG") >= 1 and count(ops, "ATOMG") == 0      # flush: fire-and-forget float REDs
    assert no_spills_in_particle_loop(ops), "the TSC tile kernel spills in its particle loop"
    assert count(ops, "BAR") <= 6
    # lanes starting on the same bank are ranked with ONE match instruction (no ballot ladder), and the rotation is
    # selects on registers: the 27 updates keep immediate offsets, nothing is indexed dynamically (no local memory, above)
    assert count(ops, "MATCH.ANY") == 1 and count(ops, "VOTE") == 0
    # the flush reads the tile and nothing else from shared memory (its mesh offsets are register arithmetic): the LDS of
    # the kernel are the flush's two code versions (with / without re-zeroing), each unrolled over the 11 rows of a plane
    assert count(ops, "LDS") <= 22


def test_cic_and_mass_tile_kernels(sass):
    ops = one(sass, r"brick_tile_kernelILi2ELb0ELb0ELi0ENS_2P3")
    assert count(ops, "ATOMS.ADD") >= 8 and not any(op.startswith("ATOMS.CAS") for op in ops)
    assert count(ops, "MATCH") == 0                  # CIC: 8 updates per particle do not pay for the vote (measured)
    ops = one(sass, r"brick_tile_kernelILi3ELb1ELb0ELi0ENS_2P4")
    assert count(ops, "ATOMS.ADD") >= 27 and not any(op.startswith("ATOMS.CAS") for op in ops)
    assert count(ops, "F2I") >= 27                   # with masses: FMUL + F2I at the chunk's own scale
    assert no_spills_in_particle_loop(ops)


def test_partition_counts_with_reds_and_scatters_with_returning_atomics(sass):
    cnt = one(sass, r"brick_count_kernelILi3EfLb1E")             # TSC, float32, SoA
    assert count(cnt, "REDG") >= 4 and count(cnt, "ATOMG") == 0
    sc = one(sass, r"brick_scatter_kernelILi3EfLb1ELb0E")
    assert count(sc, "ATOMG") >= 4 and count(sc, "STG") >= 12


def test_binning_kernel_evicts_with_float64_reds(sass):
    ops = one(sass, r"bin_power_kernelILb1ELb0ELb1ELi0")        # interlaced auto spectrum, compensated, one-pass mode
    assert sum(op.startswith("RED") and "F64" in op for op in ops) >= 4
    assert count(ops, "ATOMG") == 0 and not any(op.startswith("ATOMS") for op in ops)
    assert count(ops, "LDL") + count(ops, "STL") == 0
