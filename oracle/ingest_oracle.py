"""CPU restatement of the two ingest steps next to the P(k) path  --  TEST INFRASTRUCTURE ONLY.

* ``read_data_assign``: PowerSpectrum3D._read_data's gridder, /root/reference/src/astrild/power_spectra/power_spectrum_3d.py:142-148.
* ``unpack_poisson``: the record loop of Ecosmog.compress_snapshot, /root/reference/src/astrild/particles/ecosmog.py:184-230,
  transcribed statement by statement (struct.unpack of every block into Python lists), without the set() de-duplication
  that follows it (:236-238) -- see astrild_b200/ingest.py for why that step does not change the gridded field.
* ``write_poisson``: a writer of the same layout for the tests (the reference has none: the files come from the
  ECOSMOG Fortran code).  PARITY NOTE: the reference holds no fixture of this format, so the layout is pinned only by
  the reader above.
Importers: tests/ only.
"""
from __future__ import annotations

from struct import pack, unpack

import numpy as np


def read_data_assign(npar: int, x, y, z, values) -> np.ndarray:
    value_map = np.zeros((npar, npar, npar))
    xi = (npar * x).astype(int)
    yi = (npar * y).astype(int)
    zi = (npar * z).astype(int)
    value_map[(xi, yi, zi)] = values
    return value_map


def unpack_poisson(content: bytes, nfields: int, levelmin: int, levelmax: int, dimensions: int = 3) -> list:
    _dimfac = 2 ** dimensions
    datlis = [[] for _ in range(nfields)]
    pmin = 0
    pmax = 48
    info = unpack("i" * 3 * 4, content[pmin:pmax])
    [ncpu, _ndim, _nlevelmax, nboundary] = [info[1], info[4], info[7], info[10]]
    for _ilevel in range(levelmin, levelmax + 1):
        for _ibound in range(1, nboundary + ncpu + 1):
            pmin0 = pmax
            pmax0 = pmin0 + 4 * 3 * 2
            info = unpack("i" * 3 * 2, content[pmin0:pmax0])
            [_currlevel, ncache] = [info[1], info[4]]
            if ncache == 0:
                pmax = pmax0
                continue
            for _dim in range(1, _dimfac + 1):
                j = 0
                for N in range(1, nfields + 1):
                    pmin = pmax0 + (8 * N - 4) + (N - 1) * 8 * ncache
                    pmax = pmin + ncache * 8
                    info = unpack("d" * ncache, content[pmin:pmax])
                    for floatelem in info:
                        datlis[j].append(floatelem)
                    j += 1
                pmax0 = pmax + 4
            pmax = pmax0
    return [np.asarray(c, dtype=np.float64) for c in datlis]


def write_poisson(blocks: list, ncpu: int, nboundary: int, levelmin: int, levelmax: int, ndim: int = 3) -> bytes:
    """blocks[(level, ibound)] -> array [2**ndim][nfields][ncache] (or missing: ncache = 0).  Fortran unformatted
    records: int32 byte count, payload, int32 byte count."""
    def rec(payload: bytes) -> bytes:
        return pack("i", len(payload)) + payload + pack("i", len(payload))

    out = [rec(pack("i", ncpu)), rec(pack("i", ndim)), rec(pack("i", levelmax)), rec(pack("i", nboundary))]
    for lev in range(levelmin, levelmax + 1):
        for ib in range(1, nboundary + ncpu + 1):
            b = blocks.get((lev, ib))
            ncache = 0 if b is None else int(np.asarray(b).shape[2])
            out += [rec(pack("i", lev)), rec(pack("i", ncache))]
            if ncache:
                a = np.ascontiguousarray(b, dtype=np.float64)
                for dim in range(a.shape[0]):
                    for f in range(a.shape[1]):
                        out.append(rec(a[dim, f].tobytes()))
    return b"".join(out)
