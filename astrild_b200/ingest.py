"""Device-side ingest for the P(k) path.

* ``assign_grid``: the NGP *assignment* of ``PowerSpectrum3D._read_data``
  (/root/reference/src/astrild/power_spectra/power_spectrum_3d.py:142-148) as a CUDA kernel
  (``apk_assign_grid``): ``value_map[((N*x).astype(int), (N*y).astype(int), (N*z).astype(int))] = values``.
* ``read_poisson_output``: the record reader of ``Ecosmog.compress_snapshot``
  (/root/reference/src/astrild/particles/ecosmog.py:184-230).  The host walks the Fortran record headers of the file
  image (a few integers per AMR level and cpu), the image itself goes to the device through pinned staging memory in one
  copy, and ``apk_gather_records`` picks the float64 blocks out of it into one contiguous device column per field --
  no ``struct.unpack`` of the payload, no Python list of floats.  The reference then drops duplicated rows with
  ``set(map(tuple, ...))`` (ecosmog.py:236-238); identical rows assign identical values to identical cells, so the
  gridded field is the same with or without that step and it is not reproduced here.
"""
from __future__ import annotations

import ctypes as ct
import struct

import numpy as np
import torch

from . import _lib
from ._lib import AstrildPkError
from .engine import _ptr, get_engine


def assign_grid(x, y, z, values, npar: int, boxsize: float = 1.0, device=None) -> torch.Tensor:
    """-> float64 device tensor [npar][npar][npar] (what ``_read_data`` returns for an .h5 file)."""
    eng = get_engine(int(npar), float(boxsize), device)
    cols = [eng._to_device(np.asarray(c) if not isinstance(c, torch.Tensor) else c).contiguous() for c in (x, y, z)]
    if not (cols[0].dtype == cols[1].dtype == cols[2].dtype):
        raise AstrildPkError("x, y, z must share one dtype")
    vals = eng._to_device(np.asarray(values) if not isinstance(values, torch.Tensor) else values).contiguous()
    n = int(cols[0].shape[0])
    if not (cols[1].shape[0] == n and cols[2].shape[0] == n and vals.shape[0] == n):
        raise AstrildPkError("x, y, z and values must be equally long")
    N = eng.N
    out = torch.empty((N, N, N), dtype=torch.float64, device=eng.device)
    winner = torch.empty(N * N * N, dtype=torch.int32, device=eng.device)
    bad = torch.zeros(1, dtype=torch.int64, device=eng.device)
    code = lambda t: _lib.APK_F32 if t.dtype == torch.float32 else _lib.APK_F64   # noqa: E731
    _lib.call("apk_assign_grid", eng._plan, _ptr(cols[0]), _ptr(cols[1]), _ptr(cols[2]), code(cols[0]), _ptr(vals),
              code(vals), n, _ptr(out), _ptr(winner), _ptr(bad), eng.stream)
    nbad = int(bad.item())
    if nbad:
        raise IndexError(f"{nbad} sample(s) fall outside the {N}^3 grid (index out of bounds, as NumPy would raise)")
    return out


def poisson_record_pieces(content, nfields: int, levelmin: int, levelmax: int, dimensions: int = 3, dst0=None):
    """Walks the record headers of one ``output_poisson`` file image exactly like ecosmog.py:184-230 and returns
    (pieces, counts): pieces is int64 [n][3] = (byte offset of a float64 block, destination element in the field's
    column, number of values) per field -- a list of ``nfields`` arrays -- and counts the values each field gained."""
    buf = memoryview(content)
    dimfac = 2 ** dimensions
    info = struct.unpack("i" * 12, buf[0:48])
    ncpu, nboundary = info[1], info[10]
    pmax = 48
    dst = [0] * nfields if dst0 is None else list(dst0)
    pieces = [[] for _ in range(nfields)]
    for _ilevel in range(levelmin, levelmax + 1):
        for _ibound in range(1, nboundary + ncpu + 1):
            pmax0 = pmax + 24
            if pmax0 > len(buf):
                raise AstrildPkError("output_poisson image ends inside a level header")
            ncache = struct.unpack("i" * 6, buf[pmax:pmax0])[4]
            if ncache == 0:
                pmax = pmax0
                continue
            for _dim in range(dimfac):
                for N in range(1, nfields + 1):
                    pmin = pmax0 + (8 * N - 4) + (N - 1) * 8 * ncache
                    pmax = pmin + ncache * 8
                    if pmax > len(buf):
                        raise AstrildPkError("output_poisson image ends inside a data record")
                    pieces[N - 1].append((pmin, dst[N - 1], ncache))
                    dst[N - 1] += ncache
                pmax0 = pmax + 4
            pmax = pmax0
    return [np.asarray(p, dtype=np.int64).reshape(-1, 3) for p in pieces], dst


def read_poisson_output(files, fields, amr_levels, dimensions: int = 3, device=None) -> dict:
    """``Ecosmog.compress_snapshot``'s columns (one float64 device tensor per entry of ``fields``) from the cpu files of
    one snapshot, given as paths or as bytes-like images, in the order the reference sorts them."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    levelmin, levelmax = min(amr_levels), max(amr_levels)
    images = []
    for f in files:
        if isinstance(f, (bytes, bytearray, memoryview)):
            images.append(bytes(f))
        else:
            with open(f, "rb") as fh:
                images.append(fh.read())
    nf = len(fields)
    dst = [0] * nf
    per_file = []
    for img in images:
        pieces, dst = poisson_record_pieces(img, nf, levelmin, levelmax, dimensions, dst)
        per_file.append(pieces)
    cols = [torch.empty(dst[j], dtype=torch.float64, device=dev) for j in range(nf)]
    lib_dev = dev.index if dev.index is not None else torch.cuda.current_device()
    stream = ct.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    keep = []
    for img, pieces in zip(images, per_file):
        n = len(img)
        stage = torch.empty(n + (-n) % 8, dtype=torch.uint8, pin_memory=True)       # pinned staging -> one async copy
        stage[:n] = torch.frombuffer(bytearray(img), dtype=torch.uint8)
        raw = stage.to(dev, non_blocking=True)
        for j in range(nf):
            if len(pieces[j]) == 0:
                continue
            # pieces of at most 2^16 values: one CTA each
            pj = pieces[j]
            parts = []
            for src, d0, cnt in pj:
                for o in range(0, int(cnt), 1 << 16):
                    c = min(1 << 16, int(cnt) - o)
                    parts.append((src + 8 * o, d0 + o, c))
            tab = torch.from_numpy(np.asarray(parts, dtype=np.int64)).to(dev, non_blocking=True)
            _lib.call("apk_gather_records", _ptr(raw), _ptr(tab), len(parts), _ptr(cols[j]), lib_dev, stream)
            keep.append(tab)
        keep.extend((stage, raw))
    torch.cuda.current_stream(dev).synchronize()
    return dict(zip(fields, cols))
