"""ctypes front-end of the oracle's C twin (oracle/pk_oracle_c.c)  --  TEST INFRASTRUCTURE ONLY.

Same semantics as oracle/pk_oracle.py (PARITY UNPINNED, see there), usable at 512^3+.
FFT stays ``scipy.fft.rfftn`` in float64 (the reference runs FFTW in float64; neither FFTW
nor pfft exists in this image).  Importers: tests/, __graft_entry__.smoke(), bench.py's
cpu_baseline / --impl reference legs.  Never imported by astrild_b200/.
"""
from __future__ import annotations

import ctypes as ct
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import pk_oracle as _o

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpk_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or (
            os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "pk_oracle_c.c"))):
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ct.CDLL(_SO)
        _lib.orc_paint.restype = None
        _lib.orc_paint.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int, ct.c_int,
                                   ct.c_void_p, ct.c_int, ct.c_int64, ct.c_int, ct.c_double,
                                   ct.c_int, ct.c_double, ct.c_void_p]
        _lib.orc_paint_yblocks.restype = None
        _lib.orc_paint_yblocks.argtypes = _lib.orc_paint.argtypes + [ct.c_int, ct.c_int, ct.c_int]
        _lib.orc_bin_power.restype = None
        _lib.orc_bin_power.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_int] + [ct.c_void_p] * 4 + [
            ct.c_int, ct.c_double] + [ct.c_void_p] * 4 + [ct.c_int, ct.c_int]
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ct.c_void_p)


def paint(pos, mass, N: int, L: float, resampler: str = "cic", shift: float = 0.0,
          out: np.ndarray | None = None, threads: int = 1) -> np.ndarray:
    """Scalar-loop twin of pk_oracle.paint.  pos: (Np,3) AoS or tuple (x,y,z) SoA; f32 or f64.
    threads > 1: workers own blocks of y-rows, two phases (orc_paint_yblocks); deterministic, equal to the
    single-threaded canvas up to the order of float64 additions next to block boundaries."""
    support = _o.RESAMPLER_SUPPORT[resampler]
    if isinstance(pos, (tuple, list)):
        x, y, z = (np.ascontiguousarray(a) for a in pos)
        assert x.dtype == y.dtype == z.dtype and x.dtype in (np.float32, np.float64)
        ptrs, layout, is32, Np = (_ptr(x), _ptr(y), _ptr(z)), 1, x.dtype == np.float32, x.shape[0]
        keep = (x, y, z)
    else:
        p = np.ascontiguousarray(pos)
        if p.dtype not in (np.float32, np.float64):
            p = p.astype(np.float64)
        ptrs, layout, is32, Np = (_ptr(p), None, None), 0, p.dtype == np.float32, p.shape[0]
        keep = (p,)
    m = None
    if mass is not None and not np.isscalar(mass):
        m = np.ascontiguousarray(mass)
        if m.dtype not in (np.float32, np.float64):
            m = m.astype(np.float64)
    canvas = np.zeros((N, N, N), dtype=np.float64) if out is None else out
    args = (ptrs[0], ptrs[1], ptrs[2], layout, int(is32), _ptr(m),
            int(m is not None and m.dtype == np.float32), Np, N, float(L), support, float(shift), _ptr(canvas))
    threads = max(1, min(int(threads), N // (2 * max(support, 2))))     # blocks of at least `support` rows
    if threads == 1:
        lib().orc_paint(*args)
    else:
        fn = lib().orc_paint_yblocks
        with ThreadPoolExecutor(threads) as ex:
            for phase in (0, 1):
                list(ex.map(lambda t: fn(*args, t, threads, phase), range(threads)))
    if mass is not None and np.isscalar(mass) and mass != 1.0:
        canvas *= mass
    del keep
    return canvas


def fftpower_1d(c1, c2, N: int, L: float, kmin: float = 0.0, dk=None, kmax=None,
                k_dtype=np.float64, threads: int = 1):
    """Twin of pk_oracle.fftpower_1d (x-slabs optionally spread over threads)."""
    c1 = np.ascontiguousarray(c1, dtype=np.complex128)
    c2 = None if c2 is None else np.ascontiguousarray(c2, dtype=np.complex128)
    kx, ky, kz = _o.k_tables(N, L, k_dtype)
    edges = _o.k_edges(N, L, kmin, dk, kmax)
    e2 = edges ** 2
    nb = len(edges) + 1

    def work(rng):
        xs, yr, yi = np.zeros(nb), np.zeros(nb), np.zeros(nb)
        ns = np.zeros(nb, dtype=np.int64)
        lib().orc_bin_power(_ptr(c1), _ptr(c2), N, _ptr(kx), _ptr(ky), _ptr(kz), _ptr(e2),
                            len(edges), float(L) ** 3, _ptr(xs), _ptr(yr), _ptr(yi), _ptr(ns),
                            rng[0], rng[1])
        return xs, yr, yi, ns

    threads = max(1, min(threads, N))
    bounds = np.linspace(0, N, threads + 1).astype(int)
    ranges = list(zip(bounds[:-1], bounds[1:]))
    if threads == 1:
        parts = [work(ranges[0])]
    else:
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(work, ranges))
    xsum = sum(p[0] for p in parts)
    ysum = sum(p[1] for p in parts) + 1j * sum(p[2] for p in parts)
    Nsum = sum(p[3] for p in parts)
    with np.errstate(invalid="ignore", divide="ignore"):
        k = (xsum / Nsum)[1:-1]
        power = (ysum / Nsum)[1:-1]
    return {"k": k, "power": power, "modes": Nsum[1:-1].copy(), "edges": edges,
            "xsum": xsum, "ysum": ysum, "Nsum": Nsum}


def power_from_mesh(value_map1, value_map2, L: float, workers: int = 1, threads: int = 1):
    """Twin of pk_oracle.power_from_mesh (power_spectrum_3d.py:164-226)."""
    v1 = np.asarray(value_map1)
    N = v1.shape[0]
    c1 = _o.r2c(v1, workers)
    c2 = None if value_map2 is None else _o.r2c(np.asarray(value_map2), workers)
    r = fftpower_1d(c1, c2, N, L, kmin=2 * np.pi / L, threads=threads)
    return r["k"], r["power"].real - 0.0, r["modes"]


def _combine_compensate_inplace(c, c2, N: int, L: float, resampler: str, interlaced: bool, compensated: bool):
    """pk_oracle.interlace_combine followed by pk_oracle.compensate, x-slab by x-slab and in place (same
    expressions in the same order per element), so that a 1024^3 field needs no third and fourth full-size
    temporary.  c2 is consumed."""
    kx, ky, kz = _o.k_tables(N, L)
    H = L / N
    f = _o.compensation_1d(resampler, interlaced, N) if compensated else None
    for ix in range(N):
        if interlaced:
            kH = (kx[ix] + ky[:, None] + kz[None, :]) * H
            c[ix] = 0.5 * c[ix] + 0.5 * c2[ix] * np.exp(0.5j * kH)
        if compensated:
            c[ix] = c[ix] / (f[ix] * f[:, None] * f[None, : N // 2 + 1])
    return c


def power_from_particles(pos, mass, N: int, L: float, resampler: str = "tsc",
                         interlaced: bool = False, compensated: bool = False,
                         normalize: bool = False, workers: int = 1, threads: int = 1,
                         pos2=None, mass2=None, timings: dict | None = None, paint_L: float | None = None,
                         lean: bool = False):
    """Twin of pk_oracle.power_from_particles (stats_subfind.py:125-150 + CatalogMesh options).

    paint_L: box length in the units of ``pos`` when they differ from L (Ramses-style [0,1) coordinates:
    paint_L = 1), so that the float32 positions reach the window arithmetic unrescaled.
    lean: free every full-size array as soon as it is dead and combine / compensate in place (BASELINE sizes)."""
    import time

    t = {"deposit": 0.0, "fft": 0.0, "bin": 0.0}
    pL = L if paint_L is None else paint_L

    def field(p, m):
        dx = L / N
        t0 = time.perf_counter()
        real = paint(p, m, N, pL, resampler, threads=threads)
        t["deposit"] += time.perf_counter() - t0
        scale = (N ** 3 / real.sum()) if normalize else 1.0 / dx ** 3
        t0 = time.perf_counter()
        c = _o.r2c(real, workers)
        t["fft"] += time.perf_counter() - t0
        del real
        c *= scale
        c2_ = None
        if interlaced:
            t0 = time.perf_counter()
            real2 = paint(p, m, N, pL, resampler, shift=0.5, threads=threads)
            t["deposit"] += time.perf_counter() - t0
            t0 = time.perf_counter()
            c2_ = _o.r2c(real2, workers)
            t["fft"] += time.perf_counter() - t0
            del real2
            c2_ *= scale
        t0 = time.perf_counter()
        if lean:
            c = _combine_compensate_inplace(c, c2_, N, L, resampler, interlaced, compensated)
        else:
            if interlaced:
                c = _o.interlace_combine(c, c2_, N, L)
            if compensated:
                c = _o.compensate(c, resampler, interlaced, N)
        t["bin"] += time.perf_counter() - t0
        return c

    c1 = field(pos, mass)
    c2 = None if pos2 is None else field(pos2, mass2)
    t0 = time.perf_counter()
    r = fftpower_1d(c1, c2, N, L, kmin=2 * np.pi / L, threads=threads)
    t["bin"] += time.perf_counter() - t0
    if timings is not None:
        timings.update(t)
    return r["k"], r["power"].real - 0.0, r["modes"]
