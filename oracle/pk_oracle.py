"""CPU oracle for astrild's matter/halo P(k) hot path  --  TEST INFRASTRUCTURE ONLY.

This module is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  Nothing under ``astrild_b200/`` imports it.

PARITY UNPINNED.  The arithmetic of this path lives in third-party packages that
astrild pins but does not vendor (``/root/reference/poetry.lock:333-336`` nbodykit
0.3.14, ``:478-481`` pmesh 0.1.55, ``:449-452`` pfft-python 0.1.21) and that cannot be
installed offline; astrild's own tests never exercise the path.
The oracle therefore *restates* the published algorithms of those packages and is
anchored on astrild's call sites

  * ``src/astrild/particles/hutils/stats_subfind.py:125-150``  (paint tsc -> /dx^3 ->
    ArrayMesh -> FFTPower(mode="1d", kmin=2pi/L) -> power.real - shotnoise)
  * ``src/astrild/power_spectra/power_spectrum_3d.py:164-226`` (ArrayMesh -> FFTPower,
    auto and cross)

and on analytic known answers (tests/test_oracle_kat.py).

Every function is plain NumPy, float64 throughout (what the reference runs in).
"""
from __future__ import annotations

import numpy as np

try:  # scipy's pocketfft can use several workers; numpy's cannot
    import scipy.fft as _fft
except Exception:  # pragma: no cover
    _fft = None

RESAMPLER_SUPPORT = {"nearest": 1, "ngp": 1, "cic": 2, "linear": 2, "tsc": 3, "quadratic": 3}


# --------------------------------------------------------------------------------------
# pmesh.pm.ParticleMesh.paint   (call site: stats_subfind.py:130-131)
# --------------------------------------------------------------------------------------
def window_1d(resampler: str, g: np.ndarray):
    """Per-axis base index and weights of pmesh's window kernels.

    Restates pmesh 0.1.55 ``window.py`` / ``_window_imp.c``:
    support s; ``ipos = floor(g + shift) - left`` with ``left=(s-1)//2`` and
    ``shift = 0.5`` for odd s; weights ``W(dx - j)``, renormalised to sum 1.
    ``g`` is the position in grid units (cell i is centred on g = i).
    Returns (ipos int64 [n], w float64 [n, s]).
    """
    s = RESAMPLER_SUPPORT[resampler]
    left = (s - 1) // 2
    shift = 0.5 if (s % 2) else 0.0
    ipos = np.floor(g + shift).astype(np.int64) - left
    dx = g - ipos
    w = np.empty(g.shape + (s,), dtype=np.float64)
    for j in range(s):
        x = np.abs(dx - j)
        if s == 1:
            w[..., j] = 1.0
        elif s == 2:
            w[..., j] = np.maximum(0.0, 1.0 - x)
        else:
            w[..., j] = np.where(x <= 0.5, 0.75 - x * x, 0.5 * np.square(np.maximum(0.0, 1.5 - x)))
    w /= w.sum(axis=-1, keepdims=True)
    return ipos, w


def paint(pos, mass, N: int, L: float, resampler: str = "cic", shift: float = 0.0,
          out: np.ndarray | None = None) -> np.ndarray:
    """``pm.paint(pos, mass=mass, resampler=...)`` on one rank (stats_subfind.py:130-131).

    pos: (Np, 3) in the same units as L; mass: scalar or (Np,).  ``shift`` is pmesh's
    ``affine.shift`` in grid units (0.5 for the interlaced twin).
    Returns the float64 (N, N, N) canvas holding mass per cell (not density).
    """
    pos = np.asarray(pos, dtype=np.float64)
    Np = pos.shape[0]
    mass = np.broadcast_to(np.asarray(mass, dtype=np.float64), (Np,))
    canvas = np.zeros(N * N * N, dtype=np.float64) if out is None else out.reshape(-1)
    s = RESAMPLER_SUPPORT[resampler]
    chunk = 1 << 20
    for a in range(0, Np, chunk):
        p = pos[a:a + chunk]
        g = p * (N / L) + shift
        ix, wx = window_1d(resampler, g[:, 0])
        iy, wy = window_1d(resampler, g[:, 1])
        iz, wz = window_1d(resampler, g[:, 2])
        m = mass[a:a + chunk]
        for jx in range(s):
            cx = np.mod(ix + jx, N) * (N * N)
            for jy in range(s):
                cxy = cx + np.mod(iy + jy, N) * N
                wxy = m * wx[:, jx] * wy[:, jy]
                for jz in range(s):
                    idx = cxy + np.mod(iz + jz, N)
                    canvas += np.bincount(idx, weights=wxy * wz[:, jz], minlength=N * N * N)
    return canvas.reshape(N, N, N)


# --------------------------------------------------------------------------------------
# pmesh RealField.r2c and the k tables  
# --------------------------------------------------------------------------------------
def freq_index(N: int) -> np.ndarray:
    """Integer lattice frequency of storage index i: i for 2i < N, i - N otherwise.

    For even N the Nyquist index N/2 maps to -N/2 (pmesh convention).
    """
    i = np.arange(N, dtype=np.int64)
    return np.where(2 * i < N, i, i - N)


def k_tables(N: int, L: float, k_dtype=np.float64):
    """Per-axis wavenumber tables ``k_i = w_i * N / L`` with ``w_i = n_i * (2 pi / N)``.

    The expression order is pmesh's.  ``k_dtype`` is the dtype of the
    index ramp (Appendix C, Q1); default float64.  Returns float64 arrays
    (kx[N], ky[N], kz[N//2+1]) whose values were computed in ``k_dtype``.
    """
    n = freq_index(N).astype(k_dtype)
    w = n * k_dtype(2 * np.pi / N)
    k = (w * k_dtype(N) / k_dtype(L)).astype(np.float64)
    return k, k.copy(), k[: N // 2 + 1].copy()


def r2c(field: np.ndarray, workers: int = 1) -> np.ndarray:
    """``RealField.r2c()``: forward FFT normalised by 1/N^3, Hermitian half on the last axis."""
    f = np.asarray(field, dtype=np.float64)
    if _fft is not None:
        c = _fft.rfftn(f, workers=workers)
    else:  # pragma: no cover
        c = np.fft.rfftn(f)
    c *= 1.0 / f.size
    return c


# --------------------------------------------------------------------------------------
# nbodykit CatalogMesh semantics: interlacing and window compensation  
# --------------------------------------------------------------------------------------
def interlace_combine(c1: np.ndarray, c2: np.ndarray, N: int, L: float) -> np.ndarray:
    """``c = 0.5 c1 + 0.5 c2 exp(0.5j * sum_i k_i H_i)`` with H = L/N."""
    kx, ky, kz = k_tables(N, L)
    H = L / N
    kH = (kx[:, None, None] + ky[None, :, None] + kz[None, None, :]) * H
    return 0.5 * c1 + 0.5 * c2 * np.exp(0.5j * kH)


def compensation_1d(resampler: str, interlaced: bool, N: int) -> np.ndarray:
    """Per-axis factor the complex field is *divided* by (nbodykit ``Compensate*``).

    interlaced: ``sinc(w/2pi)^p`` (p = 2 CIC, 3 TSC);  not interlaced: the shot-noise
    variants ``(1 - 2/3 s)^(1/2)`` (CIC) and ``(1 - s + 2/15 s^2)^(1/2)`` (TSC),
    ``s = sin^2(w/2)`` (Jing 2005 eqs 18, 20).  ``w`` is the circular frequency.
    """
    w = freq_index(N).astype(np.float64) * (2 * np.pi / N)
    p = {"cic": 2, "tsc": 3}[resampler]
    if interlaced:
        return np.sinc(w / (2 * np.pi)) ** p
    s = np.sin(0.5 * w) ** 2
    if resampler == "cic":
        return (1 - 2.0 / 3 * s) ** 0.5
    return (1 - s + 2.0 / 15 * s * s) ** 0.5


def compensate(c: np.ndarray, resampler: str, interlaced: bool, N: int) -> np.ndarray:
    f = compensation_1d(resampler, interlaced, N)
    return c / (f[:, None, None] * f[None, :, None] * f[None, None, : N // 2 + 1])


# --------------------------------------------------------------------------------------
# nbodykit FFTPower(mode="1d")  (call sites power_spectrum_3d.py:189-195, 216-222;
# stats_subfind.py:142-148; semantics SURVEY.md A.5, A.6)
# --------------------------------------------------------------------------------------
def k_edges(N: int, L: float, kmin: float = 0.0, dk: float | None = None,
            kmax: float | None = None) -> np.ndarray:
    """``kedges = numpy.arange(kmin, kmax, dk)``; dk defaults to 2pi/L, kmax to pi N/L + dk/2."""
    if dk is None:
        dk = 2 * np.pi / L
    if kmax is None:
        kmax = np.pi * N / L + dk / 2
    return np.arange(kmin, kmax, dk)


def hermitian_weights(N: int) -> np.ndarray:
    """Weight of stored mode iz on the half axis: 2 where k_z > 0, else 1 (iz = 0, Nyquist)."""
    nz = freq_index(N)[: N // 2 + 1]
    return np.where(nz > 0, 2.0, 1.0)


def project_to_basis_1d(p3d: np.ndarray, N: int, L: float, kedges: np.ndarray,
                        k_dtype=np.float64):
    """``project_to_basis`` with one mu bin and no poles.

    Per x-slab: ``k2 = ((kx^2 + ky^2) + kz^2)``, ``dig = digitize(k2, kedges^2)``,
    ``xsum += w sqrt(k2)``, ``Nsum += w``, ``ysum.real += w Re p`` and ``ysum.imag += Im p``
    on singular planes only.  Returns full (Nx+2)-long arrays (under/overflow included):
    xsum f8, ysum c16, Nsum i8.
    """
    kx, ky, kz = k_tables(N, L, k_dtype)
    x2edges = kedges ** 2
    Nx = len(kedges) - 1
    xsum = np.zeros(Nx + 2)
    ysum = np.zeros(Nx + 2, dtype=np.complex128)
    Nsum = np.zeros(Nx + 2, dtype=np.int64)
    w = hermitian_weights(N)
    nonsing = w > 1.0
    kyz2 = None
    for ix in range(N):
        k2 = (kx[ix] ** 2 + (ky ** 2)[:, None]) + (kz ** 2)[None, :]
        dig = np.digitize(k2.ravel(), x2edges)
        wk = np.broadcast_to(w[None, :], k2.shape).ravel()
        xsum += np.bincount(dig, weights=np.sqrt(k2).ravel() * wk, minlength=Nx + 2)
        Nsum += np.bincount(dig, weights=wk, minlength=Nx + 2).astype(np.int64)
        y = np.array(p3d[ix], dtype=np.complex128)
        y.real[:, nonsing] *= 2.0
        y.imag[:, nonsing] = 0.0
        ysum.real += np.bincount(dig, weights=y.real.ravel(), minlength=Nx + 2)
        ysum.imag += np.bincount(dig, weights=y.imag.ravel(), minlength=Nx + 2)
    del kyz2
    return xsum, ysum, Nsum


def fftpower_1d(c1: np.ndarray, c2: np.ndarray | None, N: int, L: float,
                kmin: float = 0.0, dk: float | None = None, kmax: float | None = None,
                k_dtype=np.float64):
    """``FFTPower(first, mode='1d', second=, kmin=, dk=, kmax=)`` on complex fields.

    ``p3d = c1 conj(c2)``, DC zeroed, times V = L^3, projected on |k| shells.
    Returns dict(k, power (complex), modes (int64), edges, full sums) with empty bins NaN.
    """
    if c2 is None:
        c2 = c1
    p3d = c1 * np.conj(c2)
    p3d[0, 0, 0] = 0.0
    p3d *= float(L) ** 3
    edges = k_edges(N, L, kmin, dk, kmax)
    xsum, ysum, Nsum = project_to_basis_1d(p3d, N, L, edges, k_dtype)
    with np.errstate(invalid="ignore", divide="ignore"):
        k = (xsum / Nsum)[1:-1]
        power = (ysum / Nsum)[1:-1]
    return {"k": k, "power": power, "modes": Nsum[1:-1].copy(), "edges": edges,
            "xsum": xsum, "ysum": ysum, "Nsum": Nsum}


# --------------------------------------------------------------------------------------
# nbodykit FFTPower(mode="2d", Nmu=, poles=, los=)  -- SURVEY.md section 8f row N4 (astrild hints at redshift-space use:
# /root/reference/README.md:11, src/astrild/particles/hutils/tpcf.py:12-60).  Restated from nbodykit 0.3.14
# algorithms/fftpower.py (FFTPower.run: muedges = linspace(0, 1, Nmu + 1); project_to_basis) -- recollection, unpinned.
# --------------------------------------------------------------------------------------
def legendre(ell: int, mu):
    """P_ell(mu) for ell = 0..8 by Bonnet's recursion (what scipy.special.legendre(ell)(mu) evaluates)."""
    mu = np.asarray(mu, dtype=np.float64)
    p0, p1 = np.ones_like(mu), mu
    if ell == 0:
        return p0
    for n in range(1, ell):
        p0, p1 = p1, ((2 * n + 1) * mu * p1 - n * p0) / (n + 1)
    return p1


def project_to_basis_2d(p3d: np.ndarray, N: int, L: float, kedges: np.ndarray, Nmu: int, poles=(),
                        los=(0.0, 0.0, 1.0), k_dtype=np.float64):
    """``project_to_basis(y3d, [kedges, muedges], poles=, los=)``.

    mu = (k . los) / |k| (0 at k = 0); ``dig_mu = digitize(|mu|, linspace(0, 1, Nmu + 1))``; Hermitian weights 2 on the
    non-singular planes; for each ell of [0] + poles: ``(2 ell + 1) L_ell(mu) y`` with, on non-singular modes, the real
    part doubled and the imaginary part dropped for even ell and the reverse for odd ell (the conjugate mode has -mu);
    the internal mu == 1 column is folded into the last visible one.  Returns the FULL arrays (under/overflow kept):
    xsum, musum f8 [Nx+2][Nmu+2], ysum c16 [Nell][Nx+2][Nmu+2], Nsum i8, and the list of ells (0 first).
    """
    kx, ky, kz = k_tables(N, L, k_dtype)
    muedges = np.linspace(0.0, 1.0, Nmu + 1)
    x2edges = kedges ** 2
    Nx = len(kedges) - 1
    ells = sorted(set([0] + [int(e) for e in poles]))
    shape = (Nx + 2, Nmu + 2)
    xsum, musum = np.zeros(shape), np.zeros(shape)
    ysum = np.zeros((len(ells),) + shape, dtype=np.complex128)
    Nsum = np.zeros(shape, dtype=np.int64)
    w = hermitian_weights(N)
    nonsing = w > 1.0
    size = shape[0] * shape[1]
    for ix in range(N):
        k2 = (kx[ix] ** 2 + (ky ** 2)[:, None]) + (kz ** 2)[None, :]
        dig_x = np.digitize(k2.ravel(), x2edges)
        knorm = np.sqrt(k2)
        kdotl = (kx[ix] * los[0] + (ky * los[1])[:, None]) + (kz * los[2])[None, :]
        with np.errstate(invalid="ignore", divide="ignore"):
            mu = kdotl / knorm
        mu[knorm == 0.0] = 0.0
        dig_mu = np.digitize(np.abs(mu).ravel(), muedges)
        multi = np.ravel_multi_index([dig_x, dig_mu], shape)
        wk = np.broadcast_to(w[None, :], k2.shape)
        xsum.flat += np.bincount(multi, weights=(knorm * wk).ravel(), minlength=size)
        musum.flat += np.bincount(multi, weights=(np.abs(mu) * wk).ravel(), minlength=size)
        Nsum.flat += np.bincount(multi, weights=wk.ravel(), minlength=size).astype(np.int64)
        y = np.asarray(p3d[ix], dtype=np.complex128)
        for i, ell in enumerate(ells):
            wy = legendre(ell, mu) * y
            if ell % 2:
                wy.real[:, nonsing] = 0.0
                wy.imag[:, nonsing] *= 2.0
            else:
                wy.real[:, nonsing] *= 2.0
                wy.imag[:, nonsing] = 0.0
            wy *= 2.0 * ell + 1.0
            ysum[i].real.flat += np.bincount(multi, weights=wy.real.ravel(), minlength=size)
            ysum[i].imag.flat += np.bincount(multi, weights=wy.imag.ravel(), minlength=size)
    ysum[..., -2] += ysum[..., -1]
    musum[:, -2] += musum[:, -1]
    xsum[:, -2] += xsum[:, -1]
    Nsum[:, -2] += Nsum[:, -1]
    return xsum, musum, ysum, Nsum, ells


def fftpower_2d(c1: np.ndarray, c2: np.ndarray | None, N: int, L: float, Nmu: int = 5, poles=(), los=(0.0, 0.0, 1.0),
                kmin: float = 0.0, dk: float | None = None, kmax: float | None = None, k_dtype=np.float64):
    """``FFTPower(first, mode='2d', Nmu=, poles=, los=, second=, kmin=, dk=, kmax=)`` on complex fields -> dict with
    ``k``, ``mu``, ``power`` (complex), ``modes`` of shape (Nk, Nmu) and, if poles: ``poles`` = {"k", "modes",
    "power_<ell>"} (1-D over k: sums over the mu bins)."""
    if c2 is None:
        c2 = c1
    p3d = c1 * np.conj(c2)
    p3d[0, 0, 0] = 0.0
    p3d *= float(L) ** 3
    edges = k_edges(N, L, kmin, dk, kmax)
    xsum, musum, ysum, Nsum, ells = project_to_basis_2d(p3d, N, L, edges, Nmu, poles, los, k_dtype)
    sl = slice(1, -1)
    with np.errstate(invalid="ignore", divide="ignore"):
        out = {"k": (xsum / Nsum)[sl, sl], "mu": (musum / Nsum)[sl, sl], "power": (ysum[0] / Nsum)[sl, sl],
               "modes": Nsum[sl, sl].copy(), "edges": edges, "muedges": np.linspace(0.0, 1.0, Nmu + 1)}
        if len(poles):
            n1 = Nsum[sl, sl].sum(axis=-1)
            pol = {"k": xsum[sl, sl].sum(axis=-1) / n1, "modes": n1}
            for ell in poles:
                pol["power_%d" % ell] = ysum[ells.index(int(ell))][sl, sl].sum(axis=-1) / n1
            out["poles"] = pol
    return out


def power2d_bruteforce(delta1: np.ndarray, delta2: np.ndarray | None, L: float, kedges: np.ndarray, Nmu: int, poles=(),
                       los=(0.0, 0.0, 1.0)):
    """The same estimator WITHOUT the Hermitian half-space bookkeeping: every one of the N^3 modes of the full complex
    transform enters once with weight 1 and its own mu (known-answer check of project_to_basis_2d's doubling rules).
    Returns (power[Nk][Nmu] complex, modes, {ell: P_ell[Nk]})."""
    N = delta1.shape[0]
    c1 = np.fft.fftn(delta1) / N ** 3
    c2 = c1 if delta2 is None else np.fft.fftn(delta2) / N ** 3
    p3d = c1 * np.conj(c2) * float(L) ** 3
    p3d[0, 0, 0] = 0.0
    kx, _, _ = k_tables(N, L)
    k2 = (kx[:, None, None] ** 2 + kx[None, :, None] ** 2) + kx[None, None, :] ** 2
    kn = np.sqrt(k2)
    kd = (kx[:, None, None] * los[0] + kx[None, :, None] * los[1]) + kx[None, None, :] * los[2]
    with np.errstate(invalid="ignore", divide="ignore"):
        mu = kd / kn
    mu[kn == 0] = 0.0
    muedges = np.linspace(0.0, 1.0, Nmu + 1)
    dx = np.digitize(k2.ravel(), kedges ** 2)
    dm = np.digitize(np.abs(mu).ravel(), muedges)
    dm[dm == Nmu + 1] = Nmu
    Nx = len(kedges) - 1
    shape = (Nx + 2, Nmu + 2)
    multi = np.ravel_multi_index([dx, dm], shape)
    n = np.bincount(multi, minlength=shape[0] * shape[1]).reshape(shape)
    def binned(v):
        return (np.bincount(multi, weights=v.real.ravel(), minlength=n.size)
                + 1j * np.bincount(multi, weights=v.imag.ravel(), minlength=n.size)).reshape(shape)
    with np.errstate(invalid="ignore", divide="ignore"):
        p2 = (binned(p3d) / n)[1:-1, 1:-1]
        n1 = n[1:-1, 1:-1].sum(axis=-1)
        pol = {int(ell): binned((2 * ell + 1) * legendre(int(ell), mu) * p3d)[1:-1, 1:-1].sum(axis=-1) / n1 for ell in poles}
    return p2, n[1:-1, 1:-1], pol


# --------------------------------------------------------------------------------------
# astrild-level entry points (the two drop-in boundaries)
# --------------------------------------------------------------------------------------
def power_from_mesh(value_map1, value_map2, L: float, workers: int = 1, k_dtype=np.float64):
    """``PowerSpectrum3D._power_spectrum_3d`` (power_spectrum_3d.py:164-226).

    ArrayMesh stores compensated/interlaced/window only as metadata (SURVEY.md section 0
    item 2) so they do not appear here.  shotnoise attr is 0 for an ArrayMesh.
    Returns (k, Pk, modes).
    """
    v1 = np.asarray(value_map1)
    N = v1.shape[0]
    c1 = r2c(v1, workers)
    c2 = None if value_map2 is None else r2c(np.asarray(value_map2), workers)
    r = fftpower_1d(c1, c2, N, L, kmin=2 * np.pi / L, k_dtype=k_dtype)
    return r["k"], r["power"].real - 0.0, r["modes"]


def power_from_particles(pos, mass, N: int, L: float, resampler: str = "tsc",
                         interlaced: bool = False, compensated: bool = False,
                         normalize: bool = False, workers: int = 1, pos2=None, mass2=None,
                         k_dtype=np.float64):
    """``SubFind.power_spectrum`` (stats_subfind.py:125-150) plus the CatalogMesh options.

    Defaults reproduce astrild-as-written: TSC mass deposit, rho = paint/dx^3, no
    interlacing, no compensation, no normalisation.  ``normalize=True`` divides the field
    by its mean (1 + delta);  ``interlaced`` / ``compensated`` follow SURVEY.md A.3.
    A second catalogue (pos2, mass2) gives the cross spectrum Re(c1 conj c2).
    Returns (k, Pk, modes).
    """
    def field(p, m):
        dx = L / N
        m = 1.0 if m is None else m
        real = paint(p, m, N, L, resampler)
        scale = (N ** 3 / real.sum()) if normalize else 1.0 / dx ** 3
        c = r2c(real, workers) * scale
        if interlaced:
            real2 = paint(p, m, N, L, resampler, shift=0.5)
            c = interlace_combine(c, r2c(real2, workers) * scale, N, L)
        if compensated:
            c = compensate(c, resampler, interlaced, N)
        return c

    c1 = field(pos, mass)
    c2 = None if pos2 is None else field(pos2, mass2)
    r = fftpower_1d(c1, c2, N, L, kmin=2 * np.pi / L, k_dtype=k_dtype)
    return r["k"], r["power"].real - 0.0, r["modes"]


def mode_counts_bruteforce(N: int, L: float, kedges: np.ndarray, k_dtype=np.float64):
    """Full-lattice (all N^3 modes, no Hermitian trick) mode counts per bin, for the KATs."""
    kx, ky, _ = k_tables(N, L, k_dtype)
    kzfull = kx
    k2 = (kx[:, None, None] ** 2 + ky[None, :, None] ** 2) + kzfull[None, None, :] ** 2
    dig = np.digitize(k2.ravel(), kedges ** 2)
    return np.bincount(dig, minlength=len(kedges) + 1).astype(np.int64)
