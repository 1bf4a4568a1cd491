"""Synthetic particle sets of the BASELINE.json configs, generated on the device with torch.

Input generation only -- nothing here is on the measured path.  Ramses-style output: three
float32 columns x, y, z in box units [0, 1) (the convention of
/root/reference/src/astrild/particles/ecosmog.py:139-254 and
/root/reference/src/astrild/power_spectra/power_spectrum_3d.py:144-147).
"""
from __future__ import annotations

import math

import torch


def uniform_particles(n_side: int, seed: int, device) -> tuple:
    """Config 1: n_side^3 uniform-random particles."""
    g = torch.Generator(device=device).manual_seed(seed)
    n = n_side ** 3
    return tuple(torch.rand(n, generator=g, device=device, dtype=torch.float32) for _ in range(3))


def _bbks_power(k: torch.Tensor, gamma: float = 0.21) -> torch.Tensor:
    """P_lin(k) ~ k T^2(k), BBKS transfer function, shape parameter Gamma [h/Mpc]."""
    q = (k / gamma).clamp_min(1e-12)
    t = torch.log1p(2.34 * q) / (2.34 * q) * (1 + 3.89 * q + (16.1 * q) ** 2 + (5.46 * q) ** 3 + (6.71 * q) ** 4) ** -0.25
    return k * t * t


def zeldovich_particles(n_side: int, boxsize: float, seed: int, device, rms_cells: float = 1.0,
                        x_planes: tuple | None = None) -> tuple:
    """Configs 2/3: lattice q = (i + 1/2)/n displaced by a Zel'dovich field whose rms 1-D
    displacement is ``rms_cells`` lattice spacings.  ``x_planes=(a, b)`` returns only the particles
    whose LATTICE plane is in [a, b) (per-rank generation); the field itself is global so every
    rank draws the same realisation.  Returns (x, y, z) float32 in [0, 1).
    """
    n = n_side
    g = torch.Generator(device=device).manual_seed(seed)
    noise = torch.randn((n, n, n), generator=g, device=device, dtype=torch.float32)
    dk = torch.fft.rfftn(noise)
    del noise
    kf = 2 * math.pi / boxsize
    fx = torch.fft.fftfreq(n, d=1.0 / n, device=device) * kf
    fz = torch.fft.rfftfreq(n, d=1.0 / n, device=device) * kf
    k2 = fx[:, None, None] ** 2 + fx[None, :, None] ** 2 + fz[None, None, :] ** 2
    k2[0, 0, 0] = 1.0
    dk *= torch.sqrt(_bbks_power(torch.sqrt(k2))) / k2       # delta_k / k^2
    dk[0, 0, 0] = 0
    del k2
    a, b = (0, n) if x_planes is None else x_planes
    cols, amp = [], None
    lattice = (torch.arange(n, device=device, dtype=torch.float32) + 0.5) / n
    for axis, kvec in enumerate((fx[:, None, None], fx[None, :, None], fz[None, None, :])):
        psi = torch.fft.irfftn(dk * (1j * kvec), s=(n, n, n))            # displacement along `axis`
        if amp is None:
            amp = rms_cells / n / psi.std().item()
        psi = psi[a:b]
        shape = [1, 1, 1]
        shape[axis] = -1
        q = lattice[a:b] if axis == 0 else lattice
        pos = (q.reshape(shape) + psi * amp).reshape(-1)
        pos = pos - torch.floor(pos)
        pos = torch.where(pos >= 1.0, torch.zeros_like(pos), pos)       # float32 rounding of 1 - eps
        cols.append(pos.contiguous())
        del psi
    return tuple(cols)


def sine_displaced_particles(n_side: int, seed: int, device, rms_cells: float = 1.0, x_planes: tuple | None = None) -> tuple:
    """Config 5 (2048^3, per-rank generation): lattice q = (i + 1/2)/n displaced by a smooth analytic field
    (a few plane waves per axis, phases from ``seed``) with rms 1-D displacement ``rms_cells`` lattice
    spacings.  No global FFT is needed, so each rank builds only its own planes.  Returns (x, y, z) float32
    in [0, 1)."""
    n = n_side
    a, b = (0, n) if x_planes is None else x_planes
    g = torch.Generator(device="cpu").manual_seed(seed)
    nwave = 6
    kv = torch.randint(1, 9, (3, nwave, 3), generator=g).to(device=device, dtype=torch.float32)   # wave vectors
    ph = (torch.rand((3, nwave), generator=g) * 2 * math.pi).to(device)
    amp = rms_cells / n * math.sqrt(2.0 / nwave)
    lat = (torch.arange(n, device=device, dtype=torch.float32) + 0.5) / n
    qx, qy, qz = lat[a:b, None, None], lat[None, :, None], lat[None, None, :]
    cols = []
    for axis in range(3):
        disp = torch.zeros((b - a, n, n), device=device, dtype=torch.float32)
        for w in range(nwave):
            disp += torch.sin(2 * math.pi * (kv[axis, w, 0] * qx + kv[axis, w, 1] * qy + kv[axis, w, 2] * qz) + ph[axis, w])
        q = (qx, qy, qz)[axis]
        pos = (q + amp * disp).reshape(-1)
        del disp
        pos = pos - torch.floor(pos)
        pos = torch.where(pos >= 1.0, torch.zeros_like(pos), pos)
        cols.append(pos.contiguous())
    return tuple(cols)


def halo_subset(pos: tuple, n_halos: int, seed: int, mu: float = math.log(1e13), sigma: float = 1.0) -> tuple:
    """Config 4: ``n_halos`` positions drawn (with replacement) from the particle columns ``pos`` and log-normal
    masses exp(N(mu, sigma)) -- a mass-weighted tracer of the same density field.  Returns (x, y, z, mass),
    float32, on the device of ``pos``."""
    device = pos[0].device
    g = torch.Generator(device=device).manual_seed(seed)
    idx = torch.randint(0, pos[0].numel(), (n_halos,), generator=g, device=device)
    mass = torch.exp(mu + sigma * torch.randn(n_halos, generator=g, device=device, dtype=torch.float32))
    return tuple(c[idx].contiguous() for c in pos) + (mass.contiguous(),)
