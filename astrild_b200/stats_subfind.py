"""Drop-in for ``astrild.particles.hutils.stats_subfind.SubFind.power_spectrum``.

Reference: /root/reference/src/astrild/particles/hutils/stats_subfind.py:109-153 -- an unbound
function used as a static method and resolved BY NAME from a YAML key
(/root/reference/src/astrild/particles/halo.py:178,195-197), so the class name, the function
name and the keyword set are the contract.  The body is the reference's call sequence
(paint tsc -> / dx^3 -> ArrayMesh -> FFTPower(mode='1d', kmin=2pi/L) -> power.real - shotnoise)
against astrild_b200.lab; the painted field never leaves the GPU.

Added keywords (defaults reproduce the reference): ``resampler``, ``interlaced``,
``compensated``, ``normalize`` (the nbodykit CatalogMesh options the reference comments out at
:137-138), ``return_modes``, ``device``.
"""
from __future__ import annotations

import numpy as np

from .lab import ArrayMesh, CatalogMesh, FFTPower, ParticleMesh


class SubFind:
    def power_spectrum(
        snapshot,
        objects: str = "subhalo",
        limits: tuple = None,
        nbins: int = 512,
        boxsize: float = 500.0,
        resampler: str = "tsc",
        interlaced: bool = False,
        compensated: bool = False,
        normalize: bool = False,
        return_modes: bool = False,
        device=None,
    ):
        """
        Comput the real-space halo power spectrum

        Args:
        """
        if boxsize is None:
            boxsize = snapshot.header.boxsize / 1e3  # [Mpc/h]

        if objects == "subhalo":
            pos_field = snapshot.cat["SubhaloPos"][:] * snapshot.header.hubble / 1e3  # [Mpc/h]
            mass_field = snapshot.cat["SubhaloMass"][:] * snapshot.header.hubble / 1e10
            print(np.min(mass_field), np.max(mass_field))
        else:
            raise ValueError(f"objects={objects!r}: the reference only defines 'subhalo'")

        if interlaced or compensated or normalize:
            mesh = CatalogMesh(pos_field, boxsize, nbins, weight=mass_field, resampler=resampler,
                               interlaced=interlaced, compensated=compensated, normalize=normalize,
                               device=device)
        else:
            dx = boxsize / nbins
            pm = ParticleMesh(Nmesh=[nbins] * 3, BoxSize=boxsize, device=device)
            value_map = pm.paint(pos_field, mass=mass_field, resampler=resampler)
            value_map = value_map / dx ** 3     # stays on the device (reference: value_map.value / dx**3)
            mesh = ArrayMesh(value_map, Nmesh=nbins, compensated=False, BoxSize=boxsize)
        r = FFTPower(first=mesh, mode="1d", kmin=2 * np.pi / boxsize)
        k = np.array(r.power["k"])
        Pk = np.array(r.power["power"].real - r.power.attrs["shotnoise"])
        print("***********************************")
        print(k, Pk)
        if return_modes:
            return k, Pk, np.array(r.power["modes"])
        return k, Pk
