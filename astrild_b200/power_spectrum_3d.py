"""Drop-in for ``astrild.power_spectra.power_spectrum_3d.PowerSpectrum3D``.

Same class, methods, arguments and return values as
/root/reference/src/astrild/power_spectra/power_spectrum_3d.py:18-249; the nbodykit calls in
``_power_spectrum_3d`` (:164-226) run on the B200 through astrild_b200.lab.  Differences,
all outside the numerical path:
  * ``PowerSpectrumWarning`` (undefined name in the reference, :49,100,127) is raised as
    ``PowerSpectrum3DWarning``;
  * the snapshot-consistency assert (:109) compares with ``nansum`` so empty bins do not trip it;
  * ``_save_results`` falls back to ``.npz`` when pandas' HDF5 backend (pytables) is absent;
  * ``return_modes=True`` additionally returns Nmodes (the reference drops ``power['modes']``).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

from .lab import ArrayMesh, FFTPower


class PowerSpectrum3DWarning(BaseException):
    pass


class PowerSpectrum3D:
    """
    Attributes:
        sim_type:
        simulation: object with ``boxsize``, ``domain_level``, ``npar`` (and, for ``compute``,
            ``dir_nrs``, ``dirs``, ``get_file_paths``) -- astrild's ``Simulation``/``Ecosmog``.

    Methods:
        compute:
    """

    def __init__(self, sim_type: str, simulation, device=None):
        self.sim = simulation
        self.sim.type = sim_type
        self.device = device

    def compute(
        self,
        quantities: List[str],
        file_dsc: List[Dict[str, str]],
        snap_nrs: Optional[List[int]] = None,
        dir_out: Optional[str] = None,
        save: bool = True,
    ) -> Union[None, dict]:
        """
        Power spectrum of particle quanities.

        Args:
            quantities: [rho, phi, dphi/dt, chi, velocity, kappa, \\Delta T]
            file_dsc: {path: , root: , extention: }
        """
        descs = [dict(d) for d in file_dsc]            # the reference pops "path" from the caller's dicts
        if snap_nrs:
            assert set(snap_nrs) < set(self.sim.dir_nrs), PowerSpectrum3DWarning(
                f"Some of the snapshots {snap_nrs} do not exist" + f"in:\n{self.sim.dir_nrs}"
            )
            paths = [self.sim.get_file_paths(d, d["path"], "max") for d in descs[:2]]
        else:
            paths = []
            for i, d in enumerate(descs[:2]):
                where = d.pop("path")
                if i == 0:
                    snap_nrs = self.sim.get_file_nrs(d, where, "max")
                paths.append(self.sim.get_file_paths(d, where, "max"))

        snap_nrs = np.sort(snap_nrs)
        if len(file_dsc) > 1:
            pk = self._cross_power_spectra(quantities, snap_nrs, paths[0], paths[1])
        else:
            pk = self._auto_power_spectra(quantities, snap_nrs, paths[0])

        if save:
            self._save_results(quantities, pk)
        else:
            return pk

    def _spectra_over_snapshots(self, snap_nrs, path_lists, quantity) -> dict:
        pk = {"k": {}, "P": {}}
        for snap_nr, *files in zip(snap_nrs, *path_lists):
            maps = [self._read_data(f, quantity) for f in files]
            if any(m.ndim != 3 for m in maps):
                raise PowerSpectrum3DWarning(f"{maps[0].ndim}D is not supported :-(")
            k, Pk = self._power_spectrum_3d(*maps)
            pk["k"]["snap_%d" % snap_nr] = k
            pk["P"]["snap_%d" % snap_nr] = Pk
        if len(path_lists[0]) > 1:
            # wavenumbers of different snapshots must agree (nansum: empty bins are NaN)
            cols = list(pk["k"].keys())
            assert np.nansum(pk["k"][cols[0]]) == np.nansum(pk["k"][cols[1]])
        return pk

    def _auto_power_spectra(self, quantity: List[str], snap_nrs: np.array, _file_paths: List[str]) -> dict:
        return self._spectra_over_snapshots(snap_nrs, [_file_paths], quantity)

    def _cross_power_spectra(self, quantity: List[str], snap_nrs: np.array, _file_paths1: List[str],
                             _file_paths2: List[str]) -> dict:
        return self._spectra_over_snapshots(snap_nrs, [_file_paths1, _file_paths2], None)

    def _read_data(self, file_in: str, quantity: Optional[str] = None) -> np.ndarray:
        """ """
        value_map = np.zeros((self.sim.npar, self.sim.npar, self.sim.npar))
        if ".h5" in file_in:
            import pandas as pd

            fields = pd.read_hdf(file_in, key="df")
            x = (self.sim.npar * fields["x"].values).astype(int)
            y = (self.sim.npar * fields["y"].values).astype(int)
            z = (self.sim.npar * fields["z"].values).astype(int)
            if isinstance(quantity, (list, tuple)):
                quantity = quantity[0]
            value_map[(x, y, z)] = fields[quantity].values
        elif ".npy" in file_in:
            value_map = np.load(file_in)
        return value_map

    def _get_vector_magnitude(self, value_map: np.ndarray) -> np.ndarray:
        """ Compute vector magnitude for 3D array """
        value_map = np.sqrt(np.sum(np.square(value_map), axis=3))
        assert len(value_map.shape) == 3
        return value_map

    def _power_spectrum_3d(
        self,
        value_map1: np.ndarray,
        value_map2: Optional[np.ndarray] = None,
        return_modes: bool = False,
    ) -> Tuple[np.array, np.array]:
        """
        Compute the 3D auto or cross power spectrum.

        Args:
            value_map:
                three dimensional array containing values of interest
        Returns:
            k:
                wavenumber
            Pk:
                power at each wavenumber
        """
        _k_min = 2 * np.pi / self.sim.boxsize
        if value_map2 is None:
            _mesh1 = ArrayMesh(
                value_map1,
                Nmesh=self.sim.domain_level,
                compensated=False,
                BoxSize=self.sim.boxsize,
                device=self.device,
            )
            r = FFTPower(_mesh1, mode="1d", kmin=_k_min)
        else:
            _mesh1 = ArrayMesh(
                value_map1,
                Nmesh=self.sim.domain_level,
                compensated=True,
                interlaced=True,
                window="TSC",
                BoxSize=self.sim.boxsize,
                device=self.device,
            )
            _mesh2 = ArrayMesh(
                value_map2,
                Nmesh=self.sim.domain_level,
                compensated=True,
                interlaced=True,
                window="TSC",
                BoxSize=self.sim.boxsize,
                device=self.device,
            )
            r = FFTPower(first=_mesh1, mode="1d", second=_mesh2, kmin=_k_min)
        k = np.array(r.power["k"])
        Pk = np.array(r.power["power"].real - r.power.attrs["shotnoise"])
        print("Pk wavenumber ------>", k.min(), k.max())
        if return_modes:
            return k, Pk, np.array(r.power["modes"])
        return k, Pk

    def _save_results(self, quantity: List[str], pk: dict) -> None:
        """
        Save results each power spectrum of each simulations snapshot

        Args:
            quantity:
                Quantity of whicht the power spectrum was calculated,
                e.g. divergence velocity, matter, Phi, ...
            pk:
                Simulation power spectra for different snapshots/redshifts.
        """
        import pandas as pd

        _columns = list(pk["k"].keys())
        df = pd.DataFrame(data=pk["P"], index=pk["k"][_columns[0]])
        filename = self.sim.dirs["out"] + "pk_%s.h5" % (("_").join(quantity))
        if os.path.exists(filename):
            os.remove(filename)
        print(f"Saving results to -> {filename}")
        try:
            df.to_hdf(filename, key="df", mode="w")
        except ImportError:  # pytables absent: same table as .npz
            np.savez(filename[:-3] + ".npz", k=df.index.values, columns=np.array(_columns), P=df.values)
