"""Drop-in for ``astrild.power_spectra.power_spectrum_3d.PowerSpectrum3D``.

Same class, methods, arguments and return values as
/root/reference/src/astrild/power_spectra/power_spectrum_3d.py:18-249; the nbodykit calls in
``_power_spectrum_3d`` (:164-226) run on the B200 through astrild_b200.lab.  Differences,
all outside the numerical path:
  * ``PowerSpectrumWarning`` (undefined name in the reference, :49,100,127) is raised as
    ``PowerSpectrum3DWarning``;
  * the snapshot-consistency assert (:109) compares with ``nansum`` so empty bins do not trip it;
  * ``_read_data`` grids ``.h5`` samples on the device (astrild_b200.ingest.assign_grid) and ``_save_results`` writes
    through astrild_b200.catalog (``.npz`` with the same table when pandas' HDF5 backend, pytables, is absent);
  * the reference's unused ``_get_vector_magnitude`` is not carried over;
  * ``return_modes=True`` additionally returns Nmodes (the reference drops ``power['modes']``).
"""
from __future__ import annotations

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

    def _read_data(self, file_in: str, quantity: Optional[str] = None):
        """The gridded field of one snapshot file: ``.npy`` maps as they are (DTFE output, float32 on disk); ``.h5``
        tables of AMR-cell samples (columns x, y, z in box units [0,1) and the quantity) through the device-side
        NGP assignment ``value_map[(npar*x).astype(int), ...] = values`` (ingest.assign_grid; reference:
        power_spectrum_3d.py:140-153).  The .h5 branch returns a float64 device tensor, which ArrayMesh takes as is."""
        if ".npy" in file_in:
            return np.load(file_in)
        if ".h5" in file_in:
            import pandas as pd

            from .ingest import assign_grid

            fields = pd.read_hdf(file_in, key="df")
            name = quantity[0] if isinstance(quantity, (list, tuple)) else quantity
            return assign_grid(fields["x"].values, fields["y"].values, fields["z"].values, fields[name].values,
                               self.sim.npar, device=self.device)
        return np.zeros((self.sim.npar,) * 3)

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
        """``pk_<quantities>.h5`` in ``sim.dirs["out"]``: index = k of the first snapshot, one column ``snap_<n>`` per
        snapshot (reference: power_spectrum_3d.py:228-249); ``.npz`` with the same table when pytables is absent."""
        from .catalog import save_power_spectra

        print(f"Saving results to -> {save_power_spectra(self.sim.dirs['out'], quantity, pk)}")
