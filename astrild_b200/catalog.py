"""Catalog driver and writer for the P(k) path.

What the reference does around the hot path, per simulation:
  * ``Halos.get_subfind_stats`` (/root/reference/src/astrild/particles/halo.py:157-207): for every snapshot, look the
    statistic up BY NAME on ``SubFind`` and call it with the YAML's ``args``; collect ``bins`` / ``values`` per
    ``snap_<n>``;
  * ``PowerSpectrum3D.compute`` (/root/reference/src/astrild/power_spectra/power_spectrum_3d.py:33-81): the same loop
    over snapshots for gridded fields;
  * ``_save_results`` (halo.py:499-539, power_spectrum_3d.py:228-249): one pandas table per statistic, index = bins
    (k), one column per snapshot, written with ``DataFrame.to_hdf(key="df")``;
  * ``SimulationCollection.compress_stats`` (/root/reference/src/astrild/simcoll.py:112-177) reads those tables back.
Every snapshot is a separate, blocking nbodykit run there.

Here the snapshots of a batch go through ONE plan (cuFFT plans, workspace, meshes, binning tables and staging
buffers are made once), and nothing between two snapshots waits for the device: the shell sums of snapshot i are copied
to pinned host memory asynchronously and turned into (k, P, Nmodes) later, so the host is already uploading snapshot
i + 1's first chunks (copy stream) while the GPU transforms and bins snapshot i.  File discovery and readers stay the
caller's (they are out of scope): a snapshot is handed over as arrays or as a callable that
produces them.

The table writer keeps the reference's layout.  pandas' HDF5 backend (pytables) is not in this image, so the same table
falls back to ``.npz`` (``index``, ``columns``, ``values``); ``read_table`` reads either.
"""
from __future__ import annotations

import os
from collections import deque

import numpy as np
import torch

from . import engine as _engine
from ._lib import AstrildPkError


# ------------------------------------------------------------------------------------------------------------------
# tables: index = bins, one column per snapshot
# ------------------------------------------------------------------------------------------------------------------
def write_table(filename: str, index, columns: dict) -> str:
    """``pd.DataFrame(data=columns, index=index).to_hdf(filename, key="df", mode="w")`` as the reference's
    ``_save_results`` does; without pytables the same table goes to ``<filename minus .h5>.npz``.  Returns the path."""
    names = list(columns)
    values = np.column_stack([np.asarray(columns[c]) for c in names]) if names else np.zeros((len(index), 0))
    if values.shape[0] != len(index):
        raise AstrildPkError("write_table: columns and index differ in length")
    if os.path.exists(filename):
        os.remove(filename)
    try:
        import pandas as pd

        pd.DataFrame(data={c: values[:, i] for i, c in enumerate(names)}, index=np.asarray(index)).to_hdf(
            filename, key="df", mode="w")
        return filename
    except ImportError:
        alt = (filename[:-3] if filename.endswith(".h5") else filename) + ".npz"
        np.savez(alt, index=np.asarray(index), columns=np.array(names), values=values)
        return alt


def read_table(filename: str):
    """-> (index, {column: values}) of a table written by ``write_table`` (either format)."""
    alt = (filename[:-3] if filename.endswith(".h5") else filename) + ".npz"
    if filename.endswith(".npz") or (not os.path.exists(filename) and os.path.exists(alt)):
        z = np.load(filename if filename.endswith(".npz") else alt, allow_pickle=False)
        return z["index"], {str(c): z["values"][:, i] for i, c in enumerate(z["columns"])}
    import pandas as pd

    df = pd.read_hdf(filename, key="df")
    return df.index.values, {str(c): df[c].values for c in df.columns}


def save_power_spectra(dir_out: str, quantity, pk: dict) -> str:
    """``PowerSpectrum3D._save_results``: ``pk_<quantities>.h5`` with index = k of the first snapshot, columns snap_<n>."""
    cols = list(pk["k"])
    name = os.path.join(dir_out, "pk_%s.h5" % "_".join(quantity)) if not dir_out.endswith(os.sep) else \
        dir_out + "pk_%s.h5" % "_".join(quantity)
    return write_table(name, pk["k"][cols[0]], pk["P"])


# ------------------------------------------------------------------------------------------------------------------
# statistics by name over snapshots (halo.py:157-207)
# ------------------------------------------------------------------------------------------------------------------
def subfind_stats(snapshots, statistics: dict, stats_class=None, dir_out: str | None = None, halofinder: str = "subfind") -> dict:
    """The loop of ``Halos.get_subfind_stats``: ``snapshots`` yields (snap_nr, snapshot) pairs (or is a dict of them),
    ``statistics`` is the parsed YAML -- {stat_name: {"args": {...}, ...}} -- and every statistic is resolved by name on
    ``stats_class`` (default: astrild_b200.SubFind) and called as ``fct(snapshot, **args)`` -> (bins, values).
    Results are collected as in the reference (``results["bins"|"values"]["snap_<n>"]``) and, with ``dir_out``, written
    as ``<halofinder>_<stat>_00.h5`` tables (halo.py:499-525)."""
    if stats_class is None:
        from .stats_subfind import SubFind as stats_class
    items = snapshots.items() if isinstance(snapshots, dict) else snapshots
    for name in statistics:
        statistics[name]["results"] = {"bins": {}, "values": {}}
    for snap_nr, snapshot in items:
        if snapshot is None:
            print(f"No sub- & halos found for snapshot {snap_nr}")
            continue
        for name, stg in statistics.items():
            fct = getattr(stats_class, name)
            bins, values = fct(snapshot, **stg.get("args", {}))[:2]
            if bins is not None and values is not None:
                stg["results"]["bins"]["snap_%d" % snap_nr] = bins
                stg["results"]["values"]["snap_%d" % snap_nr] = values
    if dir_out is not None:
        for name, stg in statistics.items():
            cols = list(stg["results"]["bins"])
            if cols:
                write_table(os.path.join(dir_out, f"{halofinder}_{name}_00.h5"), stg["results"]["bins"][cols[0]],
                            stg["results"]["values"])
    return statistics


# ------------------------------------------------------------------------------------------------------------------
# many particle snapshots through one plan
# ------------------------------------------------------------------------------------------------------------------
class PkBatch:
    """P(k) of many particle sets on one mesh geometry: ``submit`` queues a snapshot and returns at once,
    ``collect`` hands back ``{"k": {snap_<n>: ...}, "P": {...}, "modes": {...}, "shotnoise": {...}}`` -- ``k`` / ``P``
    in ``PowerSpectrum3D.compute``'s layout, P with the shot noise subtracted as ``power.real - attrs["shotnoise"]``
    (power_spectrum_3d.py:224).  Options are nbodykit's CatalogMesh's (lab.CatalogMesh).

    depth: snapshots whose results may be outstanding (pinned result slots); submit waits for the oldest beyond that.
    """

    def __init__(self, Nmesh: int, BoxSize: float, resampler: str = "tsc", interlaced: bool = True,
                 compensated: bool = True, normalize: bool = True, pos_scale: float | None = None, kmin: float | None = None,
                 dk: float | None = None, kmax: float | None = None, device=None, chunk_rows: int = 1 << 25,
                 method: str = "auto", depth: int = 2):
        self.eng = _engine.get_engine(int(Nmesh), float(BoxSize), device)
        eng = self.eng
        self.resampler, self.interlaced, self.normalize = str(resampler).lower(), bool(interlaced), bool(normalize)
        self.pos_scale, self.method, self.chunk_rows = pos_scale, method, int(chunk_rows)
        comp = (self.resampler, self.interlaced) if compensated else None
        self.binning = eng.binning(2 * np.pi / eng.L if kmin is None else kmin, dk, kmax, comp, self.interlaced)
        self.meshes = [eng.new_mesh() for _ in range(2 if self.interlaced else 1)]
        nb1 = len(self.binning.edges) + 1
        self._slots = [torch.empty((5, nb1), dtype=torch.float64, pin_memory=True) for _ in range(max(1, depth))]
        self._free = deque(range(len(self._slots)))
        self._pending: deque = deque()
        self._done: dict = {"k": {}, "P": {}, "modes": {}, "shotnoise": {}}

    # -- one snapshot --------------------------------------------------------------------------------------------
    def submit(self, snap_nr: int, position, weight=None) -> None:
        """position: (Np,3) / three (Np,) columns, host (pinned for asynchronous copies) or device, or a callable
        returning them (so that reading snapshot i + 1 overlaps the GPU work on snapshot i); weight: per-particle mass,
        a scalar or None."""
        if callable(position):
            got = position()                      # positions (three columns or (Np,3)), or (positions, weight)
            position, weight = got if (isinstance(got, tuple) and len(got) == 2) else (got, weight)
        eng = self.eng
        if not self._free:
            self._finish_oldest()
        slot = self._free.popleft()
        w, unit = weight, 1.0
        if w is not None and not np.isscalar(w):
            w, unit = _pow2_unit(eng, w)
        shifts = (0.0, 0.5) if self.interlaced else (0.0,)
        eng.deposit_many(position, w, self.resampler, shifts, self.pos_scale, self.method, self.chunk_rows, out=self.meshes)
        npart = _npart(position)
        dev_raw = torch.empty((5, len(self.binning.edges) + 1), dtype=torch.float64, device=eng.device)
        unit_weights = weight is None or np.isscalar(weight)
        if self.normalize and not unit_weights:
            eng.mesh_sum(self.meshes[0], out=dev_raw[4])                     # deposited mass, fetched with the shell sums
        c = eng.r2c(self.meshes[0])
        cs = eng.r2c(self.meshes[1]) if self.interlaced else None
        nb1 = dev_raw.shape[1]
        _engine._lib.call("apk_bin_power", self.binning.handle, _engine._ptr(c), _engine._ptr(cs), None, None,
                          _engine._ptr(dev_raw[0]), _engine._ptr(dev_raw[1]), _engine._ptr(dev_raw[2]),
                          _engine._ptr(dev_raw[3]), eng.stream)
        host = self._slots[slot]
        host.copy_(dev_raw, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(eng.device))
        shot = _shotnoise(eng.L ** 3, npart, weight, self.normalize)
        self._pending.append((int(snap_nr), slot, ev, npart, weight if np.isscalar(weight) else None, unit_weights, unit,
                              shot, dev_raw))
        assert nb1 == host.shape[1]

    def _finish_oldest(self) -> None:
        snap_nr, slot, ev, npart, scalar_w, unit_weights, unit, shot, _keep = self._pending.popleft()
        ev.synchronize()
        host = self._slots[slot]
        eng = self.eng
        N, L = eng.N, eng.L
        if self.normalize:                            # 1 + delta = mesh / mean; a scalar weight is already in the mesh
            total = npart * (1.0 if scalar_w is None else float(scalar_w)) if unit_weights else float(host[4, 0])
            s = N ** 3 / total
        else:                                         # rho = mass / dx^3 (astrild's SubFind.power_spectrum)
            s = unit / (L / N) ** 3
        res = eng.finish(host[:4].clone(), self.binning, L ** 3 * s * s / float(N) ** 6)
        key = "snap_%d" % snap_nr
        self._done["k"][key] = res["k"]
        self._done["P"][key] = res["power"].real - shot
        self._done["modes"][key] = res["modes"]
        self._done["shotnoise"][key] = shot
        self._free.append(slot)

    def collect(self) -> dict:
        while self._pending:
            self._finish_oldest()
        out, self._done = self._done, {"k": {}, "P": {}, "modes": {}, "shotnoise": {}}
        return out

    def run(self, snapshots) -> dict:
        """snapshots: iterable of (snap_nr, position[, weight]); position may be a callable (see submit)."""
        for item in snapshots:
            self.submit(*item)
        return self.collect()


def _npart(position) -> int:
    first = position[0] if (isinstance(position, (tuple, list)) and len(position) == 3 and not np.isscalar(position[0])) else position
    return int(first.shape[0])


def _pow2_unit(eng, w):
    """weights / 2^e with 2^e the power of two nearest to max |w| (exact scaling; see PkEngine.pow2_scaled)."""
    if isinstance(w, torch.Tensor) and w.is_cuda:
        return eng.pow2_scaled(w)
    wa = np.asarray(w)
    top = float(np.abs(wa).max()) if wa.size else 1.0
    if not np.isfinite(top) or top <= 0.0:
        return wa, 1.0
    unit = 2.0 ** round(float(np.log2(top)))
    if wa.dtype not in (np.float32, np.float64):
        wa = wa.astype(np.float64)
    return wa * wa.dtype.type(1.0 / unit), unit


def _shotnoise(V: float, n: int, w, normalize: bool) -> float:
    """nbodykit CatalogMesh attrs: V * sum(w^2) / sum(w)^2 (V / N unweighted); un-normalised field: sum(w^2) / V."""
    if n == 0:
        return 0.0
    if w is None or np.isscalar(w):
        m = 1.0 if w is None else float(w)
        return V / n if normalize else n * m * m / V
    wd = w.double() if isinstance(w, torch.Tensor) else np.asarray(w, dtype=np.float64)
    W, W2 = float(wd.sum()), float((wd * wd).sum())
    if not normalize:
        return W2 / V
    return V * W2 / (W * W) if W != 0.0 else 0.0
