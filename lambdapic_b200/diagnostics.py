"""Device-aware versions of the reference's in-memory diagnostics (callback/utils.py:240-400, callback/hdf5.py:440-500).

`ExtractSpeciesDensity` keeps the reference's recipe -- at stage `current_deposition`, inside the species loop, reduce the
guards (`sim.sync_currents()`), and take the difference of rho before and after the target species deposited -- but it
manages the mirrors itself (`needs_host = False`): the guard reduce runs on the device, then ONLY the rho array crosses
PCIe.  HDF5 output stays outside the accelerated path."""
from __future__ import annotations

import numpy as np

from .callback import _interval_triggered, _validate_interval


class ExtractSpeciesDensity:
    """Number density of one species on the global grid (`.density`, shape (nx, ny[, nz]) or its `slice`), refreshed
    every `interval`.  Single rank: with several ranks every rank fills the cells of its own patches and leaves the rest
    zero (the reference gathers on rank 0)."""
    DEFAULT_STAGE = "current_deposition"
    needs_host = False  # the callback downloads rho itself, after the guard reduce

    def __init__(self, sim, species, interval=100, slice=None):
        _validate_interval(interval)
        self.stage = self.DEFAULT_STAGE
        self.species, self.interval, self.slice = species, interval, slice
        self.prev_rho = None
        shape = (sim.nx, sim.ny) + ((sim.nz,) if sim.dimension == 3 else ())
        self._full = np.zeros(shape)
        self.density = self._full if slice is None else np.zeros(self._full[slice].shape)

    def _interior_rho(self, sim):
        sim.sync_currents()  # guards of rho folded into the owning cells (callback/hdf5.py:463-470)
        br = sim.bridge
        if br.resident:
            br.download({"rho"})
        inner = tuple(np.s_[:n] for n in ((sim.nx_per_patch, sim.ny_per_patch) +
                                          ((sim.nz_per_patch,) if sim.dimension == 3 else ())))
        return [np.array(p.fields.rho[inner]) for p in sim.patches]

    def __call__(self, sim):
        if not _interval_triggered(sim, self.interval):
            return
        target = self.species.ispec
        if target > 0 and sim.ispec == target - 1:
            self.prev_rho = self._interior_rho(sim)
            return
        if sim.ispec != target:
            return
        rho = self._interior_rho(sim)
        n = (sim.nx_per_patch, sim.ny_per_patch) + ((sim.nz_per_patch,) if sim.dimension == 3 else ())
        for ip, p in enumerate(sim.patches):
            d = rho[ip] if self.prev_rho is None else rho[ip] - self.prev_rho[ip]
            idx = (p.ipatch_x, p.ipatch_y) + ((p.ipatch_z,) if sim.dimension == 3 else ())
            self._full[tuple(np.s_[i * m:(i + 1) * m] for i, m in zip(idx, n))] = d / self.species.q
        if self.slice is not None:
            self.density[...] = self._full[self.slice]
        self.prev_rho = None
