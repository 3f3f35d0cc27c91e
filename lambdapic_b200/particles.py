"""Per-(patch, species) particle container with the reference's attribute names (core/particles.py:8-217).

Arrays are views into the species' pinned host arena (same slot layout as the device).  ``extend`` / ``prune``
called by user code on the host detach the arrays (plain numpy, ``extended = True``); the next upload
re-allocates the device arena to the new capacities."""
from __future__ import annotations

import numpy as np

from ._lib import PART_ATTRS


class ParticlesBase:
    def __init__(self, ipatch: int = 0, rank: int = 0):
        self.attrs = list(PART_ATTRS)
        self.extended = False
        self._npart_created = 0
        self.rank, self.ipatch = int(rank or 0), int(ipatch or 0)
        assert 0 <= self.rank < 2**14 and 0 <= self.ipatch < 2**18
        self.npart = 0
        self._detached = True

    def _generate_ids(self, start: int, count: int):
        assert start + count <= 2**32
        local = np.arange(start, start + count, dtype=np.uint64)
        bits = (np.uint64(self.rank) << np.uint64(50)) | (np.uint64(self.ipatch) << np.uint64(32)) | local
        return bits.view(np.float64)

    def initialize(self, npart: int) -> None:
        assert npart >= 0
        self.npart = int(npart)
        for a in self.attrs:
            setattr(self, a, np.zeros(self.npart))
        self.inv_gamma[:] = 1
        self.is_dead = np.full(self.npart, False)
        self._id[:] = self._generate_ids(self._npart_created, self.npart)
        self._npart_created += self.npart
        self._detached = True

    def extend(self, n: int):
        if n <= 0:
            return
        for a in self.attrs:
            new = np.full(self.npart + n, np.nan)
            new[:self.npart] = getattr(self, a)
            setattr(self, a, new)
        self.w[-n:] = 0
        self._id[-n:] = self._generate_ids(self._npart_created, n)
        self._npart_created += n
        self.is_dead = np.concatenate([np.asarray(self.is_dead, dtype=bool), np.ones(n, dtype=bool)])
        self.npart += n
        self.extended = True
        self._detached = True

    def prune(self, extra_buff: float = 0.1):
        n_alive = int(self.is_alive.sum())
        npart = int(n_alive * (1 + extra_buff))
        if npart >= self.npart:
            return None
        idx = np.argsort(self.is_dead, kind="stable")
        for a in self.attrs:
            setattr(self, a, np.ascontiguousarray(getattr(self, a)[idx][:npart]))
        self.is_dead = np.ascontiguousarray(np.asarray(self.is_dead)[idx][:npart])
        self.npart = npart
        self.extended = True
        self._detached = True
        return idx

    @property
    def id(self):
        return self._id.view(np.uint64)

    @property
    def is_alive(self):
        return np.logical_not(self.is_dead)
