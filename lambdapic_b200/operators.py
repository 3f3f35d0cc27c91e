"""Operator facades owned by ``Simulation`` -- same class names, call signatures and public attributes as the
reference's L3 layer, each a thin call into the C-ABI through the DeviceBridge:

  MaxwellSolver2D/3D      core/maxwell/solver/solver.py:19-254
  ParticleSort2D/3D       core/sort/particle_sort.py:8-350
  BorisPusher             core/pusher/pusher.py:15-141
  CurrentDeposition2D/3D  core/current/deposition.py:7-208
  FieldInterpolation2D/3D core/interpolation/field_interpolation.py:9-180
"""
from __future__ import annotations

import numpy as np

from .engine import B_MASK, E_MASK  # noqa: F401


class _Op:
    def __init__(self, patches):
        self.patches = patches
        self.dimension = patches.dimension

    @property
    def bridge(self):
        return self.patches._need_bridge()

    @property
    def npatches(self):
        return self.patches.npatches

    # the reference rebuilds numba typed lists here; device arenas need no re-seating
    def generate_field_lists(self):
        pass

    def generate_particle_lists(self):
        pass

    def update_particle_lists(self, ipatch=None):
        pass

    def update_patches(self):
        pass


class MaxwellSolver(_Op):
    def update_efield(self, dt: float) -> None:
        with self.bridge.coherent():
            self.bridge.engine.update_efield(dt)

    def update_bfield(self, dt: float) -> None:
        with self.bridge.coherent():
            self.bridge.engine.update_bfield(dt)


class MaxwellSolver2D(MaxwellSolver):
    pass


class MaxwellSolver3D(MaxwellSolver):
    pass


class EnableOp(_Op):
    _enabled = True

    def enable(self):
        self._enabled = True

    def disable(self):
        self._enabled = False

    def is_enabled(self):
        return self._enabled


class ParticleSort(EnableOp):
    """Per-patch bucket sort.  Default construction = the reference's: buckets are x columns
    (nx_buckets = nx, ny_buckets = nz_buckets = 1, dy_buckets = Ly ...; simulation.py:691-698,1369-1376)."""

    def __init__(self, patches, ispec, nx_buckets=None, ny_buckets=None, nz_buckets=None,
                 dx_buckets=None, dy_buckets=None, dz_buckets=None, x0=None, y0=None, z0=None):
        super().__init__(patches)
        self.ispec = ispec
        self.nx_buckets = patches.nx if nx_buckets is None else nx_buckets
        self.ny_buckets = patches.ny if ny_buckets is None else ny_buckets
        self.dx_buckets = patches.dx if dx_buckets is None else dx_buckets
        self.dy_buckets = patches.dy if dy_buckets is None else dy_buckets
        if self.dimension == 3:
            self.nz_buckets = patches.nz if nz_buckets is None else nz_buckets
            self.dz_buckets = patches.dz if dz_buckets is None else dz_buckets
        else:
            self.nz_buckets, self.dz_buckets = 1, 1.0
        # bucket origins: half a cell below the patch origin (particle_sort.py:331-333)
        self.x0s = [p.x0 - patches.dx / 2 for p in patches]
        self.y0s = [p.y0 - patches.dy / 2 for p in patches]
        self.z0s = [getattr(p, "z0", 0.0) - (patches.dz / 2 if self.dimension == 3 else 0.0) for p in patches]
        self.reverse_x = None
        self.nbuf_last = 0
        self._configured_for = None

    def generate_field_lists(self):
        """Bucket origins follow the patch origins (particle_sort.py:178-203,316-342); called again after a
        MovingWindow shift."""
        ps = self.patches
        self.x0s = [p.x0 - ps.dx / 2 for p in ps]
        self.y0s = [p.y0 - ps.dy / 2 for p in ps]
        self.z0s = [getattr(p, "z0", 0.0) - (ps.dz / 2 if self.dimension == 3 else 0.0) for p in ps]
        self._configured_for = None

    def _configure(self):
        eng = self.bridge.engine
        if self._configured_for is not eng:
            eng.configure_sort(self.ispec, self.nx_buckets, self.ny_buckets, self.nz_buckets, self.dx_buckets,
                               self.dy_buckets, self.dz_buckets, self.x0s, self.y0s, self.z0s)
            self._configured_for = eng

    def _decide_reverse_x(self):
        """Mirrored bucket order when the species drifts towards -x, decided once (particle_sort.py:64-89)."""
        w, wux = self.bridge.engine.weighted_drift(self.ispec)
        comm = self.patches._comm
        if comm is not None and comm.Get_size() > 1:
            w, wux = comm.allreduce((w, wux), op=lambda a, b: (a[0] + b[0], a[1] + b[1]))
        return bool(w > 0.0 and wux / w < 0.0)

    def __call__(self) -> int:
        if not self._enabled:
            return 0
        with self.bridge.coherent():
            self._configure()
            if self.reverse_x is None:
                self.reverse_x = self._decide_reverse_x()
            self.nbuf_last = self.bridge.engine.sort(self.ispec, self.reverse_x)
        return self.nbuf_last

    def _arrays(self):
        self._configure()
        return self.bridge.engine.sort_arrays(self.ispec)

    def _shaped(self, a):
        shape = (self.nx_buckets, self.ny_buckets) + ((self.nz_buckets,) if self.dimension == 3 else ())
        return [a[ip].reshape(shape) for ip in range(self.npatches)]

    bucket_count_list = property(lambda s: s._shaped(s._arrays()["bucket_count"]))
    bucket_bound_min_list = property(lambda s: s._shaped(s._arrays()["bound_min"]))
    bucket_bound_max_list = property(lambda s: s._shaped(s._arrays()["bound_max"]))
    particle_index_list = property(lambda s: s._arrays()["particle_index"])


class ParticleSort2D(ParticleSort):
    pass


class ParticleSort3D(ParticleSort):
    pass


class PusherBase(EnableOp):
    def __init__(self, patches, ispec):
        super().__init__(patches)
        self.ispec = ispec
        self.q = patches.species[ispec].q
        self.m = patches.species[ispec].m

    def push_position(self, dt: float):
        """core/pusher/pusher.py:102-110: the reference only implements the 2D position push."""
        if not self._enabled:
            return
        if self.dimension == 2:
            with self.bridge.coherent():
                self.bridge.engine.push_position(self.ispec, dt)


class BorisPusher(PusherBase):
    def __call__(self, dt: float, unified: bool = False) -> None:
        if not self._enabled:
            return
        with self.bridge.coherent():
            if unified:
                self.bridge.engine.push_deposit(self.ispec, dt, self.q, self.m, write_part=self.bridge.with_part)
            else:
                self.bridge.engine.push_momentum(self.ispec, dt, self.q, self.m)


class CurrentDeposition(_Op):
    def __init__(self, patches):
        super().__init__(patches)
        self.q = [s.q for s in patches.species]

    def reset(self):
        with self.bridge.coherent():
            self.bridge.engine.reset_currents()

    def __call__(self, ispec: int, dt: float):
        if self.q[ispec] == 0:  # neutral species deposit nothing (core/current/deposition.py:162,194)
            return
        with self.bridge.coherent():
            self.bridge.engine.deposit(ispec, dt, self.q[ispec])


class CurrentDeposition2D(CurrentDeposition):
    pass


class CurrentDeposition3D(CurrentDeposition):
    pass


class FieldInterpolation(_Op):
    def __call__(self, ispec: int) -> None:
        with self.bridge.coherent():
            self.bridge.engine.interpolate(ispec)


class FieldInterpolation2D(FieldInterpolation):
    pass


class FieldInterpolation3D(FieldInterpolation):
    pass


class SingleRankMPI:
    """``sim.mpi`` for one rank: the reference's single-rank fast path (core/mpi/mpi_manager.py:111-146) --
    ``*_start`` returns None and ``*_wait(None)`` is a no-op."""

    def __init__(self, comm):
        self.comm = comm
        self.rank = comm.Get_rank()
        self.size = comm.Get_size()

    def sync_guard_fields_start(self, attrs):
        return None

    def sync_guard_fields_wait(self, handle):
        pass

    def sync_guard_fields(self, attrs):
        pass

    def sync_currents_start(self):
        return None

    def sync_currents_wait(self, handle):
        pass

    def sync_currents(self):
        pass

    def sync_particles_start(self, ispec):
        return None

    def sync_particles_wait(self, handle):
        pass

    def sync_particles(self, ispec=None):
        pass
