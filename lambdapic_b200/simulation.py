"""``Simulation`` / ``Simulation3D``: the reference's user-facing driver (simulation/simulation.py:118-1430) re-hosted
on the GPU engine.  Constructor arguments, public attributes, the 14 callback stages, the order of operators inside
one step and the timer names are the reference's; every operator is a facade over the C-ABI (operators.py).

Built on the device: periodic and CPML boundaries, the laser antenna, the moving window.  Outside the accelerated path
(raised explicitly instead of silently differing): QED, collisions, load balancing.  Multi-rank runs use
lambdapic_b200.multigpu (static block partition, NCCL halos inside the library).
"""
from __future__ import annotations

import time as _time
from dataclasses import dataclass, field
from typing import Callable, ClassVar, Dict, Optional, Sequence

import numpy as np

from .callback import _interval_triggered, callback as _callback_decorator
from .comm import default_comm
from .device import DeviceBridge
from .operators import (BorisPusher, CurrentDeposition2D, CurrentDeposition3D, FieldInterpolation2D,
                        FieldInterpolation3D, MaxwellSolver2D, MaxwellSolver3D, ParticleSort2D, ParticleSort3D,
                        SingleRankMPI)
from .patch import Patch2D, Patch3D, Patches
from .species import Species
from .workloads import C_LIGHT, make_patch_grid


class SimulationCallbacks:
    """simulation/simulation.py:1435-1509"""

    def __init__(self, callbacks, simulation):
        self.simulation = simulation
        self.stages = simulation.STAGES
        self.stage_callbacks = {stage: [] for stage in self.stages}
        for cb in callbacks or []:
            if hasattr(cb, "stage"):
                stage = cb.stage or simulation.DEFAULT_STAGE
                if stage not in self.stages:
                    raise ValueError(f"Invalid stage '{stage}'")
                self.stage_callbacks[stage].append(cb)
            else:
                wrapped = _callback_decorator(stage=simulation.DEFAULT_STAGE)(cb)
                self.stage_callbacks[wrapped.stage].append(wrapped)

    def run(self, stage: str):
        for cb in self.stage_callbacks[stage]:
            cb(self.simulation)

    def non_empty_stages(self):
        return [s for s, cbs in self.stage_callbacks.items() if cbs]

    def has_triggered_callbacks(self, stage: str) -> bool:
        return any(_interval_triggered(self.simulation, getattr(cb, "interval", 1)) for cb in self.stage_callbacks.get(stage, []))


class Timer:
    """Wall time per operator name (the reference's names; core/utils/timer.py:29-96).  With
    ``enable_timer`` every interval above 0.1 ms is also appended to ``<log_file stem>.timer.txt`` in the reference's record
    format (``... | TIMER | Rank r <name> took <ms>ms``), which is what ``lambdapic timer-stat`` aggregates
    (cli/stat.py:12).  Device work is asynchronous, so in timer mode the stream is synchronised when an interval closes:
    the numbers are per-operator device times, and the step is serialised while the timer is on."""
    totals: dict = {}
    enabled = False
    sink = None     # open text file of the TIMER records, or None
    rank = 0
    sync = None     # callable that waits for the device, set by Simulation.initialize

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if Timer.enabled:
            self.t0 = _time.perf_counter()
        return self

    def __exit__(self, *exc):
        if Timer.enabled:
            if Timer.sync is not None:
                Timer.sync()
            dt = _time.perf_counter() - self.t0
            Timer.totals[self.name] = Timer.totals.get(self.name, 0.0) + dt
            if Timer.sink is not None and dt > 1e-4:
                stamp = _time.strftime("%Y-%m-%d %H:%M:%S") + f".{int((_time.time() % 1) * 1000):03d}"
                Timer.sink.write(f"{stamp} | TIMER    | Rank {Timer.rank} {self.name} took {1e3 * dt:.1f}ms\n")
        return False

    @staticmethod
    def open_sink(log_file, truncate=True):
        """log.txt -> log.timer.txt (core/utils/logger.py:14-25)."""
        import os
        if Timer.sink is not None:
            Timer.sink.close()
            Timer.sink = None
        if not log_file:
            return None
        stem, ext = os.path.splitext(str(log_file))
        path = f"{stem}.timer{ext or '.txt'}"
        Timer.sink = open(path, "w" if truncate else "a", buffering=1)
        return path


@dataclass
class Simulation:
    nx: int
    ny: int
    nz: int = field(init=False)
    dx: float
    dy: float
    dz: float = field(init=False)
    npatch_x: int = field(default=0)
    npatch_y: int = field(default=0)
    npatch_z: int = field(init=False)
    nsteps: int | None = field(default=None)
    sim_time: float | None = field(default=None)
    dt_cfl: float = field(default=0.95)
    n_guard: int = field(default=3)
    boundary_conditions: Dict[str, str] = field(default_factory=lambda: {"xmin": "pml", "xmax": "pml", "ymin": "pml", "ymax": "pml"})
    cpml_thickness: int = field(default=6)
    log_file: Optional[str] = field(default=None)
    truncate_log: bool = field(default=True)
    enable_timer: bool = field(default=False)
    random_seed: Optional[int] = field(default=None)
    comm: object = field(default=None)
    device: int = field(default=0)               # CUDA device of this rank (extension over the reference)
    store_part_fields: bool = field(default=True)  # keep ex_part..bz_part resident (48 B/particle)

    STAGES: ClassVar[list[str]] = ["init", "start", "maxwell_1", "_push_position_1", "_interpolator", "_qed",
                                   "_push_momentum", "_push_position_2", "current_deposition", "qed_create_particles",
                                   "_laser", "maxwell_2", "end", "final"]
    DEFAULT_STAGE: ClassVar[str] = "end"
    dimension: ClassVar[int] = 2
    _auto_patch_cells: ClassVar[int] = 16  # auto patching picks ~16-cell patches (one CTA tile), not threads x 4

    # ---- configuration ---------------------------------------------------------------------------------------
    def _axes(self):
        return ("x", "y")

    def _validate(self):
        for a in self._axes():
            n, d, npt = getattr(self, f"n{a}"), getattr(self, f"d{a}"), getattr(self, f"npatch_{a}")
            if not (isinstance(n, (int, np.integer)) and n > 0):
                raise ValueError(f"n{a} must be a positive integer")
            if not d > 0:
                raise ValueError(f"d{a} must be positive")
            if npt == 0:  # auto patching (simulation.py:265-282 picks threads x 4; here: tiles of ~16 cells)
                npt = max(1, n // self._auto_patch_cells)
                while n % npt:
                    npt -= 1
                setattr(self, f"npatch_{a}", npt)
            if n % npt:
                raise ValueError(f"n{a}={n} must be divisible by npatch_{a}={npt}")
            if n // npt < self.n_guard:
                raise ValueError(f"patch size along {a} must be at least n_guard={self.n_guard}")
        if self.nsteps is not None and self.sim_time is not None:
            raise ValueError("nsteps and sim_time are mutually exclusive")
        if not 0 < self.dt_cfl <= 1.0:
            raise ValueError("dt_cfl must be in (0, 1]")
        faces = [f"{a}{s}" for a in self._axes() for s in ("min", "max")]
        bc = {k: self.boundary_conditions.get(k, "pml") for k in faces}
        for k, v in bc.items():
            if v not in ("pml", "periodic"):
                raise ValueError(f"boundary condition of {k} must be 'pml' or 'periodic'")
        for a in self._axes():
            if (bc[f"{a}min"] == "periodic") != (bc[f"{a}max"] == "periodic"):
                raise ValueError(f"{a}min and {a}max must both be periodic")
        self.boundary_conditions = bc
        if not isinstance(self.cpml_thickness, (int, np.integer)) or self.cpml_thickness < 1:
            raise ValueError("cpml_thickness must be a positive integer")

    def __post_init__(self):
        self.stages = list(self.STAGES)
        self._validate()
        inv = sum(getattr(self, f"d{a}") ** -2 for a in self._axes())
        self.dt = self.dt_cfl * inv ** -0.5 / C_LIGHT  # simulation.py:219,1288
        for a in self._axes():
            setattr(self, f"L{a}", getattr(self, f"n{a}") * getattr(self, f"d{a}"))
            setattr(self, f"n{a}_per_patch", getattr(self, f"n{a}") // getattr(self, f"npatch_{a}"))
        self.species: list[Species] = []
        self.itime, self.time = 0, 0.0
        self.rand_gen = None
        self.initialized = False
        self.collision = None
        self._current_sync_handle = None
        self.current_synced = False
        self.ispec = None
        self.istep = 0
        Timer.enabled = bool(self.enable_timer)
        if Timer.enabled:
            Timer.open_sink(self.log_file, self.truncate_log)

    # ---- species -----------------------------------------------------------------------------------------------
    def add_species(self, species: Sequence[Species]):
        if isinstance(species, Species):
            species = [species]
        names = [s.name for s in self.species]
        for s in species:
            if s.name in names:
                raise ValueError(f"Species name {s.name} already exists")
            if not s.is_compatible(self.dimension):
                raise ValueError(f"species {s.name}: density/ppc profile does not take {self.dimension} arguments")
            s.ispec = len(self.species)
            self.species.append(s)
            names.append(s.name)

    def add_collision(self, *a, **k):
        raise NotImplementedError("collisions are outside the accelerated path (SURVEY.md 2 #21)")

    # ---- initialisation (simulation.py:284-423) -----------------------------------------------------------------
    def _grid(self, rank, size):
        return make_patch_grid(2, self.npatch_x, self.npatch_y, 1, self.nx_per_patch, self.ny_per_patch, 1,
                               self.dx, self.dy, 0.0, self.n_guard, self._periodic(), rank, size)

    def _periodic(self):
        bc = self.boundary_conditions
        return tuple(bc.get(f"{a}min", "periodic") == "periodic" for a in "xyz")

    def _init_pml(self):
        """CPML faces for the patches on a non-periodic domain edge, in the reference's order xmin, xmax, ymin, ymax
        (, zmin, zmax) (simulation.py:450-464)."""
        from .pml import FACE_CLASS
        npatch = {"x": self.npatch_x, "y": self.npatch_y, "z": getattr(self, "npatch_z", 1)}
        for p in self.patches:
            for ax in self._axes():
                idx = getattr(p, f"ipatch_{ax}")
                for side, edge in (("min", 0), ("max", npatch[ax] - 1)):
                    if idx == edge and self.boundary_conditions[f"{ax}{side}"] == "pml":
                        p.add_pml_boundary(FACE_CLASS[f"{ax}{side}"](p.fields, thickness=self.cpml_thickness))
        if any(p.pml_boundary for p in self.patches):
            self.bridge.configure_pml()

    def create_patches(self, grid) -> Patches:
        patches = Patches(self.dimension)
        for k, g in enumerate(grid.index):
            ix, iy = int(g % self.npatch_x), int(g // self.npatch_x)
            # origins in the reference's arithmetic, i*L/npatch (simulation.py:482-483), not i*n*d: 1 ulp apart for some sizes
            p = Patch2D(grid.rank, int(g), ix, iy, ix * self.Lx / self.npatch_x, iy * self.Ly / self.npatch_y,
                        grid.nx, grid.ny, self.dx, self.dy)
            p.neighbor_index[:] = grid.neighbor_index[k]
            p.neighbor_ipatch[:] = grid.neighbor_ipatch[k]
            p.neighbor_rank[:] = grid.neighbor_rank[k]
            patches.append(p)
        return patches

    def _set_global_domain_bounds(self):
        self.patches.xmin_global, self.patches.xmax_global = -self.dx / 2, self.Lx - self.dx / 2
        self.patches.ymin_global, self.patches.ymax_global = -self.dy / 2, self.Ly - self.dy / 2

    def initialize(self):
        comm = self.comm if self.comm is not None else default_comm()
        rank, size = comm.Get_rank(), comm.Get_size()
        self.grid = self._grid(rank, size)
        self.patches = self.create_patches(self.grid)
        self.patches._comm = comm
        self._set_global_domain_bounds()
        self.bridge = DeviceBridge(self.patches, self.n_guard, device=self.device, with_part=self.store_part_fields,
                                   nspec=len(self.species))
        Timer.rank = rank
        Timer.sync = self.bridge.engine.sync if self.enable_timer else None
        if size > 1:
            from .multigpu import MultiRankMPI
            self.mpi = MultiRankMPI(self, comm)
        else:
            self.mpi = SingleRankMPI(comm)
        self._init_pml()
        for s in self.species:
            self.patches.add_species(s, aux_attrs=s._aux_attrs)
        # simulation.py:700-716: seed -> default_rng(seed).spawn(size)[rank]
        if self.random_seed is None:
            self.rand_gen = np.random.default_rng()
        else:
            self.rand_gen = np.random.default_rng(self.random_seed).spawn(size)[rank]
        self.patches.fill_particles(self.rand_gen)
        self.bridge.upload()  # allocates the device arenas and seats the particle views
        self.patches.sync_particles()
        three = self.dimension == 3
        self.maxwell = (MaxwellSolver3D if three else MaxwellSolver2D)(self.patches)
        self.interpolator = (FieldInterpolation3D if three else FieldInterpolation2D)(self.patches)
        self.current_depositor = (CurrentDeposition3D if three else CurrentDeposition2D)(self.patches)
        self.pusher = [BorisPusher(self.patches, i) for i in range(len(self.species))]
        self.radiation = [None] * len(self.species)
        self.pairproduction = [None] * len(self.species)
        self._init_sorter()
        self.load_balancer = None
        if self.mpi.size > 1:
            # The communicator belongs to set-up, as in the reference (MPIManager is built in initialize,
            # simulation/simulation.py:700-720): create it here and run one guard exchange of the (consistent) initial fields
            # so that NCCL's connection set-up -- seconds -- is not paid by the first step of run().
            xch = self.mpi.xch  # (straight on the device copy: the host mirrors stay authoritative until run() uploads them)
            xch.halo_start(self.mpi._mask(["ex", "ey", "ez"]), 0)
            xch.halo_wait()
            self.bridge.engine.sync()
        self.initialized = True
        comm.Barrier()

    def _init_sorter(self):
        """x-column buckets only (simulation.py:691-698)."""
        self.sorter = [ParticleSort2D(self.patches, i, nx_buckets=self.nx_per_patch, ny_buckets=1,
                                      dx_buckets=self.dx, dy_buckets=self.Ly) for i in range(len(self.species))]

    # ---- list maintenance hooks kept for API compatibility (device arenas need no re-seating) -----------------
    def generate_lists(self):
        pass

    def update_patches(self):
        pass

    def update_lists(self):
        """simulation.py:781-824 re-creates the lists of extended patches; here only the flags are cleared.  The full
        sweep over all patches runs when host code may have extended particles itself (mirrors authoritative)."""
        br = getattr(self, "bridge", None)
        if br is not None and br.resident:
            for pt in br.extended_particles:
                pt.extended = False
            br.extended_particles.clear()
            return
        for p in self.patches:
            for pt in p.particles:
                pt.extended = False
        if br is not None:
            br.extended_particles.clear()

    # ---- currents (simulation.py:1143-1188) ---------------------------------------------------------------------
    def sync_currents(self):
        if self.current_synced:
            return
        self.sync_currents_start()
        self.sync_currents_wait()

    def sync_currents_start(self):
        if self.current_synced or self._current_sync_handle is not None:
            return
        with Timer("sync_currents"):
            self.patches.sync_currents()
        with Timer("mpi.sync_currents (start)"):
            self._current_sync_handle = self.mpi.sync_currents_start()
        if self._current_sync_handle is None:
            self.current_synced = True

    def sync_currents_wait(self):
        if self._current_sync_handle is None:
            return
        with Timer("mpi.sync_currents (wait)"):
            self.mpi.sync_currents_wait(self._current_sync_handle)
        self._current_sync_handle = None
        self.current_synced = True

    def energies(self):
        """Field and kinetic energies [J] from device-side reductions (no host mirror traffic): dict with
        'electric', 'magnetic' and one entry per species name."""
        eps0, mu0 = 8.8541878188e-12, 1.25663706127e-06
        with self.bridge.coherent():
            eng = self.bridge.engine
            e2, b2 = eng.field_energy_sums()
            dV = self.dx * self.dy * (self.dz if self.dimension == 3 else 1.0)
            out = {"electric": 0.5 * eps0 * e2 * dV, "magnetic": 0.5 * b2 / mu0 * dV}
            for i, sp in enumerate(self.species):
                out[sp.name] = eng.kinetic_sum(i) * sp.m * C_LIGHT**2
        return out

    def maxwell_stage(self):
        """simulation.py:743-761: one full field advance without particles."""
        for upd, attrs in ((self.maxwell.update_efield, ["ex", "ey", "ez"]), (self.maxwell.update_bfield, ["bx", "by", "bz"])):
            upd(0.5 * self.dt)
            self._sync_guards(attrs)
        for upd, attrs in ((self.maxwell.update_bfield, ["bx", "by", "bz"]), (self.maxwell.update_efield, ["ex", "ey", "ez"])):
            upd(0.5 * self.dt)
            self._sync_guards(attrs)

    def _sync_guards(self, attrs, name="E"):
        with Timer(f"mpi sync {name} field"):
            h = self.mpi.sync_guard_fields_start(attrs)
        with Timer(f"sync {name} field"):
            self.patches.sync_guard_fields(attrs)
        with Timer(f"mpi sync {name} field (wait)"):
            self.mpi.sync_guard_fields_wait(h)

    def _handle_nsteps(self, nsteps, sim_time):
        if nsteps is not None and sim_time is not None:
            raise ValueError("Cannot specify both nsteps and sim_time in run() method")
        if nsteps is None and sim_time is None:
            if self.nsteps is not None:
                return self.nsteps
            if self.sim_time is not None:
                return int(self.sim_time / self.dt)
            raise ValueError("Must provide either nsteps or sim_time, either in Simulation or as an argument to run()")
        if sim_time is not None:
            return int(sim_time / self.dt)
        return nsteps + self.itime

    def get_cfl(self):
        inv = sum(getattr(self, f"d{a}") ** -2 for a in self._axes())
        return self.dt / (inv ** -0.5 / C_LIGHT)

    # ---- the hot loop (simulation.py:858-1141) -------------------------------------------------------------------
    def _stage(self, cbs: SimulationCallbacks, stage: str, timer_name: str):
        """Run the callbacks of a stage on coherent host mirrors: download before, upload after."""
        if not cbs.stage_callbacks[stage]:
            return
        if not cbs.has_triggered_callbacks(stage):
            return
        br = self.bridge
        # callbacks that only use device-side diagnostics (sim.energies()) declare `needs_host = False` and run
        # without the host mirrors being refreshed (mirror elision)
        triggered = [cb for cb in cbs.stage_callbacks[stage] if _interval_triggered(self, getattr(cb, "interval", 1))]
        hosted = [cb for cb in triggered if getattr(cb, "needs_host", True)]
        was_resident = br.resident and bool(hosted)
        partial = False
        if was_resident:
            # declared reads / writes: only those arrays cross PCIe (a callback without hints syncs everything)
            reads = None if any(getattr(cb, "reads", None) is None for cb in hosted) else {n for cb in hosted for n in cb.reads}
            writes = None if any(getattr(cb, "writes", None) is None for cb in hosted) else {n for cb in hosted for n in cb.writes}
            partial = reads is not None and writes is not None
            if partial:
                # the device stays authoritative for everything that was not named: operator facades and
                # sim.energies() called from these callbacks keep working on the device state
                br.download(reads | writes)  # what is written back is read first (callbacks usually modify in place)
            else:
                reads = writes = None
                br.download()
                br.resident = False
        with Timer(timer_name):
            cbs.run(stage)
        if was_resident:
            br.upload(writes)
            br.resident = True

    def run(self, nsteps: int | None = None, sim_time: float | None = None, callbacks: Optional[Sequence[Callable]] = None,
            stop_callback: Callable[..., bool] = lambda: False):
        cbs = SimulationCallbacks(callbacks or [], self)
        if not self.initialized:
            self.initialize()
        with Timer("Callbacks: init stage"):
            cbs.run("init")
        stages_in_pusher = {"_push_position_1", "_interpolator", "_qed", "_push_momentum", "_push_position_2"}
        unified_ok = not stages_in_pusher.intersection(cbs.non_empty_stages())
        use_unified_pusher = [isinstance(p, BorisPusher) and unified_ok for p in self.pusher]
        nsteps_total = self._handle_nsteps(nsteps, sim_time)
        self.mpi.comm.Barrier()
        br = self.bridge
        br.upload()          # users may have modified fields/particles between initialize() and run()
        br.resident = True
        E, B = ["ex", "ey", "ez"], ["bx", "by", "bz"]
        try:
            for self.istep in range(self.itime, nsteps_total):
                self._stage(cbs, "start", "Callbacks: start stage")
                with Timer("update E field"):
                    self.maxwell.update_efield(0.5 * self.dt)
                self._sync_guards(E, "E")
                with Timer("update B field"):
                    self.maxwell.update_bfield(0.5 * self.dt)
                self._sync_guards(B, "B")
                self._stage(cbs, "maxwell_1", "maxwell_1")

                for ispec, s in enumerate(self.patches.species):
                    if not s.is_enabled():
                        continue
                    self.ispec = ispec
                    with Timer(f"Sorting {self.species[ispec].name}"):
                        self.sorter[ispec]()

                self.current_depositor.reset()
                self.current_synced = False
                for ispec, s in enumerate(self.patches.species):
                    if not s.is_enabled():
                        continue
                    self.ispec = ispec
                    if use_unified_pusher[ispec]:
                        with Timer(f"unified pusher for {self.species[ispec].name}"):
                            self.pusher[ispec](self.dt, unified=True)
                    else:
                        with Timer("push_position"):
                            self.pusher[ispec].push_position(0.5 * self.dt)
                        self._stage(cbs, "_push_position_1", "Callbacks: _push_position_1 stage")
                        with Timer(f"Interpolation for {self.species[ispec].name}"):
                            self.interpolator(ispec)
                        self._stage(cbs, "_interpolator", "Callbacks: _interpolator stage")
                        self._stage(cbs, "_qed", "Callbacks: _qed stage")
                        with Timer(f"Pushing {self.species[ispec].name}"):
                            self.pusher[ispec](self.dt)
                        self._stage(cbs, "_push_momentum", "Callbacks: _push_momentum stage")
                        with Timer("push_position"):
                            self.pusher[ispec].push_position(0.5 * self.dt)
                        self._stage(cbs, "_push_position_2", "Callbacks: _push_position_2 stage")
                        with Timer(f"Current deposition for {self.species[ispec].name}"):
                            self.current_depositor(ispec, self.dt)
                    self.current_synced = False
                    self._stage(cbs, "current_deposition", "Callbacks: current_deposition stage")

                self.sync_currents_start()
                self.ispec = None
                with Timer("mpi.sync_particles"):
                    handles = [self.mpi.sync_particles_start(i) for i in range(len(self.patches.species))]
                    for h in handles:
                        self.mpi.sync_particles_wait(h)
                with Timer("sync_particles"):
                    self.patches.sync_particles()
                self.sync_currents_wait()
                with Timer("Updating lists"):
                    self.update_lists()
                self._stage(cbs, "qed_create_particles", "Callbacks: qed_create_particles stage")

                with Timer("update B field"):
                    self.maxwell.update_bfield(0.5 * self.dt)
                self._stage(cbs, "_laser", "laser")
                self._sync_guards(B, "B")
                with Timer("update E field"):
                    self.maxwell.update_efield(0.5 * self.dt)
                self._sync_guards(E, "E")
                self._stage(cbs, "maxwell_2", "Callbacks: maxwell_2 stage")
                self._stage(cbs, "end", "Callbacks: end stage")

                self.time += self.dt
                self.itime += 1
                if stop_callback():
                    return "stop by callback"
        finally:
            if br.resident:
                br.download()   # host mirrors are current again when run() returns
                br.resident = False
        self.mpi.comm.Barrier()
        with Timer("Callbacks: final stage"):
            cbs.run("final")


Simulation2D = Simulation


@dataclass
class Simulation3D(Simulation):
    nx: int
    ny: int
    nz: int
    dx: float
    dy: float
    dz: float
    npatch_x: int = field(default=0)
    npatch_y: int = field(default=0)
    npatch_z: int = field(default=0)
    boundary_conditions: Dict[str, str] = field(default_factory=lambda: {k: "pml" for k in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")})
    dimension: ClassVar[int] = 3

    def _axes(self):
        return ("x", "y", "z")

    def _grid(self, rank, size):
        return make_patch_grid(3, self.npatch_x, self.npatch_y, self.npatch_z, self.nx_per_patch, self.ny_per_patch,
                               self.nz_per_patch, self.dx, self.dy, self.dz, self.n_guard, self._periodic(), rank, size)

    def create_patches(self, grid) -> Patches:
        patches = Patches(3)
        for k, g in enumerate(grid.index):
            ix, iy, iz = int(g % self.npatch_x), int((g // self.npatch_x) % self.npatch_y), int(g // (self.npatch_x * self.npatch_y))
            p = Patch3D(grid.rank, int(g), ix, iy, iz, ix * self.Lx / self.npatch_x, iy * self.Ly / self.npatch_y,
                        iz * self.Lz / self.npatch_z, grid.nx, grid.ny, grid.nz, self.dx, self.dy, self.dz)  # simulation.py:1393
            p.neighbor_index[:] = grid.neighbor_index[k]
            p.neighbor_ipatch[:] = grid.neighbor_ipatch[k]
            p.neighbor_rank[:] = grid.neighbor_rank[k]
            patches.append(p)
        return patches

    def _set_global_domain_bounds(self):
        super()._set_global_domain_bounds()
        self.patches.zmin_global, self.patches.zmax_global = -self.dz / 2, self.Lz - self.dz / 2

    def _init_sorter(self):
        """simulation.py:1369-1376"""
        self.sorter = [ParticleSort3D(self.patches, i, nx_buckets=self.nx_per_patch, ny_buckets=1, nz_buckets=1,
                                      dx_buckets=self.dx, dy_buckets=self.Ly, dz_buckets=self.Lz)
                       for i in range(len(self.species))]
