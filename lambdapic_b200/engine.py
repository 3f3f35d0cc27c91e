"""DeviceEngine: owns one lpic_ctx (one GPU) plus the pinned host mirrors of its fields and particles.

The host mirrors have exactly the device layout (include/lpic_b200.h), so an upload/download of one attribute
of every patch is a single copy.  The numpy arrays handed to user code (``p.fields.ex``,
``p.particles[i].x`` ...) are views into these mirrors.  The device is authoritative between
:meth:`upload_all` and :meth:`download_all`.
"""
from __future__ import annotations

import ctypes as C
from contextlib import contextmanager

import numpy as np

from . import _lib
from ._lib import FIELD_ATTRS, PART_ATTRS, P_IS_DEAD, check

ALL_FIELDS = (1 << len(FIELD_ATTRS)) - 1
E_MASK, B_MASK, J_MASK = 0b111, 0b111000, 0b1111000000
RESIDENT_ATTRS = ["x", "y", "z", "w", "ux", "uy", "uz", "inv_gamma", "_id"]
PART_FIELD_ATTRS = PART_ATTRS[8:14]


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


@contextmanager
def numa_local(L, ctx):
    """Run the body on the CPUs of the GPU's NUMA node, so that pinned host memory allocated (= first touched) inside lands
    next to the GPU's PCIe root.  Best effort: any missing piece (sysfs, cpuset restrictions) leaves the affinity alone."""
    import os
    old = None
    try:
        buf = C.create_string_buffer(32)
        if L.lpic_device_pci_bus_id(ctx, buf, 32) == 0:
            bus = buf.value.decode().lower()
            node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
            if node >= 0:
                cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
                mine = os.sched_getaffinity(0)
                if cpus & mine and (cpus & mine) != mine:
                    old = mine
                    os.sched_setaffinity(0, cpus & mine)
    except (OSError, ValueError, AttributeError):
        old = None
    try:
        yield
    finally:
        if old is not None:
            try:
                os.sched_setaffinity(0, old)
            except OSError:
                pass


class HostBuffer:
    """numpy view over memory from lpic_host_alloc (pinned when a GPU is present)."""

    def __init__(self, nbytes: int):
        self.nbytes = max(int(nbytes), 8)
        self.ptr = _lib.lib().lpic_host_alloc(self.nbytes)
        if not self.ptr:
            raise MemoryError(f"lpic_host_alloc({self.nbytes}) failed")
        self._raw = (C.c_char * self.nbytes).from_address(self.ptr)

    def array(self, dtype, count, offset=0):
        return np.frombuffer(self._raw, dtype=dtype, count=count, offset=offset)

    def free(self):
        if self.ptr:
            self._raw = None
            _lib.lib().lpic_host_free(self.ptr)
            self.ptr = None


class SpeciesMirror:
    """Host arenas of one species, same slot layout as the device arenas."""

    def __init__(self, eng: "DeviceEngine", ispec: int, with_part: bool):
        self.eng, self.ispec, self.with_part = eng, ispec, with_part
        self.attrs = list(PART_ATTRS) if with_part else list(RESIDENT_ATTRS)
        self.buffers = {}
        self._host = None  # pinned arenas are allocated on first use (the bench's device-resident leg never needs them)
        self.off = self.pcap = self.npart = None
        self.total = 0
        self.refresh_layout()

    @property
    def host(self):
        if self._host is None:
            self._host = {}
            with numa_local(self.eng.L, self.eng.ctx):
                for a in self.attrs + ["is_dead"]:
                    need = self.total * (1 if a == "is_dead" else 8)
                    buf = self.buffers.get(a)
                    if buf is None or buf.nbytes < need:  # pinning is slow (~ms per 100 MB): keep buffers, grow with headroom
                        if buf is not None:
                            buf.free()
                        buf = HostBuffer(need + need // 4)
                        self.buffers[a] = buf
                    arr = buf.array(np.uint8 if a == "is_dead" else np.float64, self.total)
                    if a == "is_dead":
                        arr[:] = 1
                    self._host[a] = arr
        return self._host

    def refresh_layout(self):
        """Re-read the device layout; the host arenas are dropped if the arena moved (the device is authoritative)."""
        n = self.eng.npatch
        off, pcap, npart = (np.zeros(n, dtype=np.int64) for _ in range(3))
        total = C.c_int64(0)
        check(self.eng.L.lpic_species_layout(self.eng.ctx, self.ispec, _ptr(off), _ptr(pcap), _ptr(npart), C.byref(total)))
        same = self.off is not None and total.value == self.total and np.array_equal(off, self.off)
        self.off, self.pcap, self.npart, self.total = off, pcap, npart, int(total.value)
        if not same:
            self._host = None  # the views are rebuilt on next use; the pinned buffers stay while they are large enough

    def view(self, attr: str, p: int):
        a = self.host[attr][self.off[p]:self.off[p] + self.npart[p]]
        return a.view(np.bool_) if attr == "is_dead" else a

    def free(self):
        for b in self.buffers.values():
            b.free()
        self.buffers, self._host = {}, None


class DeviceEngine:
    def __init__(self, dim, npatch, nx, ny, nz, n_guard, dx, dy, dz, nspec, device=0):
        self.L = _lib.lib()
        self.dim, self.npatch, self.nspec = int(dim), int(npatch), int(nspec)
        self.nx, self.ny, self.nz, self.ng = int(nx), int(ny), int(nz if dim == 3 else 1), int(n_guard)
        self.dx, self.dy, self.dz = float(dx), float(dy), float(dz if dim == 3 else 1.0)
        self.nb = 26 if dim == 3 else 8
        self.ctx = self.L.lpic_create(dim, npatch, nx, ny, self.nz, n_guard, dx, dy, self.dz, nspec, device)
        if not self.ctx:
            raise _lib.LpicError(self.L.lpic_last_error().decode())
        self.shape = (nx + 2 * n_guard, ny + 2 * n_guard) + ((self.nz + 2 * n_guard,) if dim == 3 else ())
        self.ncell = int(self.L.lpic_field_cells(self.ctx))
        assert self.ncell == int(np.prod(self.shape))
        with numa_local(self.L, self.ctx):
            self._fbuf = HostBuffer(8 * len(FIELD_ATTRS) * npatch * self.ncell)
        self.fields_host = self._fbuf.array(np.float64, len(FIELD_ATTRS) * npatch * self.ncell).reshape(
            (len(FIELD_ATTRS), npatch) + self.shape)
        self.fields_host[...] = 0.0
        self.species = [None] * nspec
        self.sort_cfg = [None] * nspec
        self.npart_created = [None] * nspec
        self.rank = 0
        self.patch_index = np.arange(npatch, dtype=np.int64)
        self.x0 = self.y0 = self.z0 = None
        self.slot_order = False  # True: round-1 v1 particle kernel (memory order) instead of the cell-ordered one

    # ---- lifetime --------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "ctx", None):
            self.L.lpic_destroy(self.ctx)
            self.ctx = None
            for s in self.species:
                if s is not None:
                    s.free()
            self._fbuf.free()
            if getattr(self, "_psibuf", None) is not None:
                self._psibuf.free()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.L.lpic_sync(self.ctx))

    # ---- geometry --------------------------------------------------------------------------------------------
    def set_geometry(self, x0, y0, z0, neighbor_ipatch, box, glob, rank=0, patch_index=None):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
        self.x0, self.y0 = f(x0), f(y0)
        self.z0 = f(z0) if z0 is not None else np.zeros(self.npatch)
        self.nbr = np.ascontiguousarray(neighbor_ipatch, dtype=np.int64).reshape(self.npatch, self.nb)
        self.box = f(box).reshape(self.npatch, 6)
        self.glob = f(glob).reshape(6)
        self.rank = int(rank)
        if patch_index is not None:
            self.patch_index = np.ascontiguousarray(patch_index, dtype=np.int64)
        check(self.L.lpic_set_patch_geometry(self.ctx, _ptr(self.x0), _ptr(self.y0), _ptr(self.z0), _ptr(self.nbr),
                                             _ptr(self.box), _ptr(self.glob), self.rank, _ptr(self.patch_index)))

    # ---- fields ----------------------------------------------------------------------------------------------
    def field_view(self, attr: str, p: int):
        return self.fields_host[FIELD_ATTRS.index(attr), p]

    def upload_fields(self, mask=ALL_FIELDS):
        check(self.L.lpic_upload_fields(self.ctx, mask, _ptr(self.fields_host)))

    def download_fields(self, mask=ALL_FIELDS):
        check(self.L.lpic_download_fields(self.ctx, mask, _ptr(self.fields_host)))

    # ---- CPML ------------------------------------------------------------------------------------------------
    def configure_pml(self, instances):
        """instances: list of (patch, axis, slot, (e_lo, e_hi, b_lo, b_hi), profiles[6, nmax]).  Allocates the psi arena on
        the device and its pinned host mirror `psi_host` of shape (ninst, 4, nx, ny[, nz])."""
        n = len(instances)
        self.pml_instances = instances
        if getattr(self, "_psibuf", None) is not None:
            self._psibuf.free()
        self._psibuf, self.psi_host = None, None
        if n == 0:
            check(self.L.lpic_pml_configure(self.ctx, 0, None, None, None, None, None, 0))
            return
        nmax = max(self.nx, self.ny, self.nz)
        i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)  # noqa: E731
        patch, axis, slot = i64([i[0] for i in instances]), i64([i[1] for i in instances]), i64([i[2] for i in instances])
        ranges = i64([i[3] for i in instances]).reshape(n, 4)
        prof = np.ascontiguousarray(np.stack([i[4] for i in instances]), dtype=np.float64)
        assert prof.shape == (n, 6, nmax)
        check(self.L.lpic_pml_configure(self.ctx, n, _ptr(patch), _ptr(axis), _ptr(slot), _ptr(ranges), _ptr(prof), nmax))
        words = int(self.L.lpic_pml_psi_words(self.ctx))
        self._psibuf = HostBuffer(8 * words)
        interior = (self.nx, self.ny) + ((self.nz,) if self.dim == 3 else ())
        self.psi_host = self._psibuf.array(np.float64, words).reshape((n, 4) + interior)
        self.psi_host[...] = 0.0

    def upload_psi(self):
        if getattr(self, "psi_host", None) is not None:
            check(self.L.lpic_pml_upload_psi(self.ctx, _ptr(self.psi_host)))

    def download_psi(self):
        if getattr(self, "psi_host", None) is not None:
            check(self.L.lpic_pml_download_psi(self.ctx, _ptr(self.psi_host)))

    # ---- particles -------------------------------------------------------------------------------------------
    def alloc_species(self, ispec, npart, slack=1.5, min_extra=64, with_part=False, npart_created=None):
        npart = np.ascontiguousarray(npart, dtype=np.int64)
        check(self.L.lpic_species_alloc(self.ctx, ispec, _ptr(npart), float(slack), int(min_extra), int(with_part)))
        if self.species[ispec] is not None:
            self.species[ispec].free()
        self.species[ispec] = SpeciesMirror(self, ispec, bool(with_part))
        self.npart_created[ispec] = (np.array(npart_created, dtype=np.int64) if npart_created is not None else npart.copy())
        return self.species[ispec]

    def _record_table(self, m, names):
        """(mask, pointer table) of the attributes among `names` that live in the device's 64-byte records (ids 0..7)."""
        import ctypes as C
        tab = (C.c_void_p * 8)()
        mask = 0
        for a in names:
            aid = PART_ATTRS.index(a) if a != "is_dead" else -1
            if 0 <= aid < 8:
                tab[aid] = m.host[a].ctypes.data
                mask |= 1 << aid
        return mask, tab

    def upload_particles(self, ispec, attrs=None):
        m = self.species[ispec]
        names = list(attrs or m.attrs + ["is_dead"])
        mask, tab = self._record_table(m, names)
        if mask:  # x y z w ux uy uz inv_gamma in one chunked, double-buffered pass (lpic_upload_particle_records)
            check(self.L.lpic_upload_particle_records(self.ctx, ispec, mask, tab))
        for a in names:
            aid = P_IS_DEAD if a == "is_dead" else PART_ATTRS.index(a)
            if not (0 <= aid < 8):
                check(self.L.lpic_upload_particles(self.ctx, ispec, aid, _ptr(m.host[a])))

    def download_particles(self, ispec, attrs=None):
        m = self.species[ispec]
        names = list(attrs or m.attrs + ["is_dead"])
        mask, tab = self._record_table(m, names)
        if mask:
            check(self.L.lpic_download_particle_records(self.ctx, ispec, mask, tab))
        for a in names:
            aid = P_IS_DEAD if a == "is_dead" else PART_ATTRS.index(a)
            if not (0 <= aid < 8):
                check(self.L.lpic_download_particles(self.ctx, ispec, aid, _ptr(m.host[a])))

    def upload_all(self):
        self.upload_fields()
        self.upload_psi()
        for s in range(self.nspec):
            if self.species[s] is not None:
                self.upload_particles(s)
        self.sync()

    def download_all(self):
        self.download_fields()
        self.download_psi()
        for s in range(self.nspec):
            if self.species[s] is not None:
                self.download_particles(s)

    def extend(self, ispec, ext):
        """ParticlesBase.extend for every patch (core/particles.py:141-168); returns True if the arena moved."""
        ext = np.ascontiguousarray(ext, dtype=np.int64)
        if not ext.any():
            return False
        created = self.npart_created[ispec]
        ids = ((np.uint64(self.rank) << np.uint64(50)) | (self.patch_index.astype(np.uint64) << np.uint64(32))
               | created.astype(np.uint64))  # core/particles.py:91-116
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        moved = C.c_int(0)
        check(self.L.lpic_species_extend(self.ctx, ispec, _ptr(ext), _ptr(ids), C.byref(moved)))
        created += ext
        self.species[ispec].refresh_layout()
        return bool(moved.value)

    def set_npart(self, ispec, npart):
        """New slot counts inside the existing segments (each <= pcap); the host mirrors keep their arenas."""
        npart = np.ascontiguousarray(npart, dtype=np.int64)
        check(self.L.lpic_species_set_npart(self.ctx, ispec, _ptr(npart)))
        self.species[ispec].refresh_layout()

    # ---- operators (one call each; names follow the reference facades) ------------------------------------------
    def update_efield(self, dt):
        check(self.L.lpic_update_efield(self.ctx, float(dt)))

    def update_bfield(self, dt):
        check(self.L.lpic_update_bfield(self.ctx, float(dt)))

    def sync_guard_fields(self, mask):
        check(self.L.lpic_sync_guard_fields(self.ctx, int(mask)))

    def sync_currents(self):
        check(self.L.lpic_sync_currents(self.ctx))

    def reset_currents(self):
        check(self.L.lpic_reset_currents(self.ctx))

    def push_deposit(self, ispec, dt, q, m, write_part=False, slot_order=None):
        slot_order = self.slot_order if slot_order is None else slot_order
        flags = (_lib.PUSH_WRITE_PART if write_part else 0) | (_lib.PUSH_SLOT_ORDER if slot_order else 0)
        check(self.L.lpic_push_deposit(self.ctx, ispec, float(dt), float(q), float(m), flags))

    def interpolate(self, ispec):
        check(self.L.lpic_interpolate(self.ctx, ispec))

    def push_momentum(self, ispec, dt, q, m):
        check(self.L.lpic_push_momentum(self.ctx, ispec, float(dt), float(q), float(m)))

    def push_position(self, ispec, dt):
        check(self.L.lpic_push_position(self.ctx, ispec, float(dt)))

    def deposit(self, ispec, dt, q):
        check(self.L.lpic_deposit(self.ctx, ispec, float(dt), float(q)))

    def laser_bfields(self, laserpos, patches, ranges, ey_src, ez_src, dt):
        """Laser antenna at xmin for the listed edge patches; ey_src / ez_src: (n, NY[, NZ]) in the padded layout."""
        i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)  # noqa: E731
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
        patches, ranges, ey_src, ez_src = i64(patches), i64(ranges), f64(ey_src), f64(ez_src)
        assert ey_src.shape == (len(patches),) + self.shape[1:] == ez_src.shape
        check(self.L.lpic_laser_bfields(self.ctx, int(laserpos), len(patches), _ptr(patches), _ptr(ranges), _ptr(ey_src),
                                        _ptr(ez_src), float(dt)))

    def weighted_drift(self, ispec):
        out = np.zeros(2)
        check(self.L.lpic_weighted_drift(self.ctx, ispec, _ptr(out)))
        return float(out[0]), float(out[1])

    def configure_sort(self, ispec, nxb, nyb, nzb, dxb, dyb, dzb, x0s, y0s, z0s):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
        self.sort_cfg[ispec] = dict(nxb=int(nxb), nyb=int(nyb), nzb=int(nzb), dxb=float(dxb), dyb=float(dyb),
                                    dzb=float(dzb), x0s=f(x0s), y0s=f(y0s), z0s=f(z0s) if z0s is not None else np.zeros(self.npatch))

    def sort(self, ispec, reverse_x) -> int:
        c = self.sort_cfg[ispec]
        nbuf = C.c_int64(0)
        check(self.L.lpic_sort(self.ctx, ispec, int(bool(reverse_x)), c["nxb"], c["nyb"], c["nzb"], c["dxb"], c["dyb"], c["dzb"],
                               _ptr(c["x0s"]), _ptr(c["y0s"]), _ptr(c["z0s"]), C.byref(nbuf)))
        return int(nbuf.value)

    def sort_arrays(self, ispec):
        """bucket_count / bound_min / bound_max as (npatch, nbin) int64 and particle_index per patch."""
        c = self.sort_cfg[ispec]
        nbin = c["nxb"] * c["nyb"] * c["nzb"]
        out = {}
        for name, which in (("bucket_count", _lib.SORT_BUCKET_COUNT), ("bound_min", _lib.SORT_BOUND_MIN),
                            ("bound_max", _lib.SORT_BOUND_MAX)):
            a = np.zeros((self.npatch, nbin), dtype=np.int64)
            check(self.L.lpic_sort_download(self.ctx, ispec, which, _ptr(a)))
            out[name] = a
        m = self.species[ispec]
        arena = np.zeros(max(m.total, 1), dtype=np.int64)
        check(self.L.lpic_sort_download(self.ctx, ispec, _lib.SORT_PARTICLE_INDEX, _ptr(arena)))
        out["particle_index"] = [arena[m.off[p]:m.off[p] + m.npart[p]].copy() for p in range(self.npatch)]
        return out

    def migrate_count(self, ispec):
        n = self.npatch
        ext, inc, alive = (np.zeros(n, dtype=np.int64) for _ in range(3))
        out = np.zeros(n * self.nb, dtype=np.int64)
        check(self.L.lpic_migrate_count(self.ctx, ispec, _ptr(ext), _ptr(inc), _ptr(out), _ptr(alive)))
        return dict(to_extend=ext, incoming=inc, outgoing=out, alive=alive)

    def migrate_fill(self, ispec):
        check(self.L.lpic_migrate_fill(self.ctx, ispec))

    def sync_particles(self, ispec):
        """Patches.sync_particles for one species (core/patch/patch.py:705-764): count -> extend -> fill."""
        rec = self.migrate_count(ispec)
        rec["moved"] = self.extend(ispec, rec["to_extend"])
        self.migrate_fill(ispec)
        return rec

    def count_alive(self, ispec) -> int:
        out = C.c_int64(0)
        check(self.L.lpic_count_alive(self.ctx, ispec, C.byref(out)))
        return int(out.value)

    def kinetic_sum(self, ispec) -> float:
        out = C.c_double(0)
        check(self.L.lpic_kinetic_sum(self.ctx, ispec, C.byref(out)))
        return float(out.value)

    def field_energy_sums(self):
        out = np.zeros(2)
        check(self.L.lpic_field_energy_sums(self.ctx, _ptr(out)))
        return float(out[0]), float(out[1])

    def init_uniform(self, ispec, ppc, weight, uth, seed):
        check(self.L.lpic_species_init_uniform(self.ctx, ispec, int(ppc), float(weight), float(uth), int(seed)))

    def step_profiled(self, dt, q, m, reverse_x):
        """One step with a CUDA event after every operator; returns [(name, ms)] (bench.py --breakdown)."""
        marks, slot = [], [3000]

        def mark(name):
            self.record_event(slot[0])
            marks.append((name, slot[0]))
            slot[0] += 1
        mark("begin")
        self.update_efield(0.5 * dt); mark("update E field")
        self.sync_guard_fields(E_MASK); mark("sync E field")
        self.update_bfield(0.5 * dt); mark("update B field")
        self.sync_guard_fields(B_MASK); mark("sync B field")
        for s in range(self.nspec):
            self.sort(s, reverse_x[s]); mark(f"sort species {s}")
        self.reset_currents(); mark("reset J,rho")
        for s in range(self.nspec):
            self.push_deposit(s, dt, q[s], m[s]); mark(f"push+deposit species {s}")
        self.sync_currents(); mark("sync_currents")
        for s in range(self.nspec):
            self.sync_particles(s); mark(f"sync_particles species {s}")
        self.update_bfield(0.5 * dt); mark("update B field")
        self.sync_guard_fields(B_MASK); mark("sync B field")
        self.update_efield(0.5 * dt); mark("update E field")
        self.sync_guard_fields(E_MASK); mark("sync E field")
        out = []
        for (n0, s0), (n1, s1) in zip(marks[:-1], marks[1:]):
            ms = C.c_double(0)
            check(self.L.lpic_event_elapsed_ms(self.ctx, s0, s1, C.byref(ms)))
            out.append((n1, ms.value))
        return out

    # ---- one full step, periodic / unified-pusher case (simulation/simulation.py:937-1130) ---------------------
    def record_event(self, slot):
        check(self.L.lpic_event_record(self.ctx, int(slot)))

    def step(self, dt, q, m, reverse_x, write_part=False, event_slot=None, laser=None):
        self.update_efield(0.5 * dt); self.sync_guard_fields(E_MASK)
        self.update_bfield(0.5 * dt); self.sync_guard_fields(B_MASK)
        nbuf = [self.sort(s, reverse_x[s]) for s in range(self.nspec)]
        self.reset_currents()
        for s in range(self.nspec):
            if event_slot is not None:
                self.record_event(event_slot + 2 * s)
            self.push_deposit(s, dt, q[s], m[s], write_part)
            if event_slot is not None:
                self.record_event(event_slot + 2 * s + 1)
        self.sync_currents()
        mig = [self.sync_particles(s) for s in range(self.nspec)]
        self.update_bfield(0.5 * dt)
        if laser is not None:  # stage `_laser` (simulation.py:1098-1103): (laserpos, patches, ranges, ey_src, ez_src)
            self.laser_bfields(*laser, dt)
        self.sync_guard_fields(B_MASK)
        self.update_efield(0.5 * dt); self.sync_guard_fields(E_MASK)
        return nbuf, mig
