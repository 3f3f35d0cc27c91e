"""DeviceBridge: keeps the host objects of a Simulation (Patches / Fields / ParticlesBase) and the GPU state coherent.

Protocol (SURVEY.md 8b "stage <-> mirror contract"):
  * the device is authoritative while ``resident`` is True (inside ``Simulation.run``'s loop);
  * before callbacks of a stage run, everything is downloaded into the host mirrors; after them everything is
    uploaded again (callbacks may write anything);
  * operator facades called while not resident (user code outside ``run``) upload, run, and download, so they
    behave like the reference's host-array operators.
"""
from __future__ import annotations

from contextlib import contextmanager

import numpy as np

from . import _lib
from ._lib import FIELD_ATTRS, P_IS_DEAD, PART_ATTRS
from .engine import ALL_FIELDS, DeviceEngine
from .fields import Fields2D, Fields3D


class DeviceBridge:
    def __init__(self, patches, n_guard, device=0, with_part=True, slack=1.3, nspec=0):
        self.patches = patches
        dim = patches.dimension
        p0 = patches[0]
        self.engine = DeviceEngine(dim, patches.npatches, p0.nx, p0.ny, getattr(p0, "nz", 1), n_guard,
                                   p0.dx, p0.dy, getattr(p0, "dz", 0.0), nspec, device)
        self.dim = dim
        self.with_part, self.slack, self.device = with_part, slack, device
        self._resident = False
        self.host_particles_valid = True  # False while the device may hold a newer particle layout than the host objects
        self.extended_particles = []  # particle objects flagged `extended` since the last Simulation.update_lists()
        self.stats = dict(uploads=0, downloads=0, h2d_bytes=0, d2h_bytes=0)
        self._set_geometry()
        F = Fields3D if dim == 3 else Fields2D
        for ip, p in enumerate(patches):
            p.set_fields(F(self.engine, ip, dim, p.x0, p.y0, getattr(p, "z0", 0.0)))
        patches._bridge = self

    def _set_geometry(self):
        ps, dim = self.patches, self.dim
        x0 = np.array([p.x0 for p in ps]); y0 = np.array([p.y0 for p in ps])
        z0 = np.array([getattr(p, "z0", 0.0) for p in ps])
        nbr = np.stack([p.neighbor_ipatch for p in ps]).astype(np.int64)
        box = np.zeros((ps.npatches, 6))
        for i, p in enumerate(ps):  # widened by half a cell, core/patch/sync_particles_3d.c:402-411
            box[i, 0:4] = [p.xmin - p.dx / 2, p.xmax + p.dx / 2, p.ymin - p.dy / 2, p.ymax + p.dy / 2]
            if dim == 3:
                box[i, 4:6] = [p.zmin - p.dz / 2, p.zmax + p.dz / 2]
        glob = np.array([ps.xmin_global, ps.xmax_global, ps.ymin_global, ps.ymax_global,
                         ps.zmin_global or 0.0, ps.zmax_global or 0.0], dtype=float)
        rank = ps[0].rank or 0
        self.engine.set_geometry(x0, y0, z0, nbr, box, glob, rank, np.array([p.index for p in ps], dtype=np.int64))

    @property
    def resident(self):
        """True while the device is authoritative (inside Simulation.run's loop)."""
        return self._resident

    @resident.setter
    def resident(self, value):
        self._resident = bool(value)
        if value:
            self.host_particles_valid = False  # from here on the device may re-lay-out its particle arenas

    def refresh_geometry(self, pml_changed=False):
        """Re-register origins, neighbour tables and particle boxes after the host objects changed them
        (MovingWindow); `pml_changed`: the set of CPML faces changed too."""
        if pml_changed:
            self.configure_pml()
        else:
            self._set_geometry()

    # ---- CPML --------------------------------------------------------------------------------------------------
    def configure_pml(self):
        """Register every patch's PML faces with the device (after Simulation._init_pml) and seat their psi arrays."""
        from .pml import PSI_NAMES
        eng, ps = self.engine, self.patches
        nmax = max(eng.nx, eng.ny, eng.nz)
        inst, owners = [], []
        for ip, p in enumerate(ps):
            for slot, m in enumerate(p.pml_boundary):
                inst.append((ip, m.axis, slot, (m.efield_start, m.efield_end, m.bfield_start, m.bfield_end), m.profiles(nmax)))
                owners.append(m)
        old = [[np.array(getattr(m, nm)) for nm in PSI_NAMES[m.axis]] for m in owners]
        eng.configure_pml(inst)
        for e, m in enumerate(owners):
            for r, nm in enumerate(PSI_NAMES[m.axis]):
                eng.psi_host[e, r] = old[e][r]
                setattr(m, nm, eng.psi_host[e, r])
        self._set_geometry()  # particle boxes shrink by the PML thickness

    # ---- species -----------------------------------------------------------------------------------------------
    def _recreate_engine_species(self):
        """(Re)allocate the device arenas from the host particle objects (all species)."""
        ps, eng = self.patches, self.engine
        nspec = len(ps.species)
        if eng.nspec != nspec:  # species are added before initialize() finishes: rebuild the context once
            fields = eng.fields_host.copy()
            p0 = ps[0]
            eng.close()
            self.engine = eng = DeviceEngine(self.dim, ps.npatches, p0.nx, p0.ny, getattr(p0, "nz", 1), p0.fields.n_guard,
                                             p0.dx, p0.dy, getattr(p0, "dz", 0.0), nspec, self.device)
            self._set_geometry()
            eng.fields_host[...] = fields
            for ip, p in enumerate(ps):
                for a in FIELD_ATTRS:
                    setattr(p.fields, a, eng.field_view(a, ip))
            if any(p.pml_boundary for p in ps):
                self.configure_pml()
        for s in range(nspec):
            self._alloc_species_from_host(s)

    def _alloc_species_from_host(self, s):
        ps, eng = self.patches, self.engine
        parts = [p.particles[s] for p in ps]
        npart = np.array([pt.npart for pt in parts], dtype=np.int64)
        created = np.array([pt._npart_created for pt in parts], dtype=np.int64)
        old_mirror, eng.species[s] = eng.species[s], None  # host arrays may be views into it: free it after the copy
        m = eng.alloc_species(s, npart, slack=self.slack, min_extra=64, with_part=self.with_part, npart_created=created)
        for ip, pt in enumerate(parts):
            old = {a: getattr(pt, a) for a in PART_ATTRS}   # references, no copies: one patch at a time
            old_dead = pt.is_dead
            self._seat(pt, m, ip)
            for a in m.attrs:
                getattr(pt, a)[...] = old[a]
            pt.is_dead[...] = old_dead
            if not self.with_part:
                self._host_only_part_fields(pt)
        if old_mirror is not None:
            old_mirror.free()

    def _reseat_in_place(self, s) -> bool:
        """Host code replaced some patches' particle arrays (initialize / extend / prune).  If every patch still fits its
        device segment, keep all arenas (device and pinned host) and copy only the detached patches into their mirror
        slices; returns False when a full re-allocation is needed."""
        eng = self.engine
        m = eng.species[s]
        if m is None or m._host is None:
            return False
        parts = [p.particles[s] for p in self.patches]
        npart = np.array([pt.npart for pt in parts], dtype=np.int64)
        if (npart > m.pcap).any():
            return False
        changed = [ip for ip, pt in enumerate(parts) if pt._detached or pt.npart != int(m.npart[ip])]
        old = {ip: ({a: getattr(parts[ip], a) for a in m.attrs}, parts[ip].is_dead) for ip in changed}
        eng.set_npart(s, npart)  # keeps the arenas: off / total are unchanged
        assert m._host is not None
        for ip in changed:
            pt = parts[ip]
            vals, dead = old[ip]
            self._seat(pt, m, ip)
            for a in m.attrs:
                getattr(pt, a)[...] = vals[a]
            pt.is_dead[...] = dead
            if not self.with_part:
                self._host_only_part_fields(pt)
            eng.npart_created[s][ip] = pt._npart_created
        return True

    def recycle(self, recycled):
        """MovingWindow fast path (device authoritative): the listed patches were recycled on the host -- fresh particle
        arrays from the loader, vacuum fields.  Their fields and psi arrays are cleared ON the device and only their new
        particles are uploaded; every other patch keeps its device state and no mirror is refreshed.  Geometry (origins,
        neighbour tables, particle boxes) is re-registered.  Returns the bytes sent."""
        import ctypes as C
        eng, ps = self.engine, self.patches
        assert self.resident, "recycle() is the device-resident path"
        recycled = [int(ip) for ip in recycled]
        nbytes = 0
        for s in range(eng.nspec):
            m = eng.species[s]
            parts = [ps[ip].particles[s] for ip in recycled]
            if not any(pt._detached for pt in parts):
                continue  # species without a density profile: the loader left it alone
            npart = m.npart.copy()
            ext = np.zeros(eng.npatch, dtype=np.int64)
            for ip, pt in zip(recycled, parts):
                if pt.npart > m.pcap[ip]:
                    ext[ip] = pt.npart - m.npart[ip]
            if ext.any():  # a recycled patch holds more particles than its segment: grow it on the device first
                eng.extend(s, ext)
                m = eng.species[s]
                npart = m.npart.copy()
            for ip, pt in zip(recycled, parts):
                npart[ip] = pt.npart
            eng.set_npart(s, npart)
            m = eng.species[s]
            m.host  # the pinned arenas exist from here on
            for ip, pt in zip(recycled, parts):
                vals, dead = {a: getattr(pt, a) for a in m.attrs}, pt.is_dead
                self._seat(pt, m, ip)
                for a in m.attrs:
                    getattr(pt, a)[...] = vals[a]
                pt.is_dead[...] = dead
                if not self.with_part:
                    self._host_only_part_fields(pt)
                eng.npart_created[s][ip] = pt._npart_created
                for a in m.attrs + ["is_dead"]:
                    aid = P_IS_DEAD if a == "is_dead" else PART_ATTRS.index(a)
                    _lib.check(eng.L.lpic_upload_particles_patch(eng.ctx, s, aid, ip, C.c_void_p(m.host[a].ctypes.data)))
                nbytes += pt.npart * (8 * len(m.attrs) + 1)
        arr = np.ascontiguousarray(recycled, dtype=np.int64)
        _lib.check(eng.L.lpic_zero_patches(eng.ctx, len(recycled), C.c_void_p(arr.ctypes.data)))
        self._set_geometry()
        eng.sync()  # the pinned slices may be rewritten by the next shift
        self.stats["h2d_bytes"] += int(nbytes)
        self.stats["recycles"] = self.stats.get("recycles", 0) + 1
        return nbytes

    @staticmethod
    def _host_only_part_fields(pt):
        """store_part_fields=False: ex_part..bz_part are not resident on the device; the host objects expose
        read-only zero arrays of the right length (no memory behind them)."""
        z = np.broadcast_to(0.0, (pt.npart,))  # one zero-stride view serves the six names (np.broadcast_to is ~4 us a call)
        for a in PART_ATTRS[8:14]:
            setattr(pt, a, z)

    def _seat(self, pt, m, ip):
        for a in m.attrs:
            setattr(pt, a, m.view(a, ip))
        pt.is_dead = m.view("is_dead", ip)
        pt.npart = int(m.npart[ip])
        pt._detached = False

    def _layout_changed_on_host(self, s):
        m = self.engine.species[s] if s < len(self.engine.species) else None
        if m is None:
            return True
        for ip, p in enumerate(self.patches):
            pt = p.particles[s]
            if pt._detached or pt.npart != int(m.npart[ip]):
                return True
        return False

    # ---- coherency ---------------------------------------------------------------------------------------------
    @staticmethod
    def _selection(names):
        """names (callback reads/writes hints) -> (field mask, psi?, particles?); None selects everything."""
        if names is None:
            return ALL_FIELDS, True, True
        mask, psi, particles = 0, False, False
        for n in names:
            if n == "fields":
                mask = ALL_FIELDS
            elif n == "psi":
                psi = True
            elif n == "particles":
                particles = True
            elif n in FIELD_ATTRS:
                mask |= 1 << FIELD_ATTRS.index(n)
            else:
                raise ValueError(f"unknown mirror name {n!r} (field attribute, 'fields', 'psi' or 'particles')")
        return mask, psi, particles

    def upload(self, names=None):
        import time
        t0 = time.perf_counter()
        try:
            return self._upload(names)
        finally:
            self.stats["h2d_seconds"] = self.stats.get("h2d_seconds", 0.0) + time.perf_counter() - t0

    def _upload(self, names=None):
        ps, eng = self.patches, self.engine
        mask, psi, particles = self._selection(names)
        if not (mask or psi or particles):
            return  # read-only callbacks: nothing to send back
        nspec = len(ps.species)
        if eng.nspec != nspec:
            self._recreate_engine_species()
            mask, psi, particles = ALL_FIELDS, True, True
        elif particles:
            if not self.host_particles_valid:
                raise RuntimeError("the host particle mirrors are stale (the device moved on since the last download); "
                                   "a callback that writes 'particles' must also read them")
            for s in range(nspec):
                if self._layout_changed_on_host(s) and not self._reseat_in_place(s):
                    self._alloc_species_from_host(s)
        if names is None or (mask == ALL_FIELDS and psi and particles):
            eng.upload_all()
            nbytes = self.state_bytes()
        else:
            nbytes = 0
            if mask:
                eng.upload_fields(mask)
                nbytes += bin(mask).count("1") * eng.fields_host[0].nbytes
            if psi and getattr(eng, "psi_host", None) is not None:
                eng.upload_psi()
                nbytes += eng.psi_host.nbytes
            if particles:
                for s in range(eng.nspec):
                    eng.upload_particles(s)
                    m = eng.species[s]
                    nbytes += m.total * (8 * len(m.attrs) + 1)
            eng.sync()
        self.stats["uploads"] += 1
        self.stats["h2d_bytes"] += int(nbytes)

    def download(self, names=None):
        import time
        t0 = time.perf_counter()
        try:
            return self._download(names)
        finally:
            self.stats["d2h_seconds"] = self.stats.get("d2h_seconds", 0.0) + time.perf_counter() - t0

    def _download(self, names=None):
        eng = self.engine
        mask, psi, particles = self._selection(names)
        if not (mask or psi or particles):
            return
        if names is not None and not (mask == ALL_FIELDS and psi and particles):
            nbytes = 0
            if particles:
                self._download_particles()
                nbytes += sum(m.total * (8 * len(m.attrs) + 1) for m in eng.species)
            if mask:
                eng.download_fields(mask)
                nbytes += bin(mask).count("1") * eng.fields_host[0].nbytes
            if psi and getattr(eng, "psi_host", None) is not None:
                eng.download_psi()
                nbytes += eng.psi_host.nbytes
            eng.sync()
            self.stats["downloads"] += 1
            self.stats["d2h_bytes"] += int(nbytes)
            return
        self._download_particles()
        eng.download_fields(ALL_FIELDS)
        eng.download_psi()
        self.stats["downloads"] += 1
        self.stats["d2h_bytes"] += self.state_bytes()

    def _download_particles(self):
        eng = self.engine
        self.host_particles_valid = True
        for s in range(eng.nspec):
            m = eng.species[s]
            reseat = m._host is None
            eng.download_particles(s)
            for ip, p in enumerate(self.patches):
                pt = p.particles[s]
                if reseat or pt.npart != int(m.npart[ip]) or pt.x.size != int(m.npart[ip]):
                    if pt.npart != int(m.npart[ip]) or pt.x.size != int(m.npart[ip]):
                        pt.extended = True
                    self._seat(pt, m, ip)
                    if not self.with_part:
                        self._host_only_part_fields(pt)
                pt._npart_created = int(eng.npart_created[s][ip])

    def state_bytes(self):
        eng = self.engine
        n = eng.fields_host.nbytes
        for s in range(eng.nspec):
            m = eng.species[s]
            n += m.total * (8 * len(m.attrs) + 1)
        if getattr(eng, "psi_host", None) is not None:
            n += eng.psi_host.nbytes
        return int(n)

    @contextmanager
    def coherent(self):
        """Run a device operator on host-authoritative data (facade called outside the resident loop)."""
        if self.resident:
            yield
            return
        self.upload()
        yield
        self.download()

    # ---- operators used by the facades -------------------------------------------------------------------------
    def sync_guard_fields(self, attrs):
        mask = 0
        for a in attrs:
            mask |= 1 << FIELD_ATTRS.index(a)
        with self.coherent():
            self.engine.sync_guard_fields(mask)

    def sync_currents(self):
        with self.coherent():
            self.engine.sync_currents()

    def sync_particles(self):
        total = np.zeros(self.patches.npatches, dtype=np.int64)
        with self.coherent():
            xch = getattr(self.engine, "comm_exchange", None)
            for s in range(self.engine.nspec):
                if xch is not None and s in xch.migrated:  # sim.mpi.sync_particles_start already did the whole migration
                    rec = {"to_extend": xch.migrated.pop(s)}
                else:
                    rec = self.engine.sync_particles(s)
                total += rec["to_extend"]
                for ip in np.nonzero(rec["to_extend"] > 0)[0]:  # no per-patch Python loop on the (usual) quiet steps
                    pt = self.patches[int(ip)].particles[s]
                    pt.extended = True
                    self.extended_particles.append(pt)
        return total

    def laser_bfields(self, laserpos, patches, ranges, ey_src, ez_src, dt):
        with self.coherent():
            self.engine.laser_bfields(laserpos, patches, ranges, ey_src, ez_src, dt)

    def close(self):
        self.engine.close()
