"""CPU oracle driver for the lambdaPIC per-step inner loop.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package (lambdapic_b200/) never imports this module.

Two interchangeable back-ends execute one full PIC step on a host-side :class:`OState`:

* ``"port"`` -- our own C restatement ``oracle/pic_oracle.c`` (ctypes, per-patch calls);
* ``"ref"``  -- the reference's own C extension modules compiled into ``oracle/_ref/`` by
  ``oracle/Makefile`` (unified pusher, sort, guard/current sync, particle migration), driven with
  duck-typed ``fields`` / ``particles`` / ``patch`` objects exactly as the reference's Python facades
  drive them (core/pusher/pusher.py:116-141, core/sort/particle_sort.py:335-350,
  core/patch/patch.py:670-764).  The numba FDTD of the reference (core/maxwell/cpu.py) cannot travel,
  so both back-ends use the C restatement of it (bit-exact against the golden vectors).

Step order restates simulation/simulation.py:937-1130 for the periodic, unified-pusher case.
"""
from __future__ import annotations

import ctypes
import importlib.machinery
import importlib.util
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
C_LIGHT = 299792458.0
EPSILON_0 = 8.8541878188e-12  # scipy.constants.epsilon_0 (CODATA 2022), used by core/maxwell/cpu.py:91

FIELD_ATTRS = ["ex", "ey", "ez", "bx", "by", "bz", "jx", "jy", "jz", "rho"]
PART_ATTRS = ["x", "y", "z", "w", "ux", "uy", "uz", "inv_gamma",
              "ex_part", "ey_part", "ez_part", "bx_part", "by_part", "bz_part", "_id"]  # core/particles.py:63-67

_c_i64 = ctypes.c_int64
_c_dbl = ctypes.c_double
_vp = ctypes.c_void_p


def build(force: bool = False) -> None:
    """Compile the C restatement (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(HERE, "_build", "libpic_oracle.so")
    src = os.path.join(HERE, "pic_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        build()
        _LIB = ctypes.CDLL(os.path.join(HERE, "_build", "libpic_oracle.so"))
        _LIB.orc_sort_patch.restype = _c_i64
    return _LIB


def have_ref() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "pusher", "unified", "unified_pusher_3d.so"))


_REF_CACHE = {}


def ref_module(rel: str):
    """Load one of the reference's extension modules from oracle/_ref/<rel>.so."""
    if rel not in _REF_CACHE:
        path = os.path.join(HERE, "_ref", rel + ".so")
        name = os.path.basename(rel)
        loader = importlib.machinery.ExtensionFileLoader(name, path)
        spec = importlib.util.spec_from_file_location(name, path, loader=loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        _REF_CACHE[rel] = mod
    return _REF_CACHE[rel]


def _p(a: np.ndarray):
    return a.ctypes.data_as(_vp)


def _pp(arrs):
    return (_vp * len(arrs))(*[a.ctypes.data for a in arrs])


# --------------------------------------------------------------------------------------------------
# duck-typed host state (attribute names are the reference's: core/fields.py, core/particles.py, patch.py)
# --------------------------------------------------------------------------------------------------
class OFields:
    def __init__(self, dim, nx, ny, nz, ng, dx, dy, dz, x0, y0, z0):
        self.nx, self.ny, self.nz, self.n_guard = nx, ny, nz, ng
        self.dx, self.dy, self.dz = dx, dy, dz
        self.x0, self.y0, self.z0 = x0, y0, z0
        self.shape = (nx + 2 * ng, ny + 2 * ng) + ((nz + 2 * ng,) if dim == 3 else ())
        for a in FIELD_ATTRS:
            setattr(self, a, np.zeros(self.shape))


class OParticles:
    def __init__(self, ipatch, rank=0):
        self.attrs = list(PART_ATTRS)
        self.ipatch, self.rank = ipatch, rank
        self.npart = 0
        self._npart_created = 0
        self.extended = False

    def _ids(self, start, count):  # core/particles.py:91-116
        local = np.arange(start, start + count, dtype=np.uint64)
        bits = (np.uint64(self.rank) << np.uint64(50)) | (np.uint64(self.ipatch) << np.uint64(32)) | local
        return bits.view(np.float64)

    def extend(self, n):  # core/particles.py:141-168
        if n <= 0:
            return
        for a in self.attrs:
            old = getattr(self, a)
            new = np.empty(self.npart + n)
            new[:self.npart] = old
            new[self.npart:] = np.nan
            setattr(self, a, new)
        self.w[-n:] = 0
        self._id[-n:] = self._ids(self._npart_created, n)
        self._npart_created += n
        self.is_dead = np.concatenate([self.is_dead, np.ones(n, dtype=bool)])
        self.npart += n
        self.extended = True


class OPatch:
    def __init__(self, index, x0, y0, z0, nx, ny, nz, dx, dy, dz, neighbor_ipatch):
        self.index = index
        self.x0, self.y0, self.z0 = x0, y0, z0
        self.nx, self.ny, self.nz, self.dx, self.dy, self.dz = nx, ny, nz, dx, dy, dz
        self.neighbor_ipatch = np.ascontiguousarray(neighbor_ipatch, dtype=np.int64)
        self.particles = []
        self.fields = None
        self.pml = []  # OPml objects in the order the reference adds them (xmin, xmax, ymin, ymax, zmin, zmax)

    def _shrink(self, face, d):  # core/patch/patch.py:105-148: the particle box excludes the PML
        return next((m.thickness * d for m in self.pml if m.face == face), 0.0)
    xmin = property(lambda s: s.x0 + s._shrink("xmin", s.dx))
    xmax = property(lambda s: s.x0 + (s.nx - 1) * s.dx - s._shrink("xmax", s.dx))
    ymin = property(lambda s: s.y0 + s._shrink("ymin", s.dy))
    ymax = property(lambda s: s.y0 + (s.ny - 1) * s.dy - s._shrink("ymax", s.dy))
    zmin = property(lambda s: s.z0 + s._shrink("zmin", s.dz))
    zmax = property(lambda s: s.z0 + (s.nz - 1) * s.dz - s._shrink("zmax", s.dz))


AXIS = {"x": 0, "y": 1, "z": 2}
# (corrected component, sign, differenced source) pairs per axis; core/boundary/cpml.py:527-730
PSI_TABLE = {0: dict(E=[("ey", -1.0, "bz"), ("ez", +1.0, "by")], B=[("by", +1.0, "ez"), ("bz", -1.0, "ey")]),
             1: dict(E=[("ex", +1.0, "bz"), ("ez", -1.0, "bx")], B=[("bx", -1.0, "ez"), ("bz", +1.0, "ex")]),
             2: dict(E=[("ex", -1.0, "by"), ("ey", +1.0, "bx")], B=[("bx", +1.0, "ey"), ("by", -1.0, "ex")])}


class OPml:
    """One CPML face of a patch: coefficient profiles and psi arrays (core/boundary/cpml.py:11-131, 247-340)."""

    def __init__(self, face, n, d, shape, dx, thickness=6, kappa_max=20.0, a_max=0.15, sigma_max=0.7):
        self.face, self.axis, self.side = face, AXIS[face[0]], face[1:]
        self.thickness, self.n, self.d = thickness, n, d
        m, ma = 3, 1
        sig_maxval = sigma_max * C_LIGHT * 0.8 * (m + 1.0) / dx  # the reference scales every face with dx (cpml.py:60)
        self.kappa_e, self.kappa_b = np.ones(n), np.ones(n)
        self.sigma_e, self.sigma_b, self.a_e, self.a_b = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)

        def coeff(pos, sl, kappa, sigma, a):
            kappa[sl] = 1 + (kappa_max - 1) * pos**m
            sigma[sl] = sig_maxval * pos**m
            a[sl] = a_max * (1 - pos)**ma
        t = thickness
        if self.side == "min":
            coeff(1.0 - np.arange(t, dtype=float) / t, np.s_[:t], self.kappa_e, self.sigma_e, self.a_e)
            coeff(1.0 - (np.arange(t, dtype=float) + 0.5) / t, np.s_[:t], self.kappa_b, self.sigma_b, self.a_b)
            self.e_range, self.b_range = (0, t), (0, t)
        else:
            coeff(1.0 - np.arange(t, dtype=float)[::-1] / t, np.s_[n - t:n], self.kappa_e, self.sigma_e, self.a_e)
            coeff(1.0 - (np.arange(t, dtype=float) + 0.5)[::-1] / t, np.s_[n - t - 1:n - 1], self.kappa_b, self.sigma_b, self.a_b)
            self.e_range, self.b_range = (n - t, n), (n - t - 1, n - 1)
        ax = "xyz"[self.axis]
        self.names = dict(E=[f"psi_{c}_{ax}" for c, _, _ in PSI_TABLE[self.axis]["E"]],
                          B=[f"psi_{c}_{ax}" for c, _, _ in PSI_TABLE[self.axis]["B"]])
        for nm in self.names["E"] + self.names["B"]:
            setattr(self, nm, np.zeros(shape))


class OSorter:
    """Per-species sorter work arrays, core/sort/particle_sort.py:20-160."""

    def __init__(self, st: "OState", ispec: int):
        self.ispec = ispec
        self.nxb = st.nx
        self.reverse_x = None
        self.nbuf_last = 0
        n = len(st.patches)
        shape = (st.nx, 1) + ((1,) if st.dim == 3 else ())
        mk = lambda: [np.zeros(shape, dtype=np.int64) for _ in range(n)]  # noqa: E731
        self.bucket_count, self.bound_min, self.bound_max, self.count_not, self.start_counter = mk(), mk(), mk(), mk(), mk()
        self.pidx = [np.full(p.particles[ispec].npart, -1, dtype=np.int64) for p in st.patches]
        self.pref = [np.full(p.particles[ispec].npart, -1, dtype=np.int64) for p in st.patches]
        self.ptarget = [np.full(p.particles[ispec].npart, -1, dtype=np.int64) for p in st.patches]
        self.buf = [np.zeros(p.particles[ispec].npart) for p in st.patches]

    def ensure(self, st):
        for ip, p in enumerate(st.patches):
            n = p.particles[self.ispec].npart
            if self.pidx[ip].size != n:
                self.pidx[ip] = np.full(n, -1, dtype=np.int64)
                self.pref[ip] = np.full(n, -1, dtype=np.int64)
                self.ptarget[ip] = np.full(n, -1, dtype=np.int64)
                self.buf[ip] = np.zeros(n)


class OState:
    """All host state of one rank: patches (fields + particles per species) and scalars."""

    def __init__(self, dim, nx, ny, nz, ng, dx, dy, dz, dt, q, m, x0, y0, z0, neighbor_ipatch, glob):
        self.dim, self.nx, self.ny, self.nz, self.ng = dim, nx, ny, nz, ng
        self.dx, self.dy, self.dz, self.dt = dx, dy, dz, dt
        self.q, self.m = list(q), list(m)
        self.nspec = len(self.q)
        self.glob = np.asarray(glob, dtype=np.float64)  # xmin,xmax,ymin,ymax,zmin,zmax (global particle box)
        self.patches = []
        for i in range(len(x0)):
            p = OPatch(i, x0[i], y0[i], z0[i], nx, ny, nz, dx, dy, dz, neighbor_ipatch[i])
            p.fields = OFields(dim, nx, ny, nz, ng, dx, dy, dz, x0[i], y0[i], z0[i])
            for s in range(self.nspec):
                p.particles.append(OParticles(i))
            self.patches.append(p)
        self.sorters = None
        self.last_migration = None
        self._leak = []

    # ---- construction helpers -------------------------------------------------------------------
    @classmethod
    def from_golden(cls, g, tag: str) -> "OState":
        dim = int(g["meta/dim"])
        st = cls(dim, int(g["meta/nx"]), int(g["meta/ny"]), int(g["meta/nz"]), int(g["meta/n_guard"]),
                 float(g["meta/dx"]), float(g["meta/dy"]), float(g["meta/dz"]), float(g["meta/dt"]),
                 g["meta/q"], g["meta/m"], g["meta/x0"], g["meta/y0"], g["meta/z0"],
                 g["meta/neighbor_ipatch"], g["meta/bounds_global"])
        for ip, p in enumerate(st.patches):
            for a in FIELD_ATTRS:
                setattr(p.fields, a, np.ascontiguousarray(g[f"{tag}/f/{ip}/{a}"]).copy())
            for s in range(st.nspec):
                part = p.particles[s]
                for a in PART_ATTRS:
                    setattr(part, a, np.ascontiguousarray(g[f"{tag}/p/{ip}/{s}/{a}"]).copy())
                part.is_dead = np.ascontiguousarray(g[f"{tag}/p/{ip}/{s}/is_dead"]).astype(bool).copy()
                part.npart = part.x.size
                part._npart_created = part.npart  # capacity == ids handed out (prune is never on the path)
        if "meta/pml_faces" in g.files:
            n_ax = (st.nx, st.ny, st.nz)
            d_ax = (st.dx, st.dy, st.dz)
            shape = (st.nx, st.ny) + ((st.nz,) if dim == 3 else ())
            cls2face = {"PMLXmin": "xmin", "PMLXmax": "xmax", "PMLYmin": "ymin", "PMLYmax": "ymax", "PMLZmin": "zmin", "PMLZmax": "zmax"}
            for ip, p in enumerate(st.patches):
                for ipml, cname in enumerate(str(g["meta/pml_faces"][ip]).split(",")):
                    if not cname:
                        continue
                    face = cls2face[cname]
                    m = OPml(face, n_ax[AXIS[face[0]]], d_ax[AXIS[face[0]]], shape, st.dx, thickness=int(g["meta/cpml_thickness"]))
                    for nm in m.names["E"] + m.names["B"]:
                        getattr(m, nm)[...] = g[f"{tag}/pml/{ip}/{ipml}/{nm}"]
                    p.pml.append(m)
        st.sorters = [OSorter(st, s) for s in range(st.nspec)]
        return st

    def set_reverse_x(self, flags):
        for s, f in zip(self.sorters, flags):
            s.reverse_x = bool(f)

    def clone(self) -> "OState":
        import copy
        leak, self._leak = self._leak, []
        c = copy.deepcopy(self)
        self._leak = leak
        return c

    @property
    def boxes(self):  # migration boxes widened by half a cell, sync_particles_3d.c:402-409
        b = np.zeros((len(self.patches), 6))
        for i, p in enumerate(self.patches):
            b[i] = [p.xmin - 0.5 * self.dx, p.xmax + 0.5 * self.dx, p.ymin - 0.5 * self.dy, p.ymax + 0.5 * self.dy,
                    p.zmin - 0.5 * self.dz, p.zmax + 0.5 * self.dz]
        return b

    @property
    def nbr(self):
        return np.ascontiguousarray(np.stack([p.neighbor_ipatch for p in self.patches]), dtype=np.int64)

    def n_alive(self) -> int:
        return int(sum((~p.particles[s].is_dead).sum() for p in self.patches for s in range(self.nspec)))


# --------------------------------------------------------------------------------------------------
# operators
# --------------------------------------------------------------------------------------------------
def _kappas(st, p, which):
    n_ax = (st.nx, st.ny, st.nz)
    ks = [np.ones(n_ax[a]) for a in range(st.dim)]
    for m in p.pml:  # core/maxwell/solver/solver.py:88-106
        ks[m.axis] = m.kappa_e if which == "e" else m.kappa_b
    return ks


def _advance_psi(st, p, which, dt):
    """pml.advance_e_currents / advance_b_currents for every face of the patch, in list order."""
    L = lib()
    f = p.fields
    for m in p.pml:
        (c1, s1, g1), (c2, s2, g2) = PSI_TABLE[m.axis][which]
        n1, n2 = m.names[which]
        kap, sig, a = (m.kappa_e, m.sigma_e, m.a_e) if which == "E" else (m.kappa_b, m.sigma_b, m.a_b)
        lo, hi = m.e_range if which == "E" else m.b_range
        L.orc_update_psi(_p(getattr(f, c1)), _p(getattr(f, c2)), _p(getattr(f, g1)), _p(getattr(f, g2)),
                         _p(getattr(m, n1)), _p(getattr(m, n2)), _c_dbl(s1), _c_dbl(s2), _p(kap), _p(sig), _p(a),
                         _c_i64(m.axis), _c_i64(0 if which == "E" else 1), _c_i64(lo), _c_i64(hi), _c_i64(st.dim),
                         _c_i64(st.nx), _c_i64(st.ny), _c_i64(st.nz), _c_i64(st.ng), _c_dbl(m.d), _c_dbl(dt))


def update_efield(st: OState, dt: float):
    L = lib()
    bfac, jfac = dt * C_LIGHT**2, dt / EPSILON_0
    if st.dim == 3 and not any(p.pml for p in st.patches):  # all patches in one call, OpenMP over patches
        ptrs = _pp([getattr(p.fields, n) for p in st.patches for n in FIELD_ATTRS[:9]])
        L.orc_update_efield_3d_all(ptrs, _c_i64(len(st.patches)), _c_i64(st.nx), _c_i64(st.ny), _c_i64(st.nz), _c_i64(st.ng),
                                   _c_dbl(st.dx), _c_dbl(st.dy), _c_dbl(st.dz), _c_dbl(bfac), _c_dbl(jfac))
        return
    for p in st.patches:
        f = p.fields
        a = [_p(getattr(f, n)) for n in FIELD_ATTRS[:9]]
        if p.pml:
            ks = [_p(k) for k in _kappas(st, p, "e")]
            if st.dim == 3:
                L.orc_update_efield_cpml_3d(*a, *ks, _c_i64(st.nx), _c_i64(st.ny), _c_i64(st.nz), _c_i64(st.ng),
                                            _c_dbl(st.dx), _c_dbl(st.dy), _c_dbl(st.dz), _c_dbl(bfac), _c_dbl(jfac))
            else:
                L.orc_update_efield_cpml_2d(*a, *ks, _c_i64(st.nx), _c_i64(st.ny), _c_i64(st.ng),
                                            _c_dbl(st.dx), _c_dbl(st.dy), _c_dbl(bfac), _c_dbl(jfac))
        elif st.dim == 3:
            L.orc_update_efield_3d(*a, _c_i64(st.nx), _c_i64(st.ny), _c_i64(st.nz), _c_i64(st.ng),
                                   _c_dbl(st.dx), _c_dbl(st.dy), _c_dbl(st.dz), _c_dbl(bfac), _c_dbl(jfac))
        else:
            L.orc_update_efield_2d(*a, _c_i64(st.nx), _c_i64(st.ny), _c_i64(st.ng),
                                   _c_dbl(st.dx), _c_dbl(st.dy), _c_dbl(bfac), _c_dbl(jfac))
    for p in st.patches:
        _advance_psi(st, p, "E", dt)


def update_bfield(st: OState, dt: float):
    L = lib()
    if st.dim == 3 and not any(p.pml for p in st.patches):
        ptrs = _pp([getattr(p.fields, n) for p in st.patches for n in FIELD_ATTRS[:6]])
        L.orc_update_bfield_3d_all(ptrs, _c_i64(len(st.patches)), _c_i64(st.nx), _c_i64(st.ny), _c_i64(st.nz), _c_i64(st.ng),
                                   _c_dbl(st.dx), _c_dbl(st.dy), _c_dbl(st.dz), _c_dbl(dt))
        return
    for p in st.patches:
        f = p.fields
        a = [_p(getattr(f, n)) for n in FIELD_ATTRS[:6]]
        if p.pml:
            ks = [_p(k) for k in _kappas(st, p, "b")]
            if st.dim == 3:
                L.orc_update_bfield_cpml_3d(*a, *ks, _c_i64(st.nx), _c_i64(st.ny), _c_i64(st.nz), _c_i64(st.ng),
                                            _c_dbl(st.dx), _c_dbl(st.dy), _c_dbl(st.dz), _c_dbl(dt))
            else:
                L.orc_update_bfield_cpml_2d(*a, *ks, _c_i64(st.nx), _c_i64(st.ny), _c_i64(st.ng), _c_dbl(st.dx), _c_dbl(st.dy), _c_dbl(dt))
        elif st.dim == 3:
            L.orc_update_bfield_3d(*a, _c_i64(st.nx), _c_i64(st.ny), _c_i64(st.nz), _c_i64(st.ng),
                                   _c_dbl(st.dx), _c_dbl(st.dy), _c_dbl(st.dz), _c_dbl(dt))
        else:
            L.orc_update_bfield_2d(*a, _c_i64(st.nx), _c_i64(st.ny), _c_i64(st.ng), _c_dbl(st.dx), _c_dbl(st.dy), _c_dbl(dt))
    for p in st.patches:
        _advance_psi(st, p, "B", dt)


def sync_guard_fields(st: OState, attrs, backend="port"):
    n = len(st.patches)
    if backend == "ref":
        mod = ref_module("patch/sync_fields3d" if st.dim == 3 else "patch/sync_fields2d")
        fl = [p.fields for p in st.patches]
        if st.dim == 3:
            mod.sync_guard_fields_3d(fl, st.patches, list(attrs), n, st.nx, st.ny, st.nz, st.ng)
        else:
            mod.sync_guard_fields_2d(fl, st.patches, list(attrs), n, st.nx, st.ny, st.ng)
        return
    nbr = st.nbr
    for a in attrs:
        ptrs = _pp([getattr(p.fields, a) for p in st.patches])
        lib().orc_sync_guard(ptrs, _c_i64(n), _p(nbr), _c_i64(st.dim), _c_i64(st.nx), _c_i64(st.ny), _c_i64(st.nz), _c_i64(st.ng))


def sync_currents(st: OState, backend="port"):
    n = len(st.patches)
    if backend == "ref":
        mod = ref_module("patch/sync_fields3d" if st.dim == 3 else "patch/sync_fields2d")
        fl = [p.fields for p in st.patches]
        if st.dim == 3:
            mod.sync_currents_3d(fl, st.patches, n, st.nx, st.ny, st.nz, st.ng)
        else:
            mod.sync_currents_2d(fl, st.patches, n, st.nx, st.ny, st.ng)
        return
    nbr = st.nbr
    for a in ("jx", "jy", "jz", "rho"):
        ptrs = _pp([getattr(p.fields, a) for p in st.patches])
        lib().orc_sync_currents(ptrs, _c_i64(n), _p(nbr), _c_i64(st.dim), _c_i64(st.nx), _c_i64(st.ny), _c_i64(st.nz), _c_i64(st.ng))


def reset_currents(st: OState):
    for p in st.patches:  # core/current/cpu3d.c:185-240 (memset incl. guards)
        for a in ("jx", "jy", "jz", "rho"):
            getattr(p.fields, a)[...] = 0.0


def decide_reverse_x(st: OState, ispec: int) -> bool:  # core/sort/particle_sort.py:64-89
    w_sum = 0.0
    wux_sum = 0.0
    for p in st.patches:
        part = p.particles[ispec]
        alive = ~part.is_dead
        w_sum += float(part.w[alive].sum())
        wux_sum += float((part.w[alive] * part.ux[alive]).sum())
    return w_sum > 0.0 and wux_sum / w_sum < 0.0


def sort_species(st: OState, ispec: int, backend="port") -> int:
    srt = st.sorters[ispec]
    srt.ensure(st)
    if srt.reverse_x is None:
        srt.reverse_x = decide_reverse_x(st, ispec)
    n = len(st.patches)
    Ly = st.glob[3] - st.glob[2]
    Lz = st.glob[5] - st.glob[4]
    x0s = [p.x0 - st.dx / 2 for p in st.patches]
    y0s = [p.y0 - st.dy / 2 for p in st.patches]
    z0s = [p.z0 - st.dz / 2 for p in st.patches]
    parts = [p.particles[ispec] for p in st.patches]
    if backend == "ref":
        attrs_list = [getattr(pt, a) for pt in parts for a in PART_ATTRS]
        if st.dim == 3:
            nbuf = ref_module("sort/cpu3d").sort_particles_patches_3d(
                [pt.x for pt in parts], [pt.y for pt in parts], [pt.z for pt in parts], [pt.is_dead for pt in parts],
                attrs_list, x0s, y0s, z0s, srt.nxb, 1, 1, st.dx, Ly, Lz, n,
                srt.bucket_count, srt.bound_min, srt.bound_max, srt.count_not, srt.start_counter,
                srt.pidx, srt.pref, srt.ptarget, srt.buf, int(srt.reverse_x))
        else:
            nbuf = ref_module("sort/cpu2d").sort_particles_patches_2d(
                [pt.x for pt in parts], [pt.y for pt in parts], [pt.is_dead for pt in parts],
                attrs_list, x0s, y0s, srt.nxb, 1, st.dx, Ly, n,
                srt.bucket_count, srt.bound_min, srt.bound_max, srt.count_not, srt.start_counter,
                srt.pidx, srt.pref, srt.ptarget, srt.buf, int(srt.reverse_x))
    else:
        nbuf = 0
        for ip, pt in enumerate(parts):
            attrs = [getattr(pt, a) for a in PART_ATTRS]
            dead = pt.is_dead.view(np.uint8)
            nbuf += lib().orc_sort_patch(
                _p(pt.x), _p(pt.y), _p(pt.z) if st.dim == 3 else None, _p(dead), _pp(attrs), _c_i64(len(attrs)),
                _c_i64(pt.npart), _c_i64(srt.nxb), _c_i64(1), _c_i64(1), _c_dbl(st.dx), _c_dbl(Ly), _c_dbl(Lz if st.dim == 3 else 1.0),
                _c_dbl(x0s[ip]), _c_dbl(y0s[ip]), _c_dbl(z0s[ip]), ctypes.c_int(int(srt.reverse_x)),
                _p(srt.bucket_count[ip]), _p(srt.bound_min[ip]), _p(srt.bound_max[ip]),
                _p(srt.pidx[ip]), _p(srt.pref[ip]), _p(srt.ptarget[ip]))
    srt.nbuf_last = int(nbuf)
    return srt.nbuf_last


def push_deposit(st: OState, ispec: int, backend="port"):
    q, m, dt = float(st.q[ispec]), float(st.m[ispec]), st.dt
    n = len(st.patches)
    if backend == "ref":
        parts = [p.particles[ispec] for p in st.patches]
        fl = [p.fields for p in st.patches]
        if st.dim == 3:
            ref_module("pusher/unified/unified_pusher_3d").unified_boris_pusher_cpu_3d(parts, fl, n, dt, q, m)
        else:
            ref_module("pusher/unified/unified_pusher_2d").unified_boris_pusher_cpu_2d(parts, fl, n, dt, q, m)
        return
    L = lib()
    for p in st.patches:
        pt, f = p.particles[ispec], p.fields
        part = _pp([getattr(pt, a) for a in PART_ATTRS[8:14]])
        fa = [_p(getattr(f, a)) for a in FIELD_ATTRS]
        dead = pt.is_dead.view(np.uint8)
        if st.dim == 3:
            L.orc_push_deposit_3d(_p(pt.x), _p(pt.y), _p(pt.z), _p(pt.ux), _p(pt.uy), _p(pt.uz), _p(pt.inv_gamma),
                                  _p(pt.w), _p(dead), part, _c_i64(pt.npart), *fa,
                                  _c_i64(st.nx), _c_i64(st.ny), _c_i64(st.nz), _c_i64(st.ng),
                                  _c_dbl(st.dx), _c_dbl(st.dy), _c_dbl(st.dz), _c_dbl(p.x0), _c_dbl(p.y0), _c_dbl(p.z0),
                                  _c_dbl(dt), _c_dbl(q), _c_dbl(m))
        else:
            L.orc_push_deposit_2d(_p(pt.x), _p(pt.y), _p(pt.ux), _p(pt.uy), _p(pt.uz), _p(pt.inv_gamma),
                                  _p(pt.w), _p(dead), part, _c_i64(pt.npart), *fa,
                                  _c_i64(st.nx), _c_i64(st.ny), _c_i64(st.ng), _c_dbl(st.dx), _c_dbl(st.dy),
                                  _c_dbl(p.x0), _c_dbl(p.y0), _c_dbl(dt), _c_dbl(q), _c_dbl(m))


def sync_particles(st: OState, backend="port"):
    """core/patch/patch.py:705-764: count -> extend (host) -> fill, per species.  Returns per-species dicts."""
    n = len(st.patches)
    nb = 26 if st.dim == 3 else 8
    out = []
    for s in range(st.nspec):
        parts = [p.particles[s] for p in st.patches]
        if backend == "ref":
            mod = ref_module("patch/sync_particles_3d" if st.dim == 3 else "patch/sync_particles_2d")
            if st.dim == 3:
                ext, inc, outg, alive = mod.get_npart_to_extend_3d(parts, st.patches, n, st.dx, st.dy, st.dz)
            else:
                ext, inc, outg, alive = mod.get_npart_to_extend_2d(parts, st.patches, n, st.dx, st.dy)
            rec = dict(to_extend=ext.copy(), incoming=inc.copy(), outgoing=outg.copy(), alive=alive.copy())
            for ip, pt in enumerate(parts):
                pt.extend(int(ext[ip]))
            g = st.glob
            if st.dim == 3:
                mod.fill_particles_from_boundary_3d(parts, st.patches, inc, outg, n, st.dx, st.dy, st.dz,
                                                    g[0], g[1], g[2], g[3], g[4], g[5], list(PART_ATTRS))
            else:
                mod.fill_particles_from_boundary_2d(parts, st.patches, inc, outg, n, st.dx, st.dy,
                                                    g[0], g[1], g[2], g[3], list(PART_ATTRS))
            # The reference frees the data pointers of `inc`/`outg` itself (AUTOFREE on PyArray_DATA,
            # sync_particles_3d.c:537-538); keep numpy from releasing them a second time.
            for a in (inc, outg):
                ctypes.pythonapi.Py_IncRef(ctypes.py_object(a))
        else:
            L = lib()
            box, nbr = st.boxes, st.nbr
            npart = np.array([pt.npart for pt in parts], dtype=np.int64)
            outg = np.zeros(n * nb, dtype=np.int64)
            inc = np.zeros(n, dtype=np.int64)
            ext = np.zeros(n, dtype=np.int64)
            alive = np.zeros(n, dtype=np.int64)
            deads = [pt.is_dead.view(np.uint8) for pt in parts]
            L.orc_migrate_count(_pp([pt.x for pt in parts]), _pp([pt.y for pt in parts]), _pp([pt.z for pt in parts]),
                                _pp(deads), _p(npart), _c_i64(n), _c_i64(st.dim), _p(box), _p(nbr),
                                _p(outg), _p(inc), _p(ext), _p(alive))
            rec = dict(to_extend=ext.copy(), incoming=inc.copy(), outgoing=outg.copy(), alive=alive.copy())
            for ip, pt in enumerate(parts):
                pt.extend(int(ext[ip]))
            npart = np.array([pt.npart for pt in parts], dtype=np.int64)
            deads = [pt.is_dead.view(np.uint8) for pt in parts]
            attrs = [getattr(pt, a) for pt in parts for a in PART_ATTRS]
            L.orc_migrate_fill(_pp(attrs), _c_i64(len(PART_ATTRS)), _c_i64(0), _c_i64(1), _c_i64(2 if st.dim == 3 else -1),
                               _pp(deads), _p(npart), _c_i64(n), _c_i64(st.dim), _p(box), _p(nbr), _p(st.glob),
                               _c_dbl(st.dx), _c_dbl(st.dy), _c_dbl(st.dz), _p(inc))
        out.append(rec)
    st.last_migration = out
    return out


def update_lists(st: OState):
    """simulation.py:781-824: sorter index arrays of patches whose particle arrays were extended are re-created (-1)."""
    for s in range(st.nspec):
        st.sorters[s].ensure(st)
        for p in st.patches:
            p.particles[s].extended = False


def laser_stage(st: OState, sources, laserpos):
    """Stage `_laser` (callback/laser.py:109-137, 171-186, 218-241): rewrite B at the antenna plane of the xmin edge
    patches from the given source planes.  sources: {ipatch: (ey_src, ez_src)} shaped like one padded x-plane."""
    L = lib()
    for ip, (ey_s, ez_s) in sources.items():
        p = st.patches[ip]
        f = p.fields
        iy0, iy1, iz0, iz1 = 0, st.ny, 0, st.nz
        for m in p.pml:
            if m.face == "ymin": iy0 = m.thickness
            if m.face == "ymax": iy1 = st.ny - m.thickness
            if m.face == "zmin": iz0 = m.thickness
            if m.face == "zmax": iz1 = st.nz - m.thickness
        ey_s, ez_s = np.ascontiguousarray(ey_s, dtype=np.float64), np.ascontiguousarray(ez_s, dtype=np.float64)
        L.orc_laser_bfields(*[_p(getattr(f, a)) for a in FIELD_ATTRS[:9]], _c_i64(st.dim), _c_i64(st.nx), _c_i64(st.ny), _c_i64(st.nz),
                            _c_i64(st.ng), _c_dbl(st.dx), _c_dbl(st.dy), _c_dbl(st.dz), _c_dbl(st.dt), _c_i64(laserpos),
                            _c_i64(iy0), _c_i64(iy1), _c_i64(iz0), _c_i64(iz1), _p(ey_s), _p(ez_s))


def step(st: OState, backend="port", laser=None):
    """One full PIC step, simulation/simulation.py:937-1130 (periodic, unified pusher, no callbacks)."""
    dt = st.dt
    E, B = ("ex", "ey", "ez"), ("bx", "by", "bz")
    update_efield(st, 0.5 * dt); sync_guard_fields(st, E, backend)
    update_bfield(st, 0.5 * dt); sync_guard_fields(st, B, backend)
    for s in range(st.nspec):
        sort_species(st, s, backend)
    reset_currents(st)
    for s in range(st.nspec):
        push_deposit(st, s, backend)
    sync_currents(st, backend)
    sync_particles(st, backend)
    update_lists(st)
    update_bfield(st, 0.5 * dt)
    if laser is not None:  # stage `_laser` sits between the B update and its guard sync (simulation.py:1098-1103)
        laser_stage(st, *laser)
    sync_guard_fields(st, B, backend)
    update_efield(st, 0.5 * dt); sync_guard_fields(st, E, backend)


# --------------------------------------------------------------------------------------------------
# MovingWindow (callback/utils.py:471-840), stage `start`
# --------------------------------------------------------------------------------------------------
DIR2 = [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (-1, -1, 0), (1, -1, 0), (-1, 1, 0), (1, 1, 0)]  # Boundary2D order


def _boundary3d_dirs():
    """Offsets in Boundary3D enum order (core/patch/patch.py:37-69): faces, xy / xz / yz edges, corners."""
    faces = [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]
    xy = [(-1, -1, 0), (-1, 1, 0)]
    xz = [(-1, 0, -1), (-1, 0, 1)]
    xy2 = [(1, -1, 0), (1, 1, 0)]
    xz2 = [(1, 0, -1), (1, 0, 1)]
    yz = [(0, -1, -1), (0, -1, 1), (0, 1, -1), (0, 1, 1)]
    corners = [(-1, -1, -1), (-1, -1, 1), (-1, 1, -1), (-1, 1, 1), (1, -1, -1), (1, -1, 1), (1, 1, -1), (1, 1, 1)]
    return faces + xy + xz + xy2 + xz2 + yz + corners


class OMovingWindow:
    """Host bookkeeping of the moving window restated on an OState.

    npatch: (npx, npy[, npz]); ipatch: per patch (ix, iy, iz) at construction; periodic: per axis; density, ppc: callables
    of the node coordinates per species (None: species is not re-loaded); rand_gen: the rank's generator AFTER the initial
    load spawned its per-patch children (simulation.py:700-716, patch.py:864-907)."""

    def __init__(self, st: OState, velocity, npatch, ipatch, periodic, Lx, density, ppc, rand_gen, start_time=None,
                 inject_particles=True, stop_inject_time=None, density_min=0.0):
        self.velocity, self.start_time = velocity, start_time
        self.inject_particles, self.stop_inject_time = inject_particles, stop_inject_time
        self.npatch, self.periodic, self.Lx = tuple(npatch), tuple(periodic), Lx
        self.density, self.ppc, self.density_min, self.rand_gen = density, ppc, density_min, rand_gen
        self.total_shift = self.patch_this_shift = None
        self.num_shifts = 0
        self.dirs = _boundary3d_dirs() if st.dim == 3 else DIR2
        for p, ip in zip(st.patches, ipatch):
            p.ipatch = [int(v) for v in ip]
            p.xaxis = np.arange(st.nx) * st.dx + p.x0
            p.yaxis = np.arange(st.ny) * st.dy + p.y0
            p.zaxis = np.arange(st.nz) * st.dz + p.z0

    def stage(self, st: OState, time: float):
        patch_Lx = st.nx * st.dx
        if self.start_time is None:
            self.start_time = self.Lx / C_LIGHT
        if self.total_shift is None:
            self.total_shift = patch_Lx
        if self.patch_this_shift is None:
            self.patch_this_shift = patch_Lx
        if time < self.start_time:
            return
        if self.num_shifts == 0:  # callback/utils.py:545-552: the x faces stop absorbing
            for p in st.patches:
                p.pml = [m for m in p.pml if m.axis != 0]
        v = self.velocity(time) if callable(self.velocity) else self.velocity
        self.total_shift += v * st.dt
        self.patch_this_shift += v * st.dt
        self.num_shifts += 1
        if self.patch_this_shift >= patch_Lx:
            direction = 1
            self.patch_this_shift -= patch_Lx
        elif self.patch_this_shift <= -patch_Lx:
            direction = -1
            self.patch_this_shift += patch_Lx
        else:
            return
        last = self.npatch[0] - 1
        new = []
        for p in st.patches:  # callback/utils.py:591-646
            if (direction > 0 and p.ipatch[0] == 0) or (direction < 0 and p.ipatch[0] == last):
                p.ipatch[0] = last if direction > 0 else 0
                p.x0 += direction * self.Lx
                p.xaxis += direction * self.Lx
                p.fields.x0 = p.fields.x0 + direction * self.Lx
                new.append(p)
            else:
                p.ipatch[0] -= direction
        self._neighbours(st)
        self._reload(st, new, time)
        for p in new:
            for a in FIELD_ATTRS:
                getattr(p.fields, a).fill(0.0)
            for m in p.pml:
                for nm in m.names["E"] + m.names["B"]:
                    getattr(m, nm).fill(0.0)

    def _neighbours(self, st):  # core/patch/patch.py:446-592, 641-667 (single rank: index -> list position)
        dim = st.dim
        where = {tuple(p.ipatch[:dim]): i for i, p in enumerate(st.patches)}
        for p in st.patches:
            nb = np.full(len(self.dirs), -1, dtype=np.int64)
            for b, off in enumerate(self.dirs):
                pos, ok = [], True
                for a in range(dim):
                    n = p.ipatch[a] + off[a]
                    if n < 0 or n >= self.npatch[a]:
                        if not self.periodic[a]:
                            ok = False
                            break
                        n %= self.npatch[a]
                    pos.append(n)
                if ok:
                    nb[b] = where[tuple(pos)]
            p.neighbor_ipatch = nb

    def _reload(self, st, new, time):  # callback/utils.py:718-840 + core/patch/cpu.py:6-99
        if not new or not self.inject_particles:
            return
        if self.stop_inject_time is not None and time >= self.stop_inject_time:
            return
        dim = st.dim
        # the loader derives the spacings from the axes of the first patch it is given (core/patch/cpu.py:23-24,70-72)
        d = tuple(float(ax[1] - ax[0]) for ax in (new[0].xaxis, new[0].yaxis, new[0].zaxis)[:dim])
        for s in range(st.nspec):
            if self.density[s] is None:
                continue
            gens = self.rand_gen.spawn(len(new))
            for k, p in enumerate(new):
                axes = (p.xaxis, p.yaxis, p.zaxis)[:dim]
                grids = np.meshgrid(*axes, indexing="ij")
                dens = np.broadcast_to(np.asarray(self.density[s](*grids), dtype=float), grids[0].shape).ravel()
                ppc = np.broadcast_to(np.asarray(self.ppc[s](*grids), dtype=float), grids[0].shape).astype(np.int64).ravel()
                ppc = np.where(dens > self.density_min, ppc, 0)
                n = int(ppc.sum())
                part = p.particles[s]
                # ParticlesBase.initialize (core/particles.py:118-139): fresh arrays, ids continue the patch's counter
                part.npart = n
                for a in PART_ATTRS:
                    setattr(part, a, np.zeros(n))
                part.inv_gamma[:] = 1
                part.is_dead = np.zeros(n, dtype=bool)
                part._id[:] = part._ids(part._npart_created, n)
                part._npart_created += n
                part.extended = True
                if n == 0:
                    continue
                u = gens[k].random(dim * n)  # per node: ppc draws for x, then y(, then z)
                first = np.concatenate([[0], np.cumsum(ppc)[:-1]])
                node_of = np.repeat(np.arange(ppc.size), ppc)
                j = np.arange(n) - first[node_of]
                nodes = [g.ravel() for g in grids]
                for a, name in enumerate("xyz"[:dim]):
                    pos = dim * first[node_of] + a * ppc[node_of] + j
                    low, rng = -d[a] / 2, d[a] / 2 - (-d[a] / 2)
                    getattr(part, name)[:] = (low + rng * u[pos]) + nodes[a][node_of]
                wnode = dens.copy()
                for da in d:
                    wnode = wnode * da
                part.w[:] = wnode[node_of] / ppc[node_of]
        update_lists(st)
