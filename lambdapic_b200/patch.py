"""Patch containers with the reference's names (core/patch/patch.py:24-764).  Geometry and neighbour tables live
on the host; fields and particles are views into the device mirrors owned by :class:`lambdapic_b200.device.DeviceBridge`."""
from __future__ import annotations

from enum import IntEnum, auto

import numpy as np

from .particles import ParticlesBase
from .species import Species


class Boundary2D(IntEnum):  # core/patch/patch.py:24-34
    XMIN = 0
    XMAX = auto()
    YMIN = auto()
    YMAX = auto()
    XMINYMIN = auto()
    XMAXYMIN = auto()
    XMINYMAX = auto()
    XMAXYMAX = auto()


class Boundary3D(IntEnum):  # core/patch/patch.py:37-69
    XMIN = 0
    XMAX = auto()
    YMIN = auto()
    YMAX = auto()
    ZMIN = auto()
    ZMAX = auto()
    XMINYMIN = auto()
    XMINYMAX = auto()
    XMINZMIN = auto()
    XMINZMAX = auto()
    XMAXYMIN = auto()
    XMAXYMAX = auto()
    XMAXZMIN = auto()
    XMAXZMAX = auto()
    YMINZMIN = auto()
    YMINZMAX = auto()
    YMAXZMIN = auto()
    YMAXZMAX = auto()
    XMINYMINZMIN = auto()
    XMINYMINZMAX = auto()
    XMINYMAXZMIN = auto()
    XMINYMAXZMAX = auto()
    XMAXYMINZMIN = auto()
    XMAXYMINZMAX = auto()
    XMAXYMAXZMIN = auto()
    XMAXYMAXZMAX = auto()


class Patch:
    def __init__(self):
        self.particles: list[ParticlesBase] = []
        self.pml_boundary = []
        self.fields = None

    # particle box of the patch; a CPML face shrinks it by its thickness (core/patch/patch.py:105-148)
    def _pml_cells(self, face):
        return next((m.thickness for m in self.pml_boundary if m.face == face), 0)
    xmin = property(lambda s: s.x0 + s._pml_cells("xmin") * s.dx)
    xmax = property(lambda s: s.x0 + (s.nx - 1) * s.dx - s._pml_cells("xmax") * s.dx)
    ymin = property(lambda s: s.y0 + s._pml_cells("ymin") * s.dy)
    ymax = property(lambda s: s.y0 + (s.ny - 1) * s.dy - s._pml_cells("ymax") * s.dy)
    zmin = property(lambda s: s.z0 + s._pml_cells("zmin") * s.dz)
    zmax = property(lambda s: s.z0 + (s.nz - 1) * s.dz - s._pml_cells("zmax") * s.dz)

    def add_pml_boundary(self, pml) -> None:
        """core/patch/patch.py:185-187, 291-298, 382-386: at most one face per axis."""
        assert all(m.axis != pml.axis for m in self.pml_boundary), "a patch cannot hold two PML faces of one axis; use more patches"
        self.pml_boundary.append(pml)

    def add_particles(self, particles: ParticlesBase) -> None:
        self.particles.append(particles)

    def set_fields(self, fields) -> None:
        self.fields = fields


class Patch2D(Patch):
    def __init__(self, rank, index, ipatch_x, ipatch_y, x0, y0, nx, ny, dx, dy):
        super().__init__()
        self.rank, self.index, self.ipatch_x, self.ipatch_y = rank, index, ipatch_x, ipatch_y
        self.x0, self.y0, self.nx, self.ny, self.dx, self.dy = x0, y0, nx, ny, dx, dy
        self.xaxis = np.arange(nx) * dx + x0
        self.yaxis = np.arange(ny) * dy + y0
        n = len(Boundary2D)
        self.neighbor_index = np.full(n, -1, dtype=int)
        self.neighbor_rank = np.full(n, -1, dtype=int)
        self.neighbor_ipatch = np.full(n, -1, dtype=int)


class Patch3D(Patch):
    def __init__(self, rank, index, ipatch_x, ipatch_y, ipatch_z, x0, y0, z0, nx, ny, nz, dx, dy, dz):
        super().__init__()
        self.rank, self.index = rank, index
        self.ipatch_x, self.ipatch_y, self.ipatch_z = ipatch_x, ipatch_y, ipatch_z
        self.x0, self.y0, self.z0 = x0, y0, z0
        self.nx, self.ny, self.nz, self.dx, self.dy, self.dz = nx, ny, nz, dx, dy, dz
        self.xaxis = np.arange(nx) * dx + x0
        self.yaxis = np.arange(ny) * dy + y0
        self.zaxis = np.arange(nz) * dz + z0
        n = len(Boundary3D)
        self.neighbor_index = np.full(n, -1, dtype=int)
        self.neighbor_rank = np.full(n, -1, dtype=int)
        self.neighbor_ipatch = np.full(n, -1, dtype=int)


class Patches:
    """Container of this rank's patches; the ``sync_*`` methods are the drop-in boundary of
    core/patch/patch.py:670-764 and run on the GPU through the attached DeviceBridge."""

    def __init__(self, dimension: int) -> None:
        self.dimension = dimension
        self.patches: list[Patch] = []
        self.species: list[Species] = []
        self.npatches = 0
        self.xmin_global = self.xmax_global = self.ymin_global = self.ymax_global = None
        self.zmin_global = self.zmax_global = None
        self._bridge = None
        self._comm = None

    def __getitem__(self, i):
        return self.patches[i]

    def __len__(self):
        return self.npatches

    def __iter__(self):
        return iter(self.patches)

    def append(self, patch: Patch):
        self.patches.append(patch)
        self.npatches += 1

    nx = property(lambda s: s.patches[0].nx)
    ny = property(lambda s: s.patches[0].ny)
    nz = property(lambda s: s.patches[0].nz)
    dx = property(lambda s: s.patches[0].dx)
    dy = property(lambda s: s.patches[0].dy)
    dz = property(lambda s: s.patches[0].dz)
    n_guard = property(lambda s: s.patches[0].fields.n_guard)

    def update_lists(self):
        pass

    def _need_bridge(self):
        if self._bridge is None:
            raise RuntimeError("patches are not attached to a device (Simulation.initialize() does that); "
                               "lambdapic_b200 has no CPU path")
        return self._bridge

    def sync_guard_fields(self, attrs=("ex", "ey", "ez", "bx", "by", "bz")):
        self._need_bridge().sync_guard_fields(list(attrs))

    def sync_currents(self):
        self._need_bridge().sync_currents()

    def sync_particles(self):
        """Returns npart_to_extend summed over species, per patch (int64), like the reference."""
        return self._need_bridge().sync_particles()

    # ---- neighbour tables (core/patch/patch.py:446-667); used again after a MovingWindow shift ----------------------
    def _init_rect_neighbor_index(self, npatch, boundary_conditions, patch_index_map):
        from .workloads import DIR2, DIR3
        dim = self.dimension
        dirs = DIR3 if dim == 3 else DIR2
        axes = "xyz"[:dim]
        if not patch_index_map:
            patch_index_map = {tuple(getattr(p, f"ipatch_{a}") for a in axes): p.index for p in self.patches}
        ntot = int(np.prod(npatch[:dim]))
        for p in self.patches:
            p.neighbor_index.fill(-1)
            for b, off in enumerate(dirs):
                pos, ok = [], True
                for a, ax in enumerate(axes):
                    n = getattr(p, f"ipatch_{ax}") + off[a]
                    if n < 0:
                        if boundary_conditions[f"{ax}min"] != "periodic":
                            ok = False
                            break
                        n = npatch[a] - 1
                    elif n >= npatch[a]:
                        if boundary_conditions[f"{ax}max"] != "periodic":
                            ok = False
                            break
                        n = 0
                    pos.append(n)
                if not ok:
                    continue
                idx = patch_index_map.get(tuple(pos))
                if idx is None:
                    if len(patch_index_map) == ntot:
                        raise KeyError(tuple(pos))
                    continue
                p.neighbor_index[b] = idx

    def init_rect_neighbor_index_2d(self, npatch_x, npatch_y, *, boundary_conditions, patch_index_map=None):
        self._init_rect_neighbor_index((npatch_x, npatch_y), boundary_conditions, patch_index_map or {})

    def init_rect_neighbor_index_3d(self, npatch_x, npatch_y, npatch_z, boundary_conditions, patch_index_map=None):
        self._init_rect_neighbor_index((npatch_x, npatch_y, npatch_z), boundary_conditions, patch_index_map or {})

    def _init_neighbor_ipatch(self):
        """Local list position of every same-rank neighbour (core/patch/patch.py:641-667)."""
        where = {p.index: ip for ip, p in enumerate(self.patches)}
        for p in self.patches:
            p.neighbor_ipatch.fill(-1)
            for b, idx in enumerate(p.neighbor_index):
                if idx >= 0 and idx in where:
                    p.neighbor_ipatch[b] = where[idx]
    init_neighbor_ipatch_2d = init_neighbor_ipatch_3d = _init_neighbor_ipatch

    def _init_neighbor_rank(self, patch_rank_map=None):
        patch_rank_map = dict(patch_rank_map or {})
        for p in self.patches:
            patch_rank_map[p.index] = p.rank
        for p in self.patches:
            p.neighbor_rank.fill(-1)
            for b, idx in enumerate(p.neighbor_index):
                if idx >= 0 and patch_rank_map[idx] != p.rank:
                    p.neighbor_rank[b] = patch_rank_map[idx]
    init_neighbor_rank_2d = init_neighbor_rank_3d = _init_neighbor_rank

    # ---- particle loading (host side so that seeds reproduce the reference's positions) ---------------------
    def calculate_npart(self, species: Species):
        """core/patch/patch.py:796-844 + core/patch/cpu.py:6-19,46-63: sum of int(ppc) over nodes with density > min."""
        dim = self.dimension
        out = np.zeros(self.npatches, dtype=np.int64)
        if species.density is None:
            return out
        species.density_jit = Species.compile_profile(species.density, dim)
        species.ppc_jit = Species.compile_profile(species.ppc, dim)
        for ip, p in enumerate(self.patches):
            _, ppc = _node_profiles(species, p, dim)
            out[ip] = int(ppc.sum())
        return out

    def add_species(self, species: Species, aux_attrs=None) -> int:
        npart = self.calculate_npart(species)
        for ip, p in enumerate(self.patches):
            part = ParticlesBase(ipatch=p.index, rank=p.rank)
            part.attrs += list(aux_attrs or [])
            part.initialize(int(npart[ip]))
            p.add_particles(part)
        self.species.append(species)
        return int(npart.sum())

    def fill_particles(self, rand_gen: np.random.Generator):
        """core/patch/patch.py:864-907 + core/patch/cpu.py:22-43,66-99: per patch generator, node by node
        x, y(, z) = uniform(-d/2, d/2, ppc) + node, w = density*dV/ppc -- drawn in that order so that the
        random stream matches the reference's numba loop exactly."""
        gens = rand_gen.spawn(self.npatches)
        for ispec, s in enumerate(self.species):
            if s.density is None:
                continue
            d = loader_spacing(self.patches[0], self.dimension)
            for ip, p in enumerate(self.patches):
                load_patch_particles(s, p, p.particles[ispec], gens[ip], self.dimension, d)


def loader_spacing(first_patch: Patch, dim: int):
    """The loader takes dx, dy(, dz) from the axes of the FIRST patch of the list it is given (core/patch/cpu.py:23-24,
    70-72): exact for a patch at the origin, 1 ulp(x0)-level off after a MovingWindow shifted the axes in place."""
    axes = (first_patch.xaxis, first_patch.yaxis) + ((first_patch.zaxis,) if dim == 3 else ())
    return tuple(float(ax[1] - ax[0]) for ax in axes)


def load_patch_particles(s: Species, p: Patch, part: ParticlesBase, gen: np.random.Generator, dim: int, d) -> int:
    """Positions and weights of one (patch, species) from the patch's generator (core/patch/cpu.py:22-43,66-99).
    `part` must already hold int(ppc.sum()) slots; d = loader_spacing(first patch of the list being loaded).
    Returns the number of particles loaded."""
    dens, ppc = _node_profiles(s, p, dim)
    n = int(ppc.sum())
    if n == 0:
        return 0
    grids = np.meshgrid(*((p.xaxis, p.yaxis) + ((p.zaxis,) if dim == 3 else ())), indexing="ij")
    ppc_f, dens_f = ppc.ravel(), dens.ravel()
    nodes = [g.ravel() for g in grids]
    u = gen.random(dim * n)
    # stream layout: node-major, then axis, then the ppc draws of that axis
    first = np.concatenate([[0], np.cumsum(ppc_f)[:-1]])            # first particle of each node
    node_of = np.repeat(np.arange(ppc_f.size), ppc_f)               # node of each particle
    k = np.arange(n) - first[node_of]                               # index inside the node
    for a, name in enumerate(("x", "y", "z")[:dim]):
        pos = dim * first[node_of] + a * ppc_f[node_of] + k
        low, rng = -d[a] / 2, d[a] / 2 - (-d[a] / 2)
        getattr(part, name)[:n] = (low + rng * u[pos]) + nodes[a][node_of]
    wnode = dens_f.copy()
    for da in d:  # dens*dx*dy*dz / ppc in the reference's association (core/patch/cpu.py:43,99)
        wnode = wnode * da
    part.w[:n] = wnode[node_of] / ppc_f[node_of]
    return n


_NUMBA_CACHE: dict = {}


def _numba_nodes(fun, axes):
    """Evaluate a user profile node by node with numba, as the reference does (core/species.py:141-170 njit-compiles the
    profile and core/patch/cpu.py calls it per node): same scalar code path, so also the same last bits for profiles that
    use transcendental functions, and ~100x faster than a Python loop for profiles written with `if`.  Returns None when
    numba is missing or cannot compile the function (the caller falls back to numpy)."""
    key = (id(fun), len(axes))
    if key not in _NUMBA_CACHE:
        try:
            import numba
            jf = fun if isinstance(fun, numba.core.dispatcher.Dispatcher) else numba.njit(fun)
            if len(axes) == 2:
                @numba.njit
                def drv(f, x, y, out):
                    for i in range(x.size):
                        for j in range(y.size):
                            out[i, j] = f(x[i], y[j])
            else:
                @numba.njit
                def drv(f, x, y, z, out):
                    for i in range(x.size):
                        for j in range(y.size):
                            for k in range(z.size):
                                out[i, j, k] = f(x[i], y[j], z[k])
            probe = np.zeros(tuple(1 for _ in axes))
            drv(jf, *[np.ascontiguousarray(a[:1], dtype=np.float64) for a in axes], probe)  # compile now, fail now
            _NUMBA_CACHE[key] = (fun, jf, drv)  # `fun` is kept alive so that its id() stays unique
        except Exception:
            _NUMBA_CACHE[key] = (fun, None, None)
    _, jf, drv = _NUMBA_CACHE[key]
    if jf is None:
        return None
    out = np.empty(tuple(len(a) for a in axes))
    drv(jf, *[np.ascontiguousarray(a, dtype=np.float64) for a in axes], out)
    return out


def _node_profiles(species: Species, p: Patch, dim: int):
    """density and int(ppc) on the nodes of patch p (0 where density <= density_min)."""
    axes = (p.xaxis, p.yaxis) + ((p.zaxis,) if dim == 3 else ())
    shape = tuple(len(a) for a in axes)
    dfun, pfun = species.density_jit, species.ppc_jit
    grids = None

    def evaluate(fun, user_fun):
        nonlocal grids
        if callable(user_fun):  # user-written profile: the reference's numba path first
            v = _numba_nodes(fun, axes)
            if v is not None:
                return v
        if grids is None:
            grids = np.meshgrid(*axes, indexing="ij")
        try:
            v = np.asarray(fun(*grids), dtype=float)
            if v.shape == shape:
                return v
            if v.shape == ():
                return np.full(shape, float(v))
        except Exception:
            pass
        return np.vectorize(fun, otypes=[float])(*grids)
    dens = evaluate(dfun, species.density)
    ppc = evaluate(pfun, species.ppc).astype(np.int64)
    ppc = np.where(dens > species.density_min, ppc, 0)
    return dens, ppc
