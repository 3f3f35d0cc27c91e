"""Synthetic uniform thermal plasma workloads (BASELINE.json configs[0] and configs[4]) and the rectangular
periodic patch decomposition they run on.  Pure host-side geometry: no compute happens here.

Patch numbering and neighbour tables follow the reference (simulation/simulation.py:467-502,1378-1406 and
core/patch/patch.py:446-592,641-667): patch index = ix + iy*npx (+ iz*npx*npy), neighbours listed in
Boundary2D/3D enum order, -1 where a non-periodic edge has no neighbour.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

C_LIGHT = 299792458.0
E_CHARGE = 1.602176634e-19
M_E = 9.1093837139e-31
M_P = 1.67262192595e-27

DIR3 = [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1),
        (-1, -1, 0), (-1, 1, 0), (-1, 0, -1), (-1, 0, 1), (1, -1, 0), (1, 1, 0), (1, 0, -1), (1, 0, 1),
        (0, -1, -1), (0, -1, 1), (0, 1, -1), (0, 1, 1),
        (-1, -1, -1), (-1, -1, 1), (-1, 1, -1), (-1, 1, 1), (1, -1, -1), (1, -1, 1), (1, 1, -1), (1, 1, 1)]
DIR2 = [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (-1, -1, 0), (1, -1, 0), (-1, 1, 0), (1, 1, 0)]


@dataclass
class PatchGrid:
    """Rectangular patch decomposition of a (sub)domain; `rank_of` maps global patch -> owning rank."""
    dim: int
    npx: int
    npy: int
    npz: int
    nx: int          # cells per patch
    ny: int
    nz: int
    dx: float
    dy: float
    dz: float
    n_guard: int = 3
    periodic: tuple = (True, True, True)
    index: np.ndarray = field(default=None)      # global patch indices owned here (ascending)
    x0: np.ndarray = field(default=None)
    y0: np.ndarray = field(default=None)
    z0: np.ndarray = field(default=None)
    neighbor_index: np.ndarray = field(default=None)   # global index of each neighbour (-1 none)
    neighbor_ipatch: np.ndarray = field(default=None)  # local position of the neighbour, -1 if remote/none
    neighbor_rank: np.ndarray = field(default=None)    # owning rank if remote, -1 if local/none
    rank: int = 0
    nranks: int = 1

    @property
    def npatch(self):
        return len(self.index)

    @property
    def glob(self):
        """Global particle box [-d/2, L-d/2] (simulation/simulation.py:425-430)."""
        Lx, Ly, Lz = self.npx * self.nx * self.dx, self.npy * self.ny * self.dy, self.npz * self.nz * self.dz
        return np.array([-self.dx / 2, Lx - self.dx / 2, -self.dy / 2, Ly - self.dy / 2,
                         -self.dz / 2 if self.dim == 3 else 0.0, (Lz - self.dz / 2) if self.dim == 3 else 0.0])

    @property
    def boxes(self):
        """Per-patch particle boxes widened by half a cell (core/patch/sync_particles_3d.c:402-411)."""
        b = np.zeros((self.npatch, 6))
        b[:, 0], b[:, 1] = self.x0 - self.dx / 2, self.x0 + (self.nx - 1) * self.dx + self.dx / 2
        b[:, 2], b[:, 3] = self.y0 - self.dy / 2, self.y0 + (self.ny - 1) * self.dy + self.dy / 2
        if self.dim == 3:
            b[:, 4], b[:, 5] = self.z0 - self.dz / 2, self.z0 + (self.nz - 1) * self.dz + self.dz / 2
        return b


def block_rank_map(npx, npy, npz, nranks):
    """Static contiguous block partition of the patch grid over ranks: split the longest patch axes first
    (8 ranks on a cubic grid -> 2x2x2 blocks).  Returns rank_of[global patch index]."""
    splits = [1, 1, 1]
    dims = [npx, npy, npz]
    r = nranks
    while r > 1:
        assert r % 2 == 0, "rank count must be a power of two"
        ok = [a for a in range(3) if dims[a] % (splits[a] * 2) == 0]
        assert ok, "patch grid not divisible by the rank grid"
        ax = max(ok, key=lambda a: (dims[a] / splits[a], a))  # longest divisible axis; z first on ties
        splits[ax] *= 2
        r //= 2
    bx, by, bz = npx // splits[0], npy // splits[1], npz // splits[2]
    ix, iy, iz = np.meshgrid(np.arange(npx), np.arange(npy), np.arange(npz), indexing="ij")
    rk = (ix // bx) + splits[0] * ((iy // by) + splits[1] * (iz // bz))
    rank_of = np.zeros(npx * npy * npz, dtype=np.int64)
    rank_of[(ix + npx * (iy + npy * iz)).ravel()] = rk.ravel()
    return rank_of


def make_patch_grid(dim, npx, npy, npz, nx, ny, nz, dx, dy, dz, n_guard=3, periodic=(True, True, True),
                    rank=0, nranks=1, rank_of=None) -> PatchGrid:
    if dim == 2:
        npz, nz = 1, 1
    ntot = npx * npy * npz
    if rank_of is None:
        rank_of = block_rank_map(npx, npy, npz, nranks) if nranks > 1 else np.zeros(ntot, dtype=np.int64)
    mine = np.nonzero(rank_of == rank)[0].astype(np.int64)
    local_of = -np.ones(ntot, dtype=np.int64)
    local_of[mine] = np.arange(len(mine))
    # local position of every patch on its own rank (needed to address remote patches)
    pos_on_rank = np.zeros(ntot, dtype=np.int64)
    for r in range(nranks):
        sel = np.nonzero(rank_of == r)[0]
        pos_on_rank[sel] = np.arange(len(sel))
    ix, iy, iz = mine % npx, (mine // npx) % npy, mine // (npx * npy)
    dirs = DIR3 if dim == 3 else DIR2
    nb = len(dirs)
    nidx = -np.ones((len(mine), nb), dtype=np.int64)
    n = (npx, npy, npz)
    for b, (sx, sy, sz) in enumerate(dirs):
        j = [ix + sx, iy + sy, iz + sz]
        ok = np.ones(len(mine), dtype=bool)
        for a in range(3):
            if periodic[a] or (a == 2 and dim == 2):
                j[a] = j[a] % n[a]
            else:
                ok &= (j[a] >= 0) & (j[a] < n[a])
                j[a] = np.clip(j[a], 0, n[a] - 1)
        g = j[0] + npx * (j[1] + npy * j[2])
        nidx[:, b] = np.where(ok, g, -1)
    safe = np.where(nidx >= 0, nidx, 0)
    nrank = np.where(nidx >= 0, rank_of[safe], -1)
    nip = np.where((nidx >= 0) & (nrank == rank), local_of[safe], -1)
    remote_rank = np.where((nidx >= 0) & (nrank != rank), nrank, -1)
    pg = PatchGrid(dim, npx, npy, npz, nx, ny, nz, dx, dy, dz if dim == 3 else 0.0, n_guard, tuple(periodic),
                   index=mine, x0=ix * nx * dx, y0=iy * ny * dy, z0=iz * nz * (dz if dim == 3 else 0.0),
                   neighbor_index=nidx, neighbor_ipatch=nip, neighbor_rank=remote_rank, rank=rank, nranks=nranks)
    pg.remote_ipatch = np.where(remote_rank >= 0, pos_on_rank[safe], -1)
    pg.rank_of = rank_of
    return pg


@dataclass
class ThermalPlasma:
    """Uniform thermal electron-proton plasma: n = n_c(0.8um), d = lambda/20, T = 1 keV."""
    dim: int = 3
    cells: tuple = (256, 256, 256)     # global cells
    patch: tuple = (16, 16, 16)        # cells per patch
    ppc: tuple = (16, 16)              # electrons, protons per cell
    d: float = 0.8e-6 / 20
    density: float = 1.742e27
    temperature_eV: float = 1.0e3
    dt_cfl: float = 0.95
    n_guard: int = 3
    seed: int = 1234

    @property
    def npatches(self):
        return tuple(c // p for c, p in zip(self.cells, self.patch))

    @property
    def dt(self):  # simulation/simulation.py:219,1288
        return self.dt_cfl * (self.dim / self.d**2) ** -0.5 / C_LIGHT

    @property
    def q(self):
        return [-E_CHARGE, E_CHARGE]

    @property
    def m(self):
        return [M_E, M_P]

    @property
    def uth(self):  # per-component thermal momentum sqrt(kT/mc^2)
        return [float(np.sqrt(self.temperature_eV * E_CHARGE / (m * C_LIGHT**2))) for m in self.m]

    @property
    def weights(self):
        dV = self.d ** self.dim
        return [self.density * dV / p for p in self.ppc]

    def grid(self, rank=0, nranks=1) -> PatchGrid:
        np_ = self.npatches
        return make_patch_grid(self.dim, np_[0], np_[1], np_[2] if self.dim == 3 else 1, self.patch[0], self.patch[1],
                               self.patch[2] if self.dim == 3 else 1, self.d, self.d, self.d, self.n_guard,
                               (True, True, True), rank, nranks)

    def n_particles(self):
        return int(np.prod(self.cells[:self.dim])) * sum(self.ppc)


def build_engine(wl: ThermalPlasma, device=0, rank=0, nranks=1, slack=1.4, with_part=False):
    """Create a DeviceEngine for this rank's block of the workload and load the particles on the device."""
    from .engine import DeviceEngine
    pg = wl.grid(rank, nranks)
    eng = DeviceEngine(wl.dim, pg.npatch, pg.nx, pg.ny, pg.nz, pg.n_guard, pg.dx, pg.dy, pg.dz, len(wl.ppc), device)
    eng.set_geometry(pg.x0, pg.y0, pg.z0, pg.neighbor_ipatch, pg.boxes, pg.glob, rank, pg.index)
    cells = pg.nx * pg.ny * pg.nz
    Ly, Lz = pg.glob[3] - pg.glob[2], (pg.glob[5] - pg.glob[4]) if wl.dim == 3 else 1.0
    for s, ppc in enumerate(wl.ppc):
        eng.alloc_species(s, np.full(pg.npatch, cells * ppc, dtype=np.int64), slack=slack, min_extra=256, with_part=with_part)
        eng.init_uniform(s, ppc, wl.weights[s], wl.uth[s], wl.seed + 7919 * s)
        eng.configure_sort(s, pg.nx, 1, 1, pg.dx, Ly, Lz, pg.x0 - pg.dx / 2, pg.y0 - pg.dy / 2, pg.z0 - pg.dz / 2)
    eng.grid = pg
    eng.sync()
    return eng
