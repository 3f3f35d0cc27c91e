"""Host-side helpers of the reference's ``callback/utils.py`` that sit directly on the hot path's data: ``get_fields`` (the
assembled global field, or its z-slice in 3D) and ``SetTemperature`` (Maxwell-Juettner momenta at stage ``init``, used by
BASELINE.json configs[0] and configs[4]).  Same names, arguments, stages and random-number consumption as the reference,
so a script written against ``lambdapic.callback.utils`` runs unchanged and, for ``SetTemperature``, produces the same
momenta bit for bit from the same seed (tests/test_reference_callbacks.py runs both side by side).

Device-aware where it pays: inside ``Simulation.run`` (device authoritative) ``get_fields`` moves only what it returns --
one attribute of every patch in 2D, ONE interior z-plane per patch in 3D (``lpic_download_field_slice``) -- instead of
relying on a full mirror download.  Declare the calling callback ``@callback(stage, needs_host=False)`` to use that path.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Sequence

import numpy as np

from ._lib import FIELD_ATTRS, check
from .callback import Callback

E_CHARGE = 1.602176634e-19
C_LIGHT = 299792458.0


# ---------------------------------------------------------------------------------------------------------------------
# get_fields (reference: callback/utils.py:26-230)
# ---------------------------------------------------------------------------------------------------------------------
def _mask(fields):
    m = 0
    for f in fields:
        if f not in FIELD_ATTRS:
            raise ValueError(f"unknown field {f!r}")
        m |= 1 << FIELD_ATTRS.index(f)
    return m


def get_fields(sim, fields: Sequence[str], slice_at: Optional[float] = None):
    """Global (nx, ny) arrays of the named fields, guards stripped; in 3D the z-plane at `slice_at` (default Lz/2).
    Rank 0 returns the arrays, other ranks a list of None (as the reference)."""
    if sim.dimension == 3:
        return get_fields_3d(sim, fields, slice_at)
    if sim.dimension == 2:
        return get_fields_2d(sim, fields)
    raise ValueError(f"Unsupported simulation type: {type(sim)}")


def _assemble(sim, planes_by_index, where):
    """planes_by_index: {global patch index: (nx_per_patch, ny_per_patch) array}; where: {(ipx, ipy): index}."""
    out = np.zeros((sim.nx, sim.ny))
    nxp, nyp = sim.nx_per_patch, sim.ny_per_patch
    for (ipx, ipy), index in where.items():
        out[ipx * nxp:(ipx + 1) * nxp, ipy * nyp:(ipy + 1) * nyp] = planes_by_index[index]
    return out


def get_fields_2d(sim, fields: Sequence[str]):
    assert sim.dimension == 2, "Only 2D simulation is supported"
    if not fields:
        return []
    br = getattr(sim, "bridge", None)
    if br is not None and br.resident:  # device authoritative: fetch only the requested attributes
        br.engine.download_fields(_mask(fields))
    ng = sim.n_guard
    comm, rank = sim.mpi.comm, sim.mpi.rank
    where = {(p.ipatch_x, p.ipatch_y): p.index for p in sim.patches}
    where_all = comm.gather(where)
    ret = []
    for f in fields:
        mine = {p.index: np.asarray(getattr(p.fields, f))[:-2 * ng, :-2 * ng].copy() for p in sim.patches}
        parts = comm.gather(mine)
        if rank == 0:
            planes = {k: v for d in parts for k, v in d.items()}
            ret.append(_assemble(sim, planes, {k: v for d in where_all for k, v in d.items()}))
        else:
            ret.append(None)
    return ret


def get_fields_3d(sim, fields: Sequence[str], slice_at: Optional[float] = None):
    assert sim.dimension == 3, "Only 3D simulation is supported"
    if not fields:
        return []
    if slice_at is None:
        slice_at = sim.Lz / 2
    if slice_at < 0 or slice_at > sim.Lz:
        raise ValueError(f"Slice position {slice_at} is outside the simulation domain [0, {sim.Lz}]")
    nzp, ng = sim.nz_per_patch, sim.n_guard
    iz_global = int((slice_at + sim.dz / 2) / sim.dz)
    iz_local = iz_global - iz_global // nzp * nzp
    holds = [p.zmin <= slice_at <= p.zmax for p in sim.patches]
    comm, rank = sim.mpi.comm, sim.mpi.rank
    where = {(p.ipatch_x, p.ipatch_y): p.index for p, h in zip(sim.patches, holds) if h}
    where_all = comm.gather(where)
    br = getattr(sim, "bridge", None)
    planes = None
    if br is not None and br.resident:  # one z-plane per patch crosses PCIe, nothing else
        eng = br.engine
        kz = np.array([iz_local if h else -1 for h in holds], dtype=np.int64)
        planes = np.zeros((len(fields), eng.npatch, sim.nx_per_patch, sim.ny_per_patch))
        mask = _mask(fields)
        order = [f for f in FIELD_ATTRS if f in fields]  # the library packs attributes in ascending id order
        check(eng.L.lpic_download_field_slice(eng.ctx, mask, C.c_void_p(kz.ctypes.data), C.c_void_p(planes.ctypes.data)))
        br.stats["d2h_bytes"] = br.stats.get("d2h_bytes", 0) + int(sum(holds)) * len(order) * sim.nx_per_patch * sim.ny_per_patch * 8
        planes = {f: planes[order.index(f)] for f in fields}
    ret = []
    for f in fields:
        if planes is not None:
            mine = {p.index: planes[f][ip] for ip, (p, h) in enumerate(zip(sim.patches, holds)) if h}
        else:
            mine = {p.index: np.asarray(getattr(p.fields, f))[:-2 * ng, :-2 * ng, iz_local].copy()
                    for p, h in zip(sim.patches, holds) if h}
        parts = comm.gather(mine)
        if rank == 0:
            ret.append(_assemble(sim, {k: v for d in parts for k, v in d.items()}, {k: v for d in where_all for k, v in d.items()}))
        else:
            ret.append(None)
    return ret


# ---------------------------------------------------------------------------------------------------------------------
# SetTemperature (reference: callback/utils.py:922-1049)
# ---------------------------------------------------------------------------------------------------------------------
class SetTemperature(Callback):
    """Stage ``init``: momenta of one species drawn from a Maxwell-Juettner distribution of `temperature` [eV] (a list of
    three stretches uy, uz relative to ux).  The generator is ``sim.rand_gen.spawn(1)[0]`` and the patches are visited in
    order, as the reference does, so a given seed gives the reference's momenta."""
    stage = "init"
    reads = ("particles",)
    writes = ("particles",)

    def __init__(self, species, temperature, interval: Callable | int | float | None = None, add: bool = False):
        self.species = species
        self.temperature = [temperature] * 3 if isinstance(temperature, (int, float)) else list(temperature)
        self.interval = (lambda sim: sim.itime == 0) if interval is None else interval
        self.add = add

    def _call(self, sim):
        ispec = self.species.ispec
        gen, = sim.rand_gen.spawn(1)
        theta = self.temperature[0] * E_CHARGE / (self.species.m * C_LIGHT**2)
        t0, t1, t2 = self.temperature  # anisotropy as the reference writes it: (u * T1) / T0, left to right
        for p in sim.patches:
            part = p.particles[ispec]
            alive = part.is_alive
            n = int(alive.sum())
            if n == 0:
                continue
            ux, uy, uz = self.sample_maxwell_juttner(n, theta, gen)
            if self.add:
                part.ux[alive] += ux
                part.uy[alive] += uy * t1 / t0
                part.uz[alive] += uz * t2 / t0
            else:
                part.ux[alive] = ux
                part.uy[alive] = uy * t1 / t0
                part.uz[alive] = uz * t2 / t0
            part.inv_gamma[alive] = 1 / np.sqrt(1 + part.ux[alive]**2 + part.uy[alive]**2 + part.uz[alive]**2)

    @staticmethod
    def maxwell_juttner_pdf(gamma, theta):
        from scipy.special import kn
        beta = np.sqrt(1 - 1 / (gamma**2))
        return (gamma**2 * beta) / (theta * kn(2, 1 / theta)) * np.exp(-gamma / theta)

    @staticmethod
    def sample_maxwell_juttner(size: int, theta: float, rand_gen=None):
        """gamma from three regimes (theta <= 0.01: Gamma(3/2) tail of the non-relativistic limit; <= 0.5: rejection from a
        uniform proposal under the numerically located maximum of the pdf; above: Gamma(3) proposal accepted with
        probability beta), then an isotropic direction.  Draw order = the reference's."""
        import scipy.optimize
        import scipy.stats
        rand_gen = rand_gen or np.random.default_rng()
        gamma = np.zeros(size)
        if theta <= 0.01:
            gamma[:] = scipy.stats.gamma(a=1.5, scale=theta).rvs(size=size, random_state=rand_gen) + 1
        elif theta <= 0.5:
            gmax = 1 + 10 * theta
            res = scipy.optimize.minimize_scalar(lambda g: -SetTemperature.maxwell_juttner_pdf(g, theta), bounds=(1, gmax), method="bounded")
            ceiling = -res.fun * 1.1 + 1e-10
            done = 0
            while done < size:
                prop = rand_gen.uniform(1, gmax, size - done)
                keep = prop[rand_gen.uniform(0, ceiling, size - done) < SetTemperature.maxwell_juttner_pdf(prop, theta)]
                gamma[done:done + len(keep)] = keep
                done += len(keep)
        else:
            dist = scipy.stats.gamma(a=3, scale=theta)
            done = 0
            while done < size:
                prop = dist.rvs(size - done, random_state=rand_gen)
                beta = np.sqrt(1 - 1 / (np.ma.array(prop, mask=prop < 1)**2))
                keep = prop[(rand_gen.uniform(size=size - done) < beta) & (prop >= 1)]
                gamma[done:done + len(keep)] = keep
                done += len(keep)
        u = np.sqrt(gamma**2 - 1)
        phi = rand_gen.uniform(0, 2 * np.pi, size)
        cost = rand_gen.uniform(-1, 1, size)
        sint = np.sqrt(1 - cost**2)
        return u * sint * np.cos(phi), u * sint * np.sin(phi), u * cost
