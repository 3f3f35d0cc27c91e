"""The reference's OWN host-side callbacks run unchanged against lambdapic_b200's Simulation objects (drop-in claim of
SURVEY.md 8(b)), and the ports shipped in lambdapic_b200.utils reproduce them bit for bit.

Runs where /root/reference exists (the build container); the GPU box has no reference and skips.  The reference is
imported from the scratch copy that oracle/make_golden.py prepares (its import-time stubs for mpi4py / h5py / ...).
No device is involved: `SetTemperature` (stage `init`, callback/utils.py:922-1049) and `get_fields` (:26-230) work on the
host mirrors, which here are plain numpy arrays behind the same attribute names."""
import os
import sys
import types

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "lambdapic")), reason="needs /root/reference")


@pytest.fixture(scope="module")
def ref():
    """The reference package, importable (oracle/make_golden.py's scratch tree)."""
    from oracle import make_golden as mg
    mg.prepare_scratch()
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(mg.SCRATCH, "nbcache"))  # the reference's numba kernels compile once
    for sub in ("shims", "ref"):
        path = os.path.join(mg.SCRATCH, sub)
        if path not in sys.path:
            sys.path.insert(0, path)
    import lambdapic  # noqa: F401
    import importlib
    utils = importlib.import_module("lambdapic.callback.utils")  # (`lambdapic.callback` the attribute is the decorator)
    from lambdapic import Electron, Proton, Simulation, Simulation3D
    return types.SimpleNamespace(utils=utils, Electron=Electron, Proton=Proton, Simulation=Simulation, Simulation3D=Simulation3D)


class _HostEngine:
    """What lambdapic_b200.fields.Fields needs from an engine, backed by numpy (no device)."""

    def __init__(self, sim, npatch):
        three = sim.dimension == 3
        self.nx, self.ny, self.ng = sim.nx_per_patch, sim.ny_per_patch, sim.n_guard
        self.nz = sim.nz_per_patch if three else 1
        self.dx, self.dy, self.dz = sim.dx, sim.dy, (sim.dz if three else 0.0)
        self.shape = (self.nx + 2 * self.ng, self.ny + 2 * self.ng) + ((self.nz + 2 * self.ng,) if three else ())
        self.store = np.zeros((10, npatch) + self.shape)

    def field_view(self, attr, ip):
        from lambdapic_b200._lib import FIELD_ATTRS
        return self.store[FIELD_ATTRS.index(attr), ip]


def host_only(sim, seed):
    """Simulation.initialize() up to the point where the device bridge would take over: patches, species, particles loaded
    from the reference's generator stream, fields as numpy arrays."""
    from lambdapic_b200.comm import default_comm
    from lambdapic_b200.fields import Fields2D, Fields3D
    from lambdapic_b200.operators import SingleRankMPI
    comm = default_comm()
    sim.grid = sim._grid(0, 1)
    sim.patches = sim.create_patches(sim.grid)
    sim.patches._comm = comm
    sim._set_global_domain_bounds()
    sim.mpi = SingleRankMPI(comm)
    eng = _HostEngine(sim, sim.patches.npatches)
    for ip, p in enumerate(sim.patches):
        F = Fields3D if sim.dimension == 3 else Fields2D
        p.fields = F(eng, ip, sim.dimension, p.x0, p.y0, *((p.z0,) if sim.dimension == 3 else ()))
    for s in sim.species:
        sim.patches.add_species(s, aux_attrs=s._aux_attrs)
    sim.rand_gen = np.random.default_rng(seed).spawn(1)[0]
    sim.patches.fill_particles(sim.rand_gen)
    return sim


def _pair(ref, dim, seed=77):
    import lambdapic_b200 as lp
    d = 0.8e-6 / 20
    kw = dict(nx=16, ny=8, dx=d, dy=d * 1.1, npatch_x=2, npatch_y=2, dt_cfl=0.95, random_seed=seed)
    bc = {k: "periodic" for k in ("xmin", "xmax", "ymin", "ymax")}
    if dim == 3:
        kw.update(nz=8, dz=d * 0.9, npatch_z=2)
        bc.update(zmin="periodic", zmax="periodic")
    ours = (lp.Simulation3D if dim == 3 else lp.Simulation)(boundary_conditions=bc, **kw)
    theirs = (ref.Simulation3D if dim == 3 else ref.Simulation)(boundary_conditions=bc, **kw)
    dens = (lambda x, y, z: 1.7e27) if dim == 3 else (lambda x, y: 1.7e27)  # (the reference validates density as a callable)
    sp_o = [lp.Electron(density=dens, ppc=3), lp.Proton(density=dens, ppc=2)]
    sp_t = [ref.Electron(density=dens, ppc=3), ref.Proton(density=dens, ppc=2)]
    ours.add_species(sp_o)
    theirs.add_species(sp_t)
    theirs.initialize()
    host_only(ours, seed)
    return ours, theirs, sp_o, sp_t


@pytest.mark.parametrize("dim,temperature", [(2, 1.0e3), (3, 1.0e3), (3, 5.0e4), (2, [1.0e6, 2.0e6, 0.5e6])])
def test_reference_set_temperature_runs_on_our_simulation_and_the_port_matches(ref, dim, temperature):
    """Three ways to the same momenta: the reference's callback on the reference's Simulation, the SAME callback object
    class on our Simulation, and our port on our Simulation -- all bit-identical (same seed, same draw order).  The three
    temperatures cover the sampler's three regimes for electrons (theta = 0.002, 0.098, 1.96)."""
    import lambdapic_b200 as lp
    from lambdapic_b200.utils import SetTemperature
    ours, theirs, sp_o, sp_t = _pair(ref, dim)
    port_sim, _, sp_p, _ = _pair(ref, dim)
    for isp in range(2):
        ref.utils.SetTemperature(sp_t[isp], temperature)(theirs)       # reference callback, reference simulation
        ref.utils.SetTemperature(sp_o[isp], temperature)(ours)         # reference callback, OUR simulation objects
        SetTemperature(sp_p[isp], temperature)(port_sim)               # our port, our simulation objects
    assert len(ours.patches) == len(theirs.patches)
    for po, pt, pp in zip(ours.patches, theirs.patches, port_sim.patches):
        for isp in range(2):
            a, b, c = po.particles[isp], pt.particles[isp], pp.particles[isp]
            assert a.npart == b.npart == c.npart
            for attr in ("x", "y", "w", "ux", "uy", "uz", "inv_gamma"):
                assert np.array_equal(getattr(a, attr), getattr(b, attr), equal_nan=True), (attr, "reference callback on ours vs on theirs")
                assert np.array_equal(getattr(c, attr), getattr(b, attr), equal_nan=True), (attr, "port vs reference")
    assert isinstance(SetTemperature(sp_p[0], 1.0), lp.Callback) and SetTemperature.stage == "init"
    # sanity of the physics: <ux^2> ~ theta for the non-relativistic case
    if temperature == 1.0e3:
        u = np.concatenate([p.particles[0].ux for p in ours.patches])
        theta = 1.0e3 * 1.602176634e-19 / (sp_o[0].m * 299792458.0**2)
        assert abs(np.mean(u**2) / theta - 1) < 0.1


@pytest.mark.parametrize("dim", [2, 3])
def test_reference_get_fields_reads_our_fields_and_the_port_matches(ref, dim):
    """The reference's get_fields_2d / get_fields_3d (duck-typed on the simulation) assemble the global array from OUR
    field objects; lambdapic_b200.utils.get_fields returns the same arrays."""
    from lambdapic_b200.utils import get_fields
    ours, _, _, _ = _pair(ref, dim)
    rng = np.random.default_rng(3)
    for p in ours.patches:
        for a in ("ex", "bz", "rho"):
            getattr(p.fields, a)[...] = rng.standard_normal(getattr(p.fields, a).shape)
    names = ["rho", "ex", "bz"]
    if dim == 2:
        want = ref.utils.get_fields_2d(ours, names)
        got = get_fields(ours, names)
    else:
        for z in (None, 0.0, ours.Lz * 0.73):
            want = ref.utils.get_fields_3d(ours, names, z)
            got = get_fields(ours, names, z)
            for w, g in zip(want, got):
                assert w.shape == (ours.nx, ours.ny) and np.array_equal(w, g)
        with pytest.raises(ValueError):
            get_fields(ours, names, ours.Lz * 1.5)
        return
    for w, g in zip(want, got):
        assert w.shape == (ours.nx, ours.ny) and np.array_equal(w, g)
