"""GPU parity through the PUBLIC API: the same user script (Simulation / Species / callbacks) that produced the
golden vectors with the unmodified reference is run on ``lambdapic_b200`` and compared step by step.

  * seed-for-seed: positions come from the host loader (bit-exact vs the reference), momenta and seed fields from
    the same ``init`` callback as oracle/make_golden.py;
  * per step: integer state bit-exact, floats <= 1e-12 (single step from t0) / 1e-11 (accumulated);
  * 1000 steps of BASELINE.json configs[0]: total-energy history within 1e-6 of the reference's.
"""
import types

import numpy as np
import pytest

from tests.parity import check_state_against_golden
from tests.test_host_api import make_sim

pytestmark = pytest.mark.gpu


def _golden_callbacks(first):
    from lambdapic_b200 import callback

    @callback("init")
    def set_momenta(sim):  # same statements as oracle/make_golden.py:set_momenta
        rng = np.random.default_rng(99)
        for p in sim.patches:
            for isp, part in enumerate(p.particles):
                n = part.npart
                sig = 0.6 if isp == 0 else 0.05
                part.ux[:] = rng.normal(0.2 if isp == 0 else -0.02, sig, n)
                part.uy[:] = rng.normal(0.0, sig, n)
                part.uz[:] = rng.normal(0.0, sig, n)
                part.inv_gamma[:] = 1.0 / np.sqrt(1 + part.ux**2 + part.uy**2 + part.uz**2)
            f = p.fields
            for a, amp in (("ex", 3e11), ("ey", -2e11), ("ez", 1e11), ("bx", 500.0), ("by", -800.0), ("bz", 300.0)):
                arr = getattr(f, a)
                arr[...] = amp * rng.standard_normal(arr.shape)
    return set_momenta


def _view(sim):
    srt = [types.SimpleNamespace(bucket_count=s.bucket_count_list, bound_min=s.bucket_bound_min_list,
                                 bound_max=s.bucket_bound_max_list, pidx=s.particle_index_list, nbuf_last=s.nbuf_last)
           for s in sim.sorter]
    return types.SimpleNamespace(patches=sim.patches, sorters=srt)


@pytest.mark.parametrize("dim,case", [(3, "golden3d"), (2, "golden2d")])
def test_public_api_run_matches_reference_step_by_step(dim, case, request):
    g = request.getfixturevalue(case)
    sim = make_sim(dim)
    seen = {}

    from lambdapic_b200 import callback

    @callback("start")
    def at_start(sim):
        if "t0" not in seen:
            seen["t0"] = check_state_against_golden(types.SimpleNamespace(patches=sim.patches, sorters=None), g, "t0", rtol=0.0,
                                                    check_sorter=False)
    init_cb = _golden_callbacks(True)
    for it in range(3):
        sim.run(nsteps=1, callbacks=[init_cb, at_start] if it == 0 else [at_start])
        for s in range(2):
            assert sim.sorter[s].reverse_x == bool(int(g[f"t1/reverse_x/{s}"]))
        worst = check_state_against_golden(_view(sim), g, f"t{it + 1}", rtol=1e-12 if it == 0 else 1e-11)
        assert worst <= 1e-11
    assert seen["t0"] == 0.0  # the state handed to the first step is bit-identical to the reference's
    assert sim.itime == 3
    sim.bridge.close()


def test_nonunified_path_with_pusher_stage_callback_2d(golden2d):
    """A callback at a pusher stage forces the non-fused operator sequence (simulation.py:896-911,993-1038); the
    result must equal the fused path to rounding, and the callback must see *_part on the host."""
    from lambdapic_b200 import callback
    g = golden2d
    sim = make_sim(2)
    seen = []

    @callback("_interpolator")
    def peek(sim):
        seen.append(float(np.abs(sim.patches[0].particles[sim.ispec].ex_part).max()))
    sim.run(nsteps=1, callbacks=[_golden_callbacks(True), peek])
    assert len(seen) == 2 and min(seen) > 0.0
    check_state_against_golden(_view(sim), g, "t1", rtol=1e-12)
    sim.bridge.close()


def test_energy_history_1000_steps_config0():
    """BASELINE.json configs[0] (2D periodic thermal e-/p+ plasma, 64x64, 4x4 patches, 32+32 ppc): energy history
    of 1000 steps within 1e-6 (relative to the total) of the unmodified reference's (tests/golden/energy_history_2d.npz)."""
    import os
    from lambdapic_b200 import Electron, Proton, Simulation, callback
    from oracle.make_golden import energy_setup
    ref = np.load(os.path.join(os.path.dirname(__file__), "golden", "energy_history_2d.npz"))
    d, n0 = 0.8e-6 / 20, 1.742e27
    sim = Simulation(nx=64, ny=64, dx=d, dy=d, npatch_x=4, npatch_y=4, dt_cfl=0.95,
                     boundary_conditions={k: "periodic" for k in ("xmin", "xmax", "ymin", "ymax")}, random_seed=77)
    sim.add_species([Electron(density=lambda x, y: n0, ppc=32), Proton(density=lambda x, y: n0, ppc=32)])
    set_momenta, energy, hist = energy_setup(sim, callback, np)
    energy.interval = 50  # host diagnostics every 50 steps; the state stays on the device in between
    sim.run(nsteps=int(ref["nsteps"]), callbacks=[set_momenta, energy])
    got = np.array(hist)
    want = ref["history"][0::50]  # stage `end` of step k runs with itime == k
    assert got.shape == want.shape
    total = want.sum(axis=1)
    err = np.abs(got - want).max(axis=1) / total
    assert err.max() <= 1e-6, err.max()
    drift = abs(got.sum(axis=1)[-1] - got.sum(axis=1)[0]) / got.sum(axis=1)[0]
    assert drift < 0.01  # the reference's own bar (tests/test_numerical_heating.py:103-133)
    sim.bridge.close()


def test_device_side_diagnostics_skip_mirror_sync(golden2d):
    """A callback declared needs_host=False reads sim.energies() (device reductions): no download/upload happens for it,
    and the numbers equal the host-side sums."""
    from lambdapic_b200 import callback
    sim = make_sim(2)
    seen = []

    @callback("end", needs_host=False)
    def diag(sim):
        seen.append(sim.energies())
    sim.initialize()
    before = dict(sim.bridge.stats)
    sim.run(nsteps=3, callbacks=[_golden_callbacks(True), diag])
    st = sim.bridge.stats
    assert st["downloads"] - before["downloads"] == 1 and st["uploads"] - before["uploads"] == 1, (before, st)  # run entry / exit only
    assert len(seen) == 3
    eps0 = 8.8541878188e-12
    e2 = sum(float((p.fields.ex[:p.fields.nx, :p.fields.ny]**2 + p.fields.ey[:p.fields.nx, :p.fields.ny]**2
                    + p.fields.ez[:p.fields.nx, :p.fields.ny]**2).sum()) for p in sim.patches)
    assert abs(seen[-1]["electric"] - 0.5 * eps0 * e2 * sim.dx * sim.dy) <= 1e-12 * seen[-1]["electric"]
    assert set(seen[-1]) == {"electric", "magnetic", "electron", "proton"}
    sim.bridge.close()


@pytest.mark.parametrize("dim,case", [(3, "golden_pml3d"), (2, "golden_pml2d")])
def test_public_api_with_cpml_boundaries_matches_reference(dim, case, request):
    """The script of oracle/make_golden.py:run_pml_case run on lambdapic_b200: CPML on every side, patches at the edge own
    up to three faces, particle boxes shrink by the layer thickness, leavers without neighbour are killed."""
    from lambdapic_b200 import Electron, Proton, Simulation, Simulation3D, callback
    from lambdapic_b200.pml import PSI_NAMES
    g = request.getfixturevalue(case)
    d, n0 = 0.8e-6 / 20, 1.742e27
    if dim == 3:
        sim = Simulation3D(nx=16, ny=16, nz=18, dx=d, dy=d * 1.25, dz=d * 0.8, npatch_x=2, npatch_y=2, npatch_z=2, dt_cfl=0.95,
                           boundary_conditions={k: "pml" for k in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")},
                           cpml_thickness=6, random_seed=777)
        dens = lambda x, y, z: n0  # noqa: E731
    else:
        sim = Simulation(nx=32, ny=24, dx=d, dy=d * 1.25, npatch_x=2, npatch_y=2, dt_cfl=0.95,
                         boundary_conditions={k: "pml" for k in ("xmin", "xmax", "ymin", "ymax")}, cpml_thickness=6, random_seed=778)
        dens = lambda x, y: n0  # noqa: E731
    sim.add_species([Electron(density=dens, ppc=2), Proton(density=dens, ppc=1)])

    @callback("init")
    def seed(sim):  # same statements as oracle/make_golden.py:run_pml_case.seed
        rng = np.random.default_rng(5)
        for p in sim.patches:
            for isp, part in enumerate(p.particles):
                n = part.npart
                sig = 0.4 if isp == 0 else 0.03
                part.ux[:] = rng.normal(0.1 if isp == 0 else -0.01, sig, n)
                part.uy[:] = rng.normal(0.0, sig, n)
                part.uz[:] = rng.normal(0.0, sig, n)
                part.inv_gamma[:] = 1.0 / np.sqrt(1 + part.ux**2 + part.uy**2 + part.uz**2)
            f = p.fields
            for a, amp in (("ex", 3e11), ("ey", -2e11), ("ez", 1e11), ("bx", 500.0), ("by", -800.0), ("bz", 300.0)):
                arr = getattr(f, a)
                arr[...] = amp * rng.standard_normal(arr.shape)
    seen = {}

    @callback("start")
    def at_start(sim):
        if "t0" not in seen:
            seen["t0"] = check_state_against_golden(types.SimpleNamespace(patches=sim.patches, sorters=None), g, "t0", rtol=0.0,
                                                    check_sorter=False)
    for it in range(3):
        sim.run(nsteps=1, callbacks=[seed, at_start] if it == 0 else [at_start])
        worst = check_state_against_golden(types.SimpleNamespace(patches=sim.patches, sorters=None), g, f"t{it + 1}",
                                           rtol=1e-12 if it == 0 else 1e-11, check_sorter=False)
        for ip, p in enumerate(sim.patches):
            assert ",".join(type(m).__name__ for m in p.pml_boundary) == str(g["meta/pml_faces"][ip])
            for ipml, m in enumerate(p.pml_boundary):
                for nm in PSI_NAMES[m.axis]:
                    ref = g[f"t{it + 1}/pml/{ip}/{ipml}/{nm}"]
                    assert np.abs(getattr(m, nm) - ref).max() <= 1e-11 * max(np.abs(ref).max(), 1e-300), (ip, nm)
    assert seen["t0"] == 0.0 and worst <= 1e-11
    sim.bridge.close()


@pytest.mark.parametrize("dim,case,nsteps", [(3, "golden_laser3d", 2), (2, "golden_laser2d", 3)])
def test_public_api_laser_into_cpml_box_matches_reference(dim, case, nsteps, request):
    """BASELINE.json configs[1]/[3] in miniature: CPML box, plasma, and a laser antenna at xmin built from the same
    SimpleLaser + (Laguerre-)Gaussian combination as the reference run; the laser callback runs device-side
    (needs_host=False), so no mirror traffic is caused by the per-step `_laser` stage."""
    from lambdapic_b200 import Electron, Proton, Simulation, Simulation3D, callback
    from tests.test_laser_sources import golden_laser
    g = request.getfixturevalue(case)
    d, n0 = 0.8e-6 / 20, 1.742e27
    if dim == 3:
        sim = Simulation3D(nx=16, ny=16, nz=18, dx=d, dy=d * 1.25, dz=d * 0.8, npatch_x=2, npatch_y=2, npatch_z=2, dt_cfl=0.95,
                           boundary_conditions={k: "pml" for k in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")},
                           cpml_thickness=6, random_seed=777)
        dens = lambda x, y, z: n0  # noqa: E731
    else:
        sim = Simulation(nx=32, ny=24, dx=d, dy=d * 1.25, npatch_x=2, npatch_y=2, dt_cfl=0.95,
                         boundary_conditions={k: "pml" for k in ("xmin", "xmax", "ymin", "ymax")}, cpml_thickness=6, random_seed=778)
        dens = lambda x, y: n0  # noqa: E731
    sim.add_species([Electron(density=dens, ppc=2), Proton(density=dens, ppc=1)])

    @callback("init")
    def seed(sim):
        rng = np.random.default_rng(5)
        for p in sim.patches:
            for isp, part in enumerate(p.particles):
                n = part.npart
                sig = 0.4 if isp == 0 else 0.03
                part.ux[:] = rng.normal(0.1 if isp == 0 else -0.01, sig, n)
                part.uy[:] = rng.normal(0.0, sig, n)
                part.uz[:] = rng.normal(0.0, sig, n)
                part.inv_gamma[:] = 1.0 / np.sqrt(1 + part.ux**2 + part.uy**2 + part.uz**2)
            f = p.fields
            for a, amp in (("ex", 3e11), ("ey", -2e11), ("ez", 1e11), ("bx", 500.0), ("by", -800.0), ("bz", 300.0)):
                arr = getattr(f, a)
                arr[...] = amp * rng.standard_normal(arr.shape)
    laser = golden_laser(dim)
    sim.initialize()
    before = dict(sim.bridge.stats)
    for it in range(nsteps):
        sim.run(nsteps=1, callbacks=[seed, laser] if it == 0 else [laser])
        worst = check_state_against_golden(types.SimpleNamespace(patches=sim.patches, sorters=None), g, f"t{it + 1}",
                                           rtol=1e-12 if it == 0 else 1e-11, check_sorter=False)
    assert worst <= 1e-11
    st = sim.bridge.stats
    assert st["uploads"] - before["uploads"] == nsteps and st["downloads"] - before["downloads"] == nsteps  # run entry/exit only
    sim.bridge.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dim,case", [(2, "golden_mw2d"), (3, "golden_mw3d")])
def test_public_api_moving_window_matches_reference(dim, case, request):
    """BASELINE.json configs[2] in miniature: x open with CPML until the window starts, window at c
    from t = 0, density varying along x.  The reference recycles a column of patches at steps 0, ~11, ~22 (2D) / 0, 15
    (3D); after each shift and at the end the whole state -- fields, slot-exact particles incl. fresh ids of the re-loaded
    patches, origins, neighbour tables, remaining CPML faces -- must match the unmodified reference's."""
    from lambdapic_b200 import Electron, MovingWindow, Proton, Simulation, Simulation3D, callback
    g = request.getfixturevalue(case)
    d, n0 = 0.8e-6 / 20, 1.742e27
    if dim == 3:
        sim = Simulation3D(nx=24, ny=8, nz=8, dx=d, dy=d * 1.25, dz=d * 0.8, npatch_x=3, npatch_y=1, npatch_z=1, dt_cfl=0.95,
                           boundary_conditions=dict(xmin="pml", xmax="pml", ymin="periodic", ymax="periodic",
                                                    zmin="periodic", zmax="periodic"), cpml_thickness=6, random_seed=4321)
        dens = lambda x, y, z: n0 * (1.0 + x * 2.0e5)  # noqa: E731
    else:
        sim = Simulation(nx=32, ny=16, dx=d, dy=d * 1.25, npatch_x=4, npatch_y=2, dt_cfl=0.95,
                         boundary_conditions=dict(xmin="pml", xmax="pml", ymin="periodic", ymax="periodic"),
                         cpml_thickness=6, random_seed=4322)
        dens = lambda x, y: n0 * (1.0 + x * 2.0e5)  # noqa: E731
    sim.add_species([Electron(density=dens, ppc=2), Proton(density=dens, ppc=1)])
    mw = MovingWindow(velocity=299792458.0, start_time=0.0)

    @callback("init")
    def seed(sim):
        rng = np.random.default_rng(6)
        for p in sim.patches:
            for isp, part in enumerate(p.particles):
                n = part.npart
                sig = 0.3 if isp == 0 else 0.02
                part.ux[:] = rng.normal(0.05 if isp == 0 else -0.01, sig, n)
                part.uy[:] = rng.normal(0.0, sig, n)
                part.uz[:] = rng.normal(0.0, sig, n)
                part.inv_gamma[:] = 1.0 / np.sqrt(1 + part.ux**2 + part.uy**2 + part.uz**2)
            f = p.fields
            for a, amp in (("ex", 3e11), ("ey", -2e11), ("ez", 1e11), ("bx", 500.0), ("by", -800.0), ("bz", 300.0)):
                arr = getattr(f, a)
                arr[...] = amp * rng.standard_normal(arr.shape)
    sim.initialize()
    nsteps = int(g["meta/nsteps"])
    dump_after = set(int(v) for v in g["meta/dump_after"])
    shift_steps = [int(v) for v in g["meta/shift_steps"]]
    assert len(shift_steps) >= 2
    worst, seen_shifts = 0.0, []
    for it in range(nsteps):
        before = [p.x0 for p in sim.patches]
        sim.run(nsteps=1, callbacks=[seed, mw] if it == 0 else [mw])
        if before != [p.x0 for p in sim.patches]:
            seen_shifts.append(it)
        if it + 1 in dump_after:
            tag = f"t{it + 1}"
            assert np.array_equal(np.array([p.x0 for p in sim.patches]), g[f"{tag}/x0"]), "patch origins"
            assert np.array_equal(np.array([p.ipatch_x for p in sim.patches]), g[f"{tag}/ipatch_x"])
            assert np.array_equal(np.array([p.neighbor_ipatch for p in sim.patches]), g[f"{tag}/neighbor_ipatch"])
            assert [",".join(type(m).__name__ for m in p.pml_boundary) for p in sim.patches] == list(g[f"{tag}/pml_faces"])
            assert np.allclose([mw.total_shift, mw.patch_this_shift, mw.num_shifts], g[f"{tag}/mw"], rtol=1e-14, atol=0)
            rtol = 1e-12 if it == 0 else 1e-10  # multi-step drift of the summation order; single steps are <= 1e-12
            worst = max(worst, check_state_against_golden(types.SimpleNamespace(patches=sim.patches, sorters=None), g, tag,
                                                          rtol=rtol, check_sorter=False))
    assert seen_shifts == shift_steps
    assert worst <= 1e-10
    sim.bridge.close()


@pytest.mark.gpu
def test_callback_read_write_hints_move_only_the_named_arrays():
    """Mirror elision: a diagnostic that declares reads=("ex", "rho"), writes=() sees exactly what a
    fully synced callback sees, downloads two field arrays per trigger and uploads nothing; a callback that declares
    writes=("bz",) has its change picked up by the device."""
    from lambdapic_b200 import Electron, Proton, Simulation, callback
    d, n0 = 0.8e-6 / 20, 1.742e27

    def build():
        sim = Simulation(nx=32, ny=32, dx=d, dy=d, npatch_x=2, npatch_y=2, dt_cfl=0.95, random_seed=5,
                         boundary_conditions={k: "periodic" for k in ("xmin", "xmax", "ymin", "ymax")})
        sim.add_species([Electron(density=lambda x, y: n0, ppc=4), Proton(density=lambda x, y: n0, ppc=4)])

        @callback("init")
        def heat(sim):
            rng = np.random.default_rng(3)
            for p in sim.patches:
                for part in p.particles:
                    part.ux[:] = rng.normal(0.0, 0.05, part.npart)
                    part.inv_gamma[:] = 1.0 / np.sqrt(1 + part.ux**2 + part.uy**2 + part.uz**2)
        return sim, heat
    seen = {"full": [], "hint": []}

    def probe(tag):
        def f(sim):
            seen[tag].append([np.array(p.fields.ex).copy() for p in sim.patches] + [np.array(p.fields.rho).copy() for p in sim.patches])
        return f
    sim_a, heat_a = build()
    sim_a.run(nsteps=4, callbacks=[heat_a, callback("end", interval=2)(probe("full"))])
    sim_b, heat_b = build()
    sim_b.initialize()
    before = dict(sim_b.bridge.stats)
    field_bytes = sim_b.bridge.engine.fields_host[0].nbytes

    @callback("maxwell_1", interval=3, reads=("bz",), writes=("bz",))
    def kick(sim):
        for p in sim.patches:
            p.fields.bz[...] += 1.0e6
    energies = []

    @callback("end", needs_host=False)
    def device_diag(sim):  # shares the stage with the hinted probe: must keep working on the DEVICE state (no upload)
        energies.append(sim.energies())
    sim_b.run(nsteps=4, callbacks=[heat_b, callback("end", interval=2, reads=("ex", "rho"), writes=())(probe("hint")), device_diag])
    assert len(energies) == 4 and all(e["electric"] >= 0.0 for e in energies)
    st = sim_b.bridge.stats
    full = sim_b.bridge.state_bytes()
    # run() entry/exit move everything once each; the two triggers of the diagnostic add 2 x 2 field arrays down, nothing up
    assert st["d2h_bytes"] - before["d2h_bytes"] == full + 2 * 2 * field_bytes
    assert st["h2d_bytes"] - before["h2d_bytes"] == full
    assert len(seen["full"]) == len(seen["hint"]) == 2
    for a, b in zip(seen["full"], seen["hint"]):
        for x, y in zip(a, b):  # two separate runs: equal up to the order of the fp64 atomics
            assert np.allclose(x, y, rtol=0.0, atol=1e-10 * max(float(np.abs(x).max()), 1e-300))
    sim_b.run(nsteps=1, callbacks=[kick])  # itime == 4: not triggered (interval 3)
    assert abs(np.mean([np.mean(p.fields.bz) for p in sim_b.patches])) < 1.0e5
    sim_b.run(nsteps=2, callbacks=[kick])  # itime 5, 6: triggered at 6 -> the uniform offset written on the host reaches the device
    assert np.mean([np.mean(p.fields.bz) for p in sim_b.patches]) > 5.0e5
    sim_a.bridge.close(); sim_b.bridge.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [2, 3])
def test_extract_species_density_conserves_the_species_charge(dim):
    """ExtractSpeciesDensity (callback/utils.py:240-400): the density of every species, taken as the rho difference around
    its deposit after a device-side guard reduce, integrates to the species' total weight (Esirkepov/TSC deposits are
    charge conserving) and only the rho array crosses PCIe."""
    from lambdapic_b200 import Electron, ExtractSpeciesDensity, Proton, Simulation, Simulation3D
    d, n0 = 0.8e-6 / 20, 1.742e27
    if dim == 3:
        sim = Simulation3D(nx=16, ny=16, nz=16, dx=d, dy=d, dz=d, npatch_x=2, npatch_y=2, npatch_z=2, random_seed=11,
                           boundary_conditions={k: "periodic" for k in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")})
        dens = lambda x, y, z: n0 * (1.0 + 0.5 * np.sin(x / (16 * d) * 2 * np.pi))  # noqa: E731
    else:
        sim = Simulation(nx=32, ny=32, dx=d, dy=d, npatch_x=2, npatch_y=2, random_seed=11,
                         boundary_conditions={k: "periodic" for k in ("xmin", "xmax", "ymin", "ymax")})
        dens = lambda x, y: n0 * (1.0 + 0.5 * np.sin(x / (32 * d) * 2 * np.pi))  # noqa: E731
    ele, pro = Electron(density=dens, ppc=4), Proton(density=dens, ppc=2)
    sim.add_species([ele, pro])
    sim.initialize()
    ne, npr = ExtractSpeciesDensity(sim, ele, interval=2), ExtractSpeciesDensity(sim, pro, interval=2)
    before = dict(sim.bridge.stats)
    sim.run(nsteps=3, callbacks=[ne, npr])
    dV = d ** dim
    for diag, isp in ((ne, 0), (npr, 1)):
        total_w = sum(float(p.particles[isp].w[~np.asarray(p.particles[isp].is_dead)].sum()) for p in sim.patches)
        assert diag.density.shape == ((16, 16, 16) if dim == 3 else (32, 32))
        assert abs(diag.density.sum() * dV - total_w) <= 1e-10 * total_w
        assert diag.density.min() > 0.2 * n0 and diag.density.max() < 2.5 * n0
    st = sim.bridge.stats
    rho_bytes = sim.bridge.engine.fields_host[0].nbytes
    # steps 0 and 2 trigger; per trigger rho is fetched after species 0 (density of e-, previous rho of p+) and after species 1
    assert st["d2h_bytes"] - before["d2h_bytes"] == sim.bridge.state_bytes() + 2 * 3 * rho_bytes
    sim.bridge.close()


@pytest.mark.gpu
def test_get_fields_inside_run_moves_only_the_slice():
    """lambdapic_b200.get_fields (callback/utils.py:26-230) called from a device-side callback: in 3D one interior z-plane
    per patch crosses PCIe (lpic_download_field_slice), and it equals the slice of the full mirror downloaded at run() exit."""
    import lambdapic_b200 as lp
    d = 0.8e-6 / 20
    sim = lp.Simulation3D(nx=32, ny=16, nz=32, dx=d, dy=d, dz=d, npatch_x=2, npatch_y=1, npatch_z=2, dt_cfl=0.95,
                          boundary_conditions={k: "periodic" for k in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")}, random_seed=5)
    species = [lp.Electron(density=lambda x, y, z: 1.7e27, ppc=4), lp.Proton(density=lambda x, y, z: 1.7e27, ppc=2)]
    sim.add_species(species)
    sim.initialize()
    z_at = sim.Lz * 0.6
    seen = {}

    @lp.callback("end", needs_host=False)
    def grab(sim):
        before = sim.bridge.stats["d2h_bytes"]
        seen["fields"] = lp.get_fields(sim, ["ex", "jz", "rho"], z_at)
        seen["bytes"] = sim.bridge.stats["d2h_bytes"] - before
    sim.run(nsteps=3, callbacks=[lp.SetTemperature(species[0], 2.0e4), lp.SetTemperature(species[1], 2.0e4), grab])
    plane_bytes = 3 * sim.nx * sim.ny * 8
    assert 0 < seen["bytes"] <= plane_bytes, (seen["bytes"], plane_bytes)
    after = lp.get_fields(sim, ["ex", "jz", "rho"], z_at)   # host mirrors are current again after run()
    for a, b in zip(seen["fields"], after):
        assert a.shape == (sim.nx, sim.ny) and np.array_equal(a, b)
    assert np.abs(after[1]).max() > 0 and np.abs(after[2]).max() > 0
    sim.bridge.close()
