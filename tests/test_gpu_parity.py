"""GPU parity tests proper: the CUDA path (through the C-ABI) against the golden vectors written by the
unmodified reference and against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): integer state bit-exact (is_dead, _id slot permutation, capacities, sorter
tables, particle_index, nbuf, migration counts); floating point <= 1e-12 of each array's max-abs per step.
FDTD / guard copy / current reduce are additionally required to be bit-exact."""
import numpy as np
import pytest

from tests.parity import FIELD_ATTRS, check_state_against_golden, rel_err

pytestmark = pytest.mark.gpu


def _harness():
    from tests import gpu_harness
    return gpu_harness


def _reverse(g, nspec):
    return [bool(int(g[f"t1/reverse_x/{s}"])) for s in range(nspec)]


@pytest.mark.parametrize("case", ["golden3d", "golden2d"])
@pytest.mark.parametrize("k", [0, 1, 2])
def test_single_step_matches_reference_golden(case, k, request):
    g = request.getfixturevalue(case)
    h = _harness()
    eng, meta = h.engine_from_golden(g, f"t{k}")
    nbuf, mig = eng.step(meta["dt"], meta["q"], meta["m"], _reverse(g, eng.nspec), write_part=True)
    st = h.host_view(eng, nbuf)
    worst = check_state_against_golden(st, g, f"t{k + 1}", rtol=1e-12)
    assert worst <= 1e-12
    eng.close()


@pytest.mark.parametrize("case", ["golden3d", "golden2d"])
def test_three_steps_from_t0_with_relayout(case, request):
    """Tight physical capacity (slack 1.0) forces the arena re-layout path on the first migration."""
    g = request.getfixturevalue(case)
    h = _harness()
    eng, meta = h.engine_from_golden(g, "t0", with_part=False, slack=1.0, min_extra=0)
    moved = False
    for _ in range(3):
        nbuf, mig = eng.step(meta["dt"], meta["q"], meta["m"], _reverse(g, eng.nspec))
        moved = moved or any(r["moved"] for r in mig)
    assert moved, "expected at least one re-layout with zero slack"
    st = h.host_view(eng, nbuf)
    check_state_against_golden(st, g, "t3", rtol=1e-11, check_part_fields=False)
    eng.close()


@pytest.mark.parametrize("case", ["golden3d", "golden2d"])
def test_field_solver_and_guard_sync_bit_exact(case, request):
    """FDTD half steps + guard copies + current reduce against the CPU oracle: 0 ULP."""
    from oracle import oracle as orc
    g = request.getfixturevalue(case)
    h = _harness()
    eng, meta = h.engine_from_golden(g, "t1")
    ost = orc.OState.from_golden(g, "t1")
    rng = np.random.default_rng(5)
    # non-trivial J/rho including guards, identical on both sides
    for ip, p in enumerate(ost.patches):
        for a in ("jx", "jy", "jz", "rho"):
            v = rng.standard_normal(getattr(p.fields, a).shape) * 1e9
            getattr(p.fields, a)[...] = v
            eng.field_view(a, ip)[...] = v
    eng.upload_fields()
    dt = meta["dt"]
    from lambdapic_b200.engine import E_MASK, B_MASK
    eng.sync_currents(); orc.sync_currents(ost)
    eng.update_efield(0.5 * dt); orc.update_efield(ost, 0.5 * dt)
    eng.sync_guard_fields(E_MASK); orc.sync_guard_fields(ost, ("ex", "ey", "ez"))
    eng.update_bfield(0.5 * dt); orc.update_bfield(ost, 0.5 * dt)
    eng.sync_guard_fields(B_MASK); orc.sync_guard_fields(ost, ("bx", "by", "bz"))
    eng.download_fields()
    for ip, p in enumerate(ost.patches):
        for a in FIELD_ATTRS:
            assert np.array_equal(eng.field_view(a, ip), getattr(p.fields, a)), (a, ip)
    eng.close()


def test_migration_counts_match_oracle(golden3d):
    from oracle import oracle as orc
    h = _harness()
    g = golden3d
    eng, meta = h.engine_from_golden(g, "t1")
    ost = orc.OState.from_golden(g, "t1")
    ost.set_reverse_x(_reverse(g, ost.nspec))
    # push both so that there are leavers, then compare the migration bookkeeping only
    for s in range(eng.nspec):
        eng.push_deposit(s, meta["dt"], meta["q"][s], meta["m"][s])
        orc.push_deposit(ost, s, "port")
    ref = orc.sync_particles(ost, "port")
    for s in range(eng.nspec):
        rec = eng.sync_particles(s)
        for k in ("to_extend", "incoming", "outgoing", "alive"):
            assert np.array_equal(rec[k], ref[s][k]), (s, k)
    eng.close()


def test_charge_conservation_known_answer(golden3d):
    """tests/core/current/test_current_deposition.py:517-557 of the reference: sum(rho) = q sum(w)/dV."""
    h = _harness()
    g = golden3d
    eng, meta = h.engine_from_golden(g, "t0")
    eng.reset_currents()
    expect = scale = 0.0
    dV = eng.dx * eng.dy * eng.dz
    for s in range(eng.nspec):
        eng.push_deposit(s, meta["dt"], meta["q"][s], meta["m"][s])
        m = eng.species[s]
        for ip in range(eng.npatch):
            alive = ~m.view("is_dead", ip)
            expect += meta["q"][s] * float(m.view("w", ip)[alive].sum()) / dV
            scale += abs(meta["q"][s]) * float(m.view("w", ip)[alive].sum()) / dV
    eng.download_fields()
    total = float(sum(eng.field_view("rho", ip).sum() for ip in range(eng.npatch)))
    assert abs(total - expect) <= 1e-10 * scale
    eng.close()


def test_slot_order_kernel_matches_cell_ordered_kernel(golden3d):
    """The v1 kernel (one thread per slot in memory order) and the cell-ordered warp-cooperative kernel do the same
    arithmetic per particle; J/rho differ only by summation order."""
    h = _harness()
    g = golden3d
    e1, meta = h.engine_from_golden(g, "t1")
    e2, _ = h.engine_from_golden(g, "t1")
    e2.slot_order = True
    for e in (e1, e2):
        e.step(meta["dt"], meta["q"], meta["m"], _reverse(g, e.nspec), write_part=True)
    a, b = h.host_view(e1, with_sorter=False), h.host_view(e2, with_sorter=False)
    for ip in range(e1.npatch):
        for at in FIELD_ATTRS:
            assert rel_err(getattr(b.patches[ip].fields, at), getattr(a.patches[ip].fields, at)) <= 1e-13
        for s in range(e1.nspec):
            pa, pb = a.patches[ip].particles[s], b.patches[ip].particles[s]
            assert np.array_equal(pa.is_dead, pb.is_dead) and np.array_equal(pa._id.view(np.uint64), pb._id.view(np.uint64))
            alive = ~pa.is_dead
            for at in ("x", "y", "z", "ux", "uy", "uz", "inv_gamma", "ex_part", "bz_part"):
                assert rel_err(getattr(pb, at)[alive], getattr(pa, at)[alive]) <= 1e-13
    e1.close(); e2.close()


def test_nonfused_stages_equal_fused(golden3d):
    """interpolate -> Boris -> (positions) -> deposit path against the fused kernel (simulation.py:993-1038)."""
    h = _harness()
    g = golden3d
    e1, meta = h.engine_from_golden(g, "t1")
    e2, _ = h.engine_from_golden(g, "t1")
    dt = meta["dt"]
    for s in range(e1.nspec):
        q, m = meta["q"][s], meta["m"][s]
        e1.push_deposit(s, dt, q, m, write_part=True)
        e2.push_position(s, 0.5 * dt); e2.interpolate(s); e2.push_momentum(s, dt, q, m); e2.push_position(s, 0.5 * dt)
        e2.deposit(s, dt, q)
    a, b = h.host_view(e1, with_sorter=False), h.host_view(e2, with_sorter=False)
    for ip in range(e1.npatch):
        for at in ("jx", "jy", "jz", "rho"):
            assert rel_err(getattr(b.patches[ip].fields, at), getattr(a.patches[ip].fields, at)) <= 1e-12
        for s in range(e1.nspec):
            alive = ~a.patches[ip].particles[s].is_dead
            for at in ("x", "y", "z", "ux", "uy", "uz", "inv_gamma", "ex_part", "bz_part"):
                assert rel_err(getattr(b.patches[ip].particles[s], at)[alive], getattr(a.patches[ip].particles[s], at)[alive]) <= 1e-12
    e1.close(); e2.close()


@pytest.mark.parametrize("patch", [8, 16])
def test_particle_kernels_agree_on_thermal_plasma(patch, monkeypatch):
    """Slot-order kernel and the cell-ordered kernel in its unrolled and rolled (LPIC_PUSH_COMPACT) builds on a synthetic
    thermal plasma with random E/B: same particles out, J/rho equal up to summation order."""
    from lambdapic_b200.workloads import ThermalPlasma, build_engine
    h = _harness()
    wl = ThermalPlasma(dim=3, cells=(2 * patch, 2 * patch, 2 * patch), patch=(patch,) * 3, ppc=(6, 3), temperature_eV=2.0e4)
    rng = np.random.default_rng(3)
    views = []
    for mode in ("cell", "slot", "compact"):
        eng = build_engine(wl, with_part=True)
        eng.slot_order = mode == "slot"
        if mode == "compact":
            monkeypatch.setenv("LPIC_PUSH_COMPACT", "1")
        r = np.random.default_rng(11)
        for a, amp in (("ex", 3e11), ("ey", -2e11), ("ez", 1e11), ("bx", 500.0), ("by", -800.0), ("bz", 300.0)):
            eng.fields_host[FIELD_ATTRS.index(a)] = amp * r.standard_normal(eng.fields_host[0].shape)
        eng.upload_fields()
        for _ in range(2):
            eng.step(wl.dt, wl.q, wl.m, [False, False], write_part=True)
        views.append((eng, h.host_view(eng, with_sorter=False)))
    monkeypatch.delenv("LPIC_PUSH_COMPACT", raising=False)
    ref = views[0][1]
    for eng, v in views[1:]:
        for ip in range(eng.npatch):
            for at in FIELD_ATTRS:
                assert rel_err(getattr(v.patches[ip].fields, at), getattr(ref.patches[ip].fields, at)) <= 1e-12, at
            for s in range(eng.nspec):
                pa, pb = ref.patches[ip].particles[s], v.patches[ip].particles[s]
                assert np.array_equal(pa.is_dead, pb.is_dead) and np.array_equal(pa._id.view(np.uint64), pb._id.view(np.uint64))
                alive = ~pa.is_dead
                for at in ("x", "y", "z", "ux", "uy", "uz", "inv_gamma", "ex_part", "by_part"):
                    assert rel_err(getattr(pb, at)[alive], getattr(pa, at)[alive]) <= 1e-12, at
    del rng
    for eng, _ in views:
        eng.close()


@pytest.mark.parametrize("case", ["golden_pml3d", "golden_pml2d"])
@pytest.mark.parametrize("k", [0, 1, 2])
def test_cpml_single_step_matches_reference_golden(case, k, request):
    """Open (CPML) boundaries on the device: fields / currents / particles <= 1e-12, slot order bit-exact, psi arrays
    <= 1e-12 of their max-abs (the coefficients contain an exp() evaluated on the host)."""
    g = request.getfixturevalue(case)
    h = _harness()
    eng, meta = h.engine_from_pml_golden(g, f"t{k}")
    rev = [bool(int(g[f"t1/reverse_x/{s}"])) for s in range(eng.nspec)]
    eng.step(meta["dt"], meta["q"], meta["m"], rev, write_part=True)
    st = h.host_view(eng, with_sorter=False)
    assert check_state_against_golden(st, g, f"t{k + 1}", rtol=1e-12, check_sorter=False) <= 1e-12
    for e, (ip, slot, nms) in enumerate(eng.psi_names):
        for r, nm in enumerate(nms):
            ref = g[f"t{k + 1}/pml/{ip}/{slot}/{nm}"]
            assert rel_err(eng.psi_host[e, r], ref) <= 1e-12, (ip, slot, nm)
    eng.close()


@pytest.mark.parametrize("case,nsteps", [("golden_laser3d", 2), ("golden_laser2d", 3)])
def test_laser_antenna_single_step_matches_reference_golden(case, nsteps, request):
    """Stage `_laser` on the device, fed with the reference's source planes (the antenna kernel is bit-exact by
    construction; the rest of the step holds the usual 1e-12)."""
    g = request.getfixturevalue(case)
    h = _harness()
    lp = int(g["meta/cpml_thickness"]) + 2
    for k in range(nsteps):
        eng, meta = h.engine_from_pml_golden(g, f"t{k}")
        rev = [bool(int(g[f"t1/reverse_x/{s}"])) for s in range(eng.nspec)]
        patches = [ip for ip in range(eng.npatch) if f"src/{k}/{ip}/ey" in g.files]
        ranges = []
        for ip in patches:
            faces = str(g["meta/pml_faces"][ip])
            t = int(g["meta/cpml_thickness"])
            ranges.append([t if "PMLYmin" in faces else 0, eng.ny - t if "PMLYmax" in faces else eng.ny,
                           t if "PMLZmin" in faces else 0, eng.nz - t if "PMLZmax" in faces else eng.nz])
        ey = np.stack([g[f"src/{k}/{ip}/ey"] for ip in patches])
        ez = np.stack([g[f"src/{k}/{ip}/ez"] for ip in patches])
        eng.step(meta["dt"], meta["q"], meta["m"], rev, write_part=True, laser=(lp, patches, ranges, ey, ez))
        st = h.host_view(eng, with_sorter=False)
        assert check_state_against_golden(st, g, f"t{k + 1}", rtol=1e-12, check_sorter=False) <= 1e-12
        eng.close()
