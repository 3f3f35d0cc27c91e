"""CPU tests that PIN the oracle: both oracle back-ends must reproduce the golden vectors written by the
unmodified reference (oracle/make_golden.py).  `ref` (the reference's own C extensions + our C FDTD) must be
bit-exact; `port` (our C restatement, built without FMA contraction) must agree to 1e-13."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.parity import check_state_against_golden


def _load(g, k):
    st = orc.OState.from_golden(g, f"t{k}")
    st.set_reverse_x([int(g[f"t1/reverse_x/{s}"]) for s in range(st.nspec)])
    return st


@pytest.mark.parametrize("case", ["golden3d", "golden2d"])
@pytest.mark.parametrize("k", [0, 1, 2])
def test_port_single_step_matches_reference(case, k, request):
    g = request.getfixturevalue(case)
    st = _load(g, k)
    orc.step(st, "port")
    worst = check_state_against_golden(st, g, f"t{k + 1}", rtol=1e-13)
    assert worst < 1e-13


@pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", ["golden3d", "golden2d"])
@pytest.mark.parametrize("k", [0, 1, 2])
def test_ref_extensions_single_step_bit_exact(case, k, request):
    g = request.getfixturevalue(case)
    st = _load(g, k)
    orc.step(st, "ref")
    worst = check_state_against_golden(st, g, f"t{k + 1}", rtol=0.0)
    assert worst == 0.0


@pytest.mark.parametrize("case", ["golden3d", "golden2d"])
def test_port_three_steps_from_t0(case, request):
    g = request.getfixturevalue(case)
    st = _load(g, 0)
    for _ in range(3):
        orc.step(st, "port")
    check_state_against_golden(st, g, "t3", rtol=1e-12)


def test_reverse_x_decision_matches_reference(golden3d):
    st = orc.OState.from_golden(golden3d, "t0")
    for s in range(st.nspec):
        assert orc.decide_reverse_x(st, s) == bool(int(golden3d[f"t1/reverse_x/{s}"]))


def test_charge_conservation_known_answer(golden3d):
    """Known answer of the reference's own tests (tests/core/current/test_current_deposition.py:517-557):
    sum(rho) = q * sum(w) / dV over alive particles."""
    st = _load(golden3d, 0)
    orc.reset_currents(st)
    for s in range(st.nspec):
        orc.push_deposit(st, s, "port")
    dV = st.dx * st.dy * st.dz
    total = sum(float(p.fields.rho.sum()) for p in st.patches)
    expect = sum(st.q[s] * float(p.particles[s].w[~p.particles[s].is_dead].sum()) for p in st.patches for s in range(st.nspec)) / dV
    scale = sum(abs(st.q[s]) * float(p.particles[s].w.sum()) for p in st.patches for s in range(st.nspec)) / dV
    assert abs(total - expect) <= 1e-10 * scale


def _psi_worst(st, g, tag):
    worst = 0.0
    for ip, p in enumerate(st.patches):
        for ipml, m in enumerate(p.pml):
            for nm in m.names["E"] + m.names["B"]:
                ref = g[f"{tag}/pml/{ip}/{ipml}/{nm}"]
                worst = max(worst, float(np.abs(getattr(m, nm) - ref).max()) / max(float(np.abs(ref).max()), 1e-300))
    return worst


@pytest.mark.parametrize("case", ["golden_pml3d", "golden_pml2d"])
@pytest.mark.parametrize("k", [0, 1, 2])
def test_cpml_port_matches_reference(case, k, request):
    """Open boundaries: kappa-scaled FDTD, psi currents, shrunk particle boxes, leavers without neighbour
    (core/boundary/cpml.py) against the unmodified reference's dumps; psi arrays bit-exact."""
    g = request.getfixturevalue(case)
    st = _load(g, k)
    orc.step(st, "port")
    assert check_state_against_golden(st, g, f"t{k + 1}", rtol=1e-13, check_sorter=False) < 1e-13
    assert _psi_worst(st, g, f"t{k + 1}") == 0.0


@pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", ["golden_pml3d", "golden_pml2d"])
def test_cpml_with_reference_extensions_bit_exact(case, request):
    g = request.getfixturevalue(case)
    st = _load(g, 0)
    for k in range(3):
        orc.step(st, "ref")
        assert check_state_against_golden(st, g, f"t{k + 1}", rtol=0.0, check_sorter=False) == 0.0
        assert _psi_worst(st, g, f"t{k + 1}") == 0.0


def _laser_sources(g, k, npatch):
    return {ip: (g[f"src/{k}/{ip}/ey"], g[f"src/{k}/{ip}/ez"]) for ip in range(npatch) if f"src/{k}/{ip}/ey" in g.files}


@pytest.mark.parametrize("case,nsteps", [("golden_laser3d", 2), ("golden_laser2d", 3)])
def test_laser_antenna_port_and_reference_extensions(case, nsteps, request):
    """Stage `_laser` (callback/laser.py:17-77) restated in C, fed with the reference's own source planes: bit-exact with
    the reference extensions as back-end, <= 1e-13 with the C port of the particle path."""
    g = request.getfixturevalue(case)
    lp = int(g["meta/cpml_thickness"]) + 2
    for backend, rtol in (("port", 1e-13), ("ref", 0.0)):
        if backend == "ref" and not orc.have_ref():
            continue
        for k in range(nsteps):
            st = _load(g, k)
            orc.step(st, backend, laser=(_laser_sources(g, k, len(st.patches)), lp))
            assert check_state_against_golden(st, g, f"t{k + 1}", rtol=rtol, check_sorter=False) <= rtol


@pytest.mark.parametrize("case", ["golden_mw2d", "golden_mw3d"])
def test_moving_window_port_and_reference_extensions(case, request):
    """MovingWindow (callback/utils.py:471-840) restated on the oracle state: patch recycling, neighbour tables, PMLX
    removal and the re-load of the recycled patches from the reference's generator stream.  Started from the reference's
    state at stage `init` of step 0, the whole run must reproduce every snapshot of the unmodified reference."""
    g = request.getfixturevalue(case)
    dim = int(g["meta/dim"])
    n0 = 1.742e27
    dens = (lambda x, y, z: n0 * (1.0 + x * 2.0e5)) if dim == 3 else (lambda x, y: n0 * (1.0 + x * 2.0e5))
    ppcs = [(lambda *a: 2.0), (lambda *a: 1.0)]
    npatch = [int(g[f"meta/npatch_{a}"]) for a in "xyz"[:dim]]
    nsteps = int(g["meta/nsteps"])
    dump_after = set(int(v) for v in g["meta/dump_after"])
    # port: 17-24 steps of accumulated summation-order differences in J (single steps are <= 1e-13); ref: bit-exact
    for backend, rtol in (("port", 1e-9), ("ref", 0.0)):
        if backend == "ref" and not orc.have_ref():
            continue
        st = orc.OState.from_golden(g, "t0")
        rand_gen = np.random.default_rng(int(g["meta/seed"])).spawn(1)[0]
        rand_gen.spawn(len(st.patches))  # the initial load took one child per patch
        mw = orc.OMovingWindow(st, 299792458.0, npatch, g["meta/ipatch"], g["meta/periodic"], float(g["meta/Lx"]),
                               [dens, dens], ppcs, rand_gen, start_time=0.0)
        for it in range(nsteps):
            mw.stage(st, it * st.dt)
            orc.step(st, backend)
            if it + 1 in dump_after:
                tag = f"t{it + 1}"
                assert np.array_equal(np.array([p.x0 for p in st.patches]), g[f"{tag}/x0"])
                assert np.array_equal(st.nbr, g[f"{tag}/neighbor_ipatch"])
                assert [",".join({"xmin": "PMLXmin", "xmax": "PMLXmax", "ymin": "PMLYmin", "ymax": "PMLYmax", "zmin": "PMLZmin",
                                  "zmax": "PMLZmax"}[m.face] for m in p.pml) for p in st.patches] == list(g[f"{tag}/pml_faces"])
                assert check_state_against_golden(st, g, tag, rtol=rtol, check_sorter=False) <= rtol
        assert mw.num_shifts == int(g[f"t{nsteps}/mw"][2])
