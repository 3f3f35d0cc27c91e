"""Parity at the geometry that is benchmarked: 2x2x2 grids of 16^3- and 32^3-cell patches at 16+16 particles per cell,
a hot plasma in which most particles change cell every step (general 125-point deposit), and a patch too large for the
tile kernel's histogram (falls back to the cell-ordered kernel with collapsed keys).  State generated on the fly, the
CUDA path and the CPU oracle start from identical bits and run the same steps.

Bars: integer state (is_dead, capacities, _id slot permutation) bit-exact; fields, currents, momenta, positions
<= 1e-12 of each array's max-abs (reference bars: tests/core/pusher/test_unified_pusher_3d.py:234-260,
tests/core/current/test_current_deposition.py:517-557 of the reference)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(wl, nsteps, field_amp=1.0, env=None, monkeypatch=None, rtol=1e-12):
    from oracle import oracle as orc
    from tests import gpu_harness as h
    if env:
        for k, v in env.items():
            monkeypatch.setenv(k, v)
    g = h.synthetic_snapshot(wl, field_amp=field_amp)
    eng, meta = h.engine_from_golden(g, "t0", with_part=True, slack=1.5, min_extra=256)
    ost = orc.OState.from_golden(g, "t0")
    del g
    rev = [False] * eng.nspec
    ost.set_reverse_x(rev)
    backend = "ref" if orc.have_ref() else "port"
    for _ in range(nsteps):
        eng.step(meta["dt"], meta["q"], meta["m"], rev, write_part=True)
        orc.step(ost, backend)
    worst = h.compare_with_oracle(eng, ost, rtol=rtol)
    launches = eng.L.lpic_launch_count()
    eng.close()
    return worst, launches


@pytest.mark.parametrize("patch", [16, 32])
def test_bench_geometry_two_steps_match_oracle(patch):
    from lambdapic_b200.workloads import ThermalPlasma
    wl = ThermalPlasma(dim=3, cells=(2 * patch,) * 3, patch=(patch,) * 3, ppc=(16, 16), temperature_eV=1.0e3)
    worst, _ = _run(wl, 2)
    assert max(worst.values()) <= 1e-12, worst


def test_hot_plasma_general_deposit_matches_oracle():
    """1 MeV: thermal momenta ~ 1.4 mc, most particles change cell during a step and take the general deposit; many leave
    their patch every step."""
    from lambdapic_b200.workloads import ThermalPlasma
    wl = ThermalPlasma(dim=3, cells=(32, 32, 32), patch=(16, 16, 16), ppc=(4, 4), temperature_eV=1.0e6)
    worst, _ = _run(wl, 2)
    assert max(worst.values()) <= 1e-12, worst


def test_patch_larger_than_the_histograms_takes_the_collapsed_key_path():
    """40x40x36 cells in ONE periodic patch: 300 tiles x 256 keys and (40+2)(40+2)(36+2) predicted keys both exceed the
    57344-entry shared-memory histogram, so the step runs k_push_sorted with keys collapsed along z."""
    from lambdapic_b200.workloads import ThermalPlasma
    wl = ThermalPlasma(dim=3, cells=(40, 40, 36), patch=(40, 40, 36), ppc=(2, 1), temperature_eV=2.0e4)
    worst, _ = _run(wl, 2)
    assert max(worst.values()) <= 1e-12, worst


@pytest.mark.parametrize("shape", [(5, 4, 6), (16, 16, 8), (7, 9, 19)])
def test_tile_kernel_on_odd_patch_shapes(shape):
    """Patch edges that are not multiples of the 4x4x16 tile: partially filled tiles, tiles whose halo ends at the guard."""
    from lambdapic_b200.workloads import ThermalPlasma
    wl = ThermalPlasma(dim=3, cells=tuple(2 * s for s in shape), patch=shape, ppc=(5, 3), temperature_eV=5.0e4)
    worst, _ = _run(wl, 3, rtol=2e-12)
    assert max(worst.values()) <= 2e-12, worst


def test_tile_and_sorted_kernels_agree(monkeypatch):
    """A/B of the round-2 tile kernel against the round-1 cell-ordered kernel on the same state."""
    from lambdapic_b200.workloads import ThermalPlasma
    wl = ThermalPlasma(dim=3, cells=(32, 32, 32), patch=(16, 16, 16), ppc=(6, 3), temperature_eV=2.0e4)
    a, _ = _run(wl, 2)
    b, _ = _run(wl, 2, env={"LPIC_PUSH_SORTED": "1"}, monkeypatch=monkeypatch)
    assert max(a.values()) <= 1e-12 and max(b.values()) <= 1e-12, (a, b)


@pytest.mark.parametrize("dim", [2, 3])
def test_record_layout_and_separate_arrays_agree(dim, monkeypatch):
    """The device keeps x y z w ux uy uz inv_gamma as 64-byte records (default) or as eight arrays (LPIC_PARTICLE_LAYOUT=soa,
    chosen when a species is allocated).  After three full steps from zero slack (sort, push, deposit, migration with growth
    and an arena re-layout) both must sit on the oracle, their integer state (capacities, is_dead, _id = the slot permutation)
    must be identical, and their floats must agree to the rounding noise of the deposit's atomic adds (whose order between
    CTAs is not fixed: 1e-13 of each array's max-abs)."""
    from oracle import oracle as orc
    from tests import gpu_harness as h
    from lambdapic_b200.workloads import ThermalPlasma
    if dim == 3:
        wl = ThermalPlasma(dim=3, cells=(32, 32, 32), patch=(16, 16, 16), ppc=(5, 3), temperature_eV=5.0e4)
    else:
        wl = ThermalPlasma(dim=2, cells=(64, 64), patch=(16, 32), ppc=(7, 5), temperature_eV=5.0e4)
    views = []
    for layout in ("rec", "soa"):
        if layout == "soa":
            monkeypatch.setenv("LPIC_PARTICLE_LAYOUT", "soa")
        else:
            monkeypatch.delenv("LPIC_PARTICLE_LAYOUT", raising=False)
        g = h.synthetic_snapshot(wl)
        eng, meta = h.engine_from_golden(g, "t0", with_part=True, slack=1.0, min_extra=0)  # zero slack: growth re-lays the arena out
        ost = orc.OState.from_golden(g, "t0")
        rev = [False] * eng.nspec
        ost.set_reverse_x(rev)
        for _ in range(3):
            eng.step(meta["dt"], meta["q"], meta["m"], rev, write_part=True)
            orc.step(ost, "ref" if orc.have_ref() else "port")
        worst = h.compare_with_oracle(eng, ost, rtol=1e-12)
        assert max(worst.values()) <= 1e-12, (layout, worst)
        st = h.host_view(eng, with_sorter=False)
        snap = {}
        for ip, patch in enumerate(st.patches):  # copies: the mirrors go away with the engine
            for name, arr in vars(patch.fields).items():
                snap[(ip, name)] = np.array(arr, copy=True)
            for isp, pt in enumerate(patch.particles):
                for name, arr in vars(pt).items():
                    if arr is not None:
                        snap[(ip, isp, name)] = np.array(arr, copy=True)
        views.append(snap)
        eng.close()
    a, b = views
    assert set(a) == set(b)
    for k in a:
        assert a[k].shape == b[k].shape, k
        if k[-1] in ("is_dead", "_id"):
            assert a[k].tobytes() == b[k].tobytes(), k
        else:
            alive = np.isfinite(a[k]) & np.isfinite(b[k])
            assert (np.isfinite(a[k]) == np.isfinite(b[k])).all(), k
            scale = max(float(np.abs(a[k][alive]).max()) if alive.any() else 0.0, 1e-300)
            assert float(np.abs(a[k][alive] - b[k][alive]).max() if alive.any() else 0.0) <= 1e-13 * scale, k
