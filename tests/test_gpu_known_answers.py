"""Known-answer and invariant tests in the style of the reference's own component tests:
tests/core/current/test_current_deposition.py (sums of J and rho), tests/core/interpolation/test_field_interpolation_*.py
(constant fields), tests/core/pusher/test_unified_pusher_*.py (dead particles untouched, dead+alive == alive-only),
tests/test_sort.py (bucket tables == histograms, sortedness, nbuf == 0 on re-sort, re-sort after extend).
They run on the CUDA path through the C-ABI; no golden vectors involved."""
import numpy as np
import pytest

from lambdapic_b200._lib import FIELD_ATTRS
from lambdapic_b200.workloads import C_LIGHT, ThermalPlasma, build_engine

pytestmark = pytest.mark.gpu


def _plasma(dim, **kw):
    cells = (16, 16, 16) if dim == 3 else (32, 32, 1)
    patch = (8, 8, 8) if dim == 3 else (16, 16, 1)
    return ThermalPlasma(dim=dim, cells=cells, patch=patch, ppc=(5, 3), **kw)


def _field_sum(eng, name):
    ng = eng.ng
    inner = tuple(slice(0, n + 2 * ng) for n in ((eng.nx, eng.ny, eng.nz)[:eng.dim]))  # guards are folded by sync_currents
    return float(sum(eng.field_view(name, ip)[inner].sum() for ip in range(eng.npatch)))


@pytest.mark.parametrize("dim", [2, 3])
def test_deposited_current_and_charge_sums(dim):
    """sum(J) dV = q sum(w v), sum(rho) dV = q sum(w) for random relativistic momenta (most particles change cell):
    test_current_deposition.py:328-369, 517-557."""
    wl = _plasma(dim)
    eng = build_engine(wl, with_part=True)
    rng = np.random.default_rng(1)
    dV = wl.d ** dim
    for s in range(eng.nspec):
        m = eng.species[s]
        eng.download_particles(s)
        eng.sync()
        for a in ("ux", "uy", "uz"):
            m.host[a][:] = rng.uniform(-10.0, 10.0, m.total)
        m.host["inv_gamma"][:] = 1.0 / np.sqrt(1 + m.host["ux"]**2 + m.host["uy"]**2 + m.host["uz"]**2)
        eng.upload_particles(s)
    eng.reset_currents()
    expect = np.zeros(4)
    for s in range(eng.nspec):
        m = eng.species[s]
        alive = m.host["is_dead"] == 0
        w, ig = m.host["w"][alive], m.host["inv_gamma"][alive]
        for c, a in enumerate(("ux", "uy", "uz")):
            expect[c] += wl.q[s] * float((w * m.host[a][alive] * ig).sum()) * C_LIGHT
        expect[3] += wl.q[s] * float(w.sum())
        eng.deposit(s, wl.dt, wl.q[s])  # standalone deposition operator (current_deposition_cpu_*)
    eng.sync_currents()
    eng.download_fields()
    got = np.array([_field_sum(eng, a) for a in ("jx", "jy", "jz", "rho")]) * dV
    expect_ok = [0, 1, 2, 3]
    scale = abs(wl.q[0]) * float(eng.species[0].host["w"].sum()) * C_LIGHT
    for c in expect_ok:
        ref_scale = scale if c < 3 else scale / C_LIGHT
        assert abs(got[c] - expect[c]) <= 1e-10 * ref_scale, (("jx", "jy", "jz", "rho")[c], got[c], expect[c])
    eng.close()


@pytest.mark.parametrize("dim", [2, 3])
def test_constant_fields_interpolate_to_the_constant(dim):
    """A field that is constant on its staggered grid is seen as that constant by every particle
    (test_field_interpolation_3d.py:259-291); dead slots keep their old *_part values."""
    wl = _plasma(dim)
    eng = build_engine(wl, with_part=True)
    consts = dict(ex=1.5, ey=-2.5, ez=3.5, bx=-4.5, by=5.5, bz=-6.5)
    for a, v in consts.items():
        eng.fields_host[FIELD_ATTRS.index(a)][...] = v
    eng.upload_fields()
    m = eng.species[0]
    eng.download_particles(0)
    eng.sync()
    m.host["is_dead"][::7] = 1
    for a in consts:
        m.host[a + "_part"][:] = 99.0
    eng.upload_particles(0)
    eng.interpolate(0)
    eng.download_particles(0)
    alive = m.host["is_dead"] == 0
    for a, v in consts.items():
        got = m.host[a + "_part"]
        assert np.allclose(got[alive], v, rtol=1e-13, atol=0)
        assert np.all(got[~alive] == 99.0)
    eng.close()


@pytest.mark.parametrize("dim", [2, 3])
def test_dead_particles_are_untouched_and_do_not_deposit(dim):
    """test_unified_pusher_3d.py:226-260: a dead slot keeps every attribute; fields from (dead + alive) equal fields
    from the alive particles alone."""
    wl = _plasma(dim, temperature_eV=5.0e4)
    results = []
    for kill in (False, True):
        eng = build_engine(wl, with_part=True)
        r = np.random.default_rng(7)
        for a, amp in (("ex", 3e11), ("ey", -2e11), ("ez", 1e11), ("bx", 500.0), ("by", -800.0), ("bz", 300.0)):
            eng.fields_host[FIELD_ATTRS.index(a)][...] = amp * r.standard_normal(eng.fields_host[0].shape)
        eng.upload_fields()
        snap = {}
        for s in range(eng.nspec):
            m = eng.species[s]
            eng.download_particles(s)
            eng.sync()
            dead = np.zeros(m.total, dtype=bool)
            dead[3::5] = True
            dead &= m.host["is_dead"] == 0
            if kill:   # variant B: the slots are dead
                m.host["is_dead"][dead] = 1
            else:      # variant A: the same particles are removed by giving them zero weight far from any effect on J
                m.host["w"][dead] = 0.0
            snap[s] = (dead, {a: m.host[a][dead].copy() for a in ("x", "y", "ux", "uy", "uz", "inv_gamma")})
            eng.upload_particles(s)
        eng.reset_currents()
        for s in range(eng.nspec):
            eng.push_deposit(s, wl.dt, wl.q[s], wl.m[s], True)
        eng.sync_currents()
        eng.download_all()
        if kill:
            for s in range(eng.nspec):
                dead, before = snap[s]
                for a, v in before.items():
                    assert np.array_equal(eng.species[s].host[a][dead], v), a
        results.append([eng.fields_host[FIELD_ATTRS.index(a)].copy() for a in ("jx", "jy", "jz", "rho")])
        eng.close()
    for a, b in zip(*results):
        assert np.abs(a - b).max() <= 1e-10 * max(float(np.abs(a).max()), 1e-300)


@pytest.mark.parametrize("dim", [2, 3])
def test_sort_tables_sortedness_and_idempotence(dim):
    """tests/test_sort.py:38-72,143-200: bucket_count == histogram of the x columns (taken before the move, dead slots
    inheriting the key of the slot before them), bounds == cumulative sums, alive particles ordered by bucket after the
    sort, nbuf == 0 when sorting again, and the same after the arrays were extended (the new dead tail inherits the last
    bucket)."""
    wl = _plasma(dim)
    eng = build_engine(wl, with_part=False)
    pg = eng.grid
    m = eng.species[0]
    eng.download_particles(0)
    eng.sync()
    rng = np.random.default_rng(5)
    for ip in range(eng.npatch):  # shuffle the slots of every patch: the first sort has work to do
        perm = rng.permutation(int(m.npart[ip]))
        for a in m.attrs:
            v = m.view(a, ip)
            v[...] = v[perm]
    eng.upload_particles(0)

    def keys_of(ip):
        x, dead = m.view("x", ip), m.view("is_dead", ip)
        with np.errstate(invalid="ignore"):
            col = np.floor((x - (pg.x0[ip] - pg.dx / 2)) / pg.dx)
        key = np.where((col >= 0) & (col < pg.nx), col, pg.nx - 1).astype(np.int64)
        run, keys = 0, np.zeros(len(x), dtype=np.int64)
        for i in range(len(x)):  # dead slots inherit the key of the slot before them
            if not dead[i]:
                run = key[i]
            keys[i] = run
        return keys, np.asarray(dead).astype(bool)
    for round_ in range(2):
        if round_ == 1:
            eng.extend(0, np.full(eng.npatch, 37, dtype=np.int64))
            eng.download_particles(0)
            eng.sync()
        before = [keys_of(ip)[0] for ip in range(eng.npatch)]
        nbuf = eng.sort(0, False)
        assert (nbuf > 0) == (round_ == 0)
        tabs = eng.sort_arrays(0)
        eng.download_particles(0)
        eng.sync()
        for ip in range(eng.npatch):
            cnt = np.bincount(before[ip], minlength=pg.nx)
            assert np.array_equal(tabs["bucket_count"][ip], cnt)
            assert np.array_equal(tabs["bound_max"][ip], np.cumsum(cnt))
            assert np.array_equal(tabs["bound_min"][ip], np.cumsum(cnt) - cnt)
            assert np.array_equal(tabs["particle_index"][ip], before[ip])
            keys, dead = keys_of(ip)
            assert np.all(np.diff(keys[~dead]) >= 0), "alive particles are ordered by bucket after the sort"
        assert eng.sort(0, False) == 0, "a sorted species has nothing to move"
    eng.close()


@pytest.mark.parametrize("dim", [2, 3])
def test_shift_by_one_cell_migration_counts(dim):
    """tests/mpi/test_syncparticles.py:63-116 on one rank: move every particle by +dx; each patch then sends exactly the
    particles of its last cell column through xmax, receives as many as its xmin neighbour sent (periodic), nothing
    crosses any other boundary, and the number of alive particles is conserved."""
    wl = _plasma(dim, temperature_eV=1.0)
    eng = build_engine(wl, with_part=False)
    pg = eng.grid
    m = eng.species[0]
    eng.download_particles(0)
    eng.sync()
    alive0 = int((m.host["is_dead"] == 0).sum())
    expect_out = np.zeros(eng.npatch, dtype=np.int64)
    for ip in range(eng.npatch):
        x = m.view("x", ip)
        x += pg.dx
        xmax_box = pg.x0[ip] + (pg.nx - 1) * pg.dx + pg.dx / 2
        expect_out[ip] = int((x > xmax_box).sum())
    eng.upload_particles(0)
    rec = eng.migrate_count(0)
    out = rec["outgoing"].reshape(eng.npatch, eng.nb)
    assert np.array_equal(out[:, 1], expect_out) and expect_out.min() > 0      # boundary 1 = xmax
    assert out[:, [b for b in range(eng.nb) if b != 1]].sum() == 0
    xmin_nbr = pg.neighbor_ipatch[:, 0]                                          # boundary 0 = xmin
    assert np.array_equal(rec["incoming"], expect_out[xmin_nbr])
    eng.extend(0, rec["to_extend"])
    eng.migrate_fill(0)
    assert eng.count_alive(0) == alive0
    eng.download_particles(0)
    eng.sync()
    for ip in range(eng.npatch):  # everybody is inside its patch box again (periodic wrap applied to the newcomers)
        x, dead = m.view("x", ip), m.view("is_dead", ip).astype(bool)
        lo, hi = pg.x0[ip] - pg.dx / 2, pg.x0[ip] + (pg.nx - 1) * pg.dx + pg.dx / 2
        assert np.all((x[~dead] >= lo) & (x[~dead] <= hi))
    eng.close()


@pytest.mark.gpu
def test_tiled_fdtd_kernels_are_bit_identical_to_the_per_cell_kernels(golden3d):
    """LPIC_FDTD_TILED=1 (shared-memory-tiled Yee update, csrc/fields.cu) in a subprocess, since the switch is read once per
    process: E and B after two half steps equal the default per-cell kernels' bit for bit."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from tests import gpu_harness as h\n"
        "g = np.load(%r)\n"
        "eng, meta = h.engine_from_golden(g, 't1')\n"
        "for _ in range(2):\n"
        "    eng.update_efield(0.5 * meta['dt']); eng.update_bfield(0.5 * meta['dt'])\n"
        "eng.download_fields()\n"
        "np.save(sys.argv[1], eng.fields_host[:6].copy())\n"
        "eng.close()\n") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                            os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_step_3d.npz"))
    outs = []
    for tiled in (False, True):
        env = dict(os.environ)
        env.pop("LPIC_FDTD_TILED", None)
        if tiled:
            env["LPIC_FDTD_TILED"] = "1"
        path = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"lpic_fdtd_{int(tiled)}.npy")
        subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=300)
        outs.append(np.load(path))
    assert np.array_equal(outs[0], outs[1])
    assert np.abs(outs[0]).max() > 0
