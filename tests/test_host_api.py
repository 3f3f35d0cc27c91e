"""CPU tests of the host-side mirror of the reference API (no GPU compute): configuration validation, patch
decomposition and neighbour tables, the host particle loader (bit-exact against the reference's numba loader on the
same seed), callback trigger rules, and the comm shim over a world_size-2 gloo group."""
import os
import subprocess
import sys

import numpy as np
import pytest

from lambdapic_b200 import Electron, Proton, Simulation, Simulation3D, callback
from lambdapic_b200.callback import _interval_triggered
from lambdapic_b200.simulation import SimulationCallbacks
from lambdapic_b200.workloads import block_rank_map, make_patch_grid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N0, D = 1.742e27, 0.8e-6 / 20


def make_sim(dim):
    if dim == 3:
        sim = Simulation3D(nx=10, ny=8, nz=12, dx=D, dy=D * 1.25, dz=D * 0.8, npatch_x=2, npatch_y=2, npatch_z=2, dt_cfl=0.95,
                           boundary_conditions={k: "periodic" for k in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")}, random_seed=1234)
        dens = lambda x, y, z: N0  # noqa: E731
    else:
        sim = Simulation(nx=16, ny=12, dx=D, dy=D * 1.25, npatch_x=2, npatch_y=3, dt_cfl=0.95,
                         boundary_conditions={k: "periodic" for k in ("xmin", "xmax", "ymin", "ymax")}, random_seed=4321)
        dens = lambda x, y: N0  # noqa: E731
    sim.add_species([Electron(density=dens, ppc=3), Proton(density=dens, ppc=2)])
    return sim


@pytest.mark.parametrize("dim,case", [(3, "golden3d"), (2, "golden2d")])
def test_loader_matches_reference_seed_for_seed(dim, case, request):
    """Same seed => bit-identical x, y, z, w, _id per patch as the reference (tests/test_random_seed.py there)."""
    g = request.getfixturevalue(case)
    sim = make_sim(dim)
    assert sim.dt == float(g["meta/dt"])
    patches = sim.create_patches(sim._grid(0, 1))
    for s in sim.species:
        patches.add_species(s)
    patches.fill_particles(np.random.default_rng(sim.random_seed).spawn(1)[0])
    assert np.array_equal(np.stack([p.neighbor_ipatch for p in patches]), g["meta/neighbor_ipatch"])
    for ip, p in enumerate(patches):
        for s in range(2):
            for a in ("x", "y", "z", "w", "_id"):
                if dim == 2 and a == "z":
                    continue
                assert np.array_equal(getattr(p.particles[s], a).view(np.uint64), g[f"t0/p/{ip}/{s}/{a}"].view(np.uint64)), (ip, s, a)


def test_config_validation():
    per = {k: "periodic" for k in ("xmin", "xmax", "ymin", "ymax")}
    with pytest.raises(ValueError):
        Simulation(nx=10, ny=8, dx=D, dy=D, npatch_x=3, npatch_y=2, boundary_conditions=per)
    with pytest.raises(ValueError):
        Simulation(nx=16, ny=16, dx=D, dy=D, npatch_x=2, npatch_y=2, nsteps=1, sim_time=1e-15, boundary_conditions=per)
    sim = Simulation(nx=16, ny=16, dx=D, dy=D, npatch_x=2, npatch_y=2)  # default boundaries: CPML on every side
    assert sim._periodic()[:2] == (False, False) and sim.cpml_thickness == 6
    grid = sim._grid(0, 1)
    assert (grid.neighbor_ipatch[0] >= 0).sum() == 3  # a corner patch of an open 2x2 domain keeps 3 neighbours
    with pytest.raises(ValueError):
        Simulation(nx=16, ny=16, dx=D, dy=D, npatch_x=2, npatch_y=2, boundary_conditions={"xmin": "periodic", "xmax": "pml"})
    sim = Simulation(nx=64, ny=32, dx=D, dy=D, boundary_conditions=per)
    assert sim.npatch_x == 4 and sim.npatch_y == 2  # auto patching: 16-cell tiles
    assert sim.STAGES[0] == "init" and sim.DEFAULT_STAGE == "end" and len(sim.STAGES) == 14
    sim.add_species([Electron(density=lambda x, y: N0, ppc=1)])
    with pytest.raises(ValueError):
        sim.add_species([Electron(density=lambda x, y: N0, ppc=1)])
    with pytest.raises(ValueError):
        sim.add_species([Proton(density=lambda x, y, z: N0, ppc=1)])


def test_callbacks_stage_and_interval_rules():
    per = {k: "periodic" for k in ("xmin", "xmax", "ymin", "ymax")}
    sim = Simulation(nx=32, ny=32, dx=D, dy=D, npatch_x=2, npatch_y=2, boundary_conditions=per)

    @callback("maxwell_1", interval=3)
    def a(sim):
        pass

    def plain(sim):
        pass
    cbs = SimulationCallbacks([a, plain], sim)
    assert cbs.non_empty_stages() == ["maxwell_1", "end"]
    sim.itime = 3
    assert cbs.has_triggered_callbacks("maxwell_1")
    sim.itime = 4
    assert not cbs.has_triggered_callbacks("maxwell_1") and cbs.has_triggered_callbacks("end")
    sim.time, sim.dt = 2.05e-15, 1e-16
    assert _interval_triggered(sim, 1e-15) and not _interval_triggered(sim, 0.8e-15)
    with pytest.raises(ValueError):
        SimulationCallbacks([callback("nonsense")(plain)], sim)
    with pytest.raises(ValueError):
        callback("end", interval=0)(plain)


def test_block_partition_and_remote_tables():
    rk = block_rank_map(4, 4, 4, 8)
    assert sorted(np.bincount(rk)) == [8] * 8
    grids = [make_patch_grid(3, 4, 4, 4, 8, 8, 8, 1.0, 1.0, 1.0, rank=r, nranks=8) for r in range(8)]
    seen = np.concatenate([g.index for g in grids])
    assert sorted(seen) == list(range(64))
    for g in grids:
        local, remote = g.neighbor_ipatch >= 0, g.neighbor_rank >= 0
        assert not (local & remote).any() and (local | remote).all()  # periodic: every neighbour exists somewhere
        # the remote position really addresses that patch on its owner
        for k in range(g.npatch):
            for b in np.nonzero(remote[k])[0]:
                owner = grids[g.neighbor_rank[k, b]]
                assert owner.index[g.remote_ipatch[k, b]] == g.neighbor_index[k, b]


def test_comm_shim_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import torch.distributed as dist\n"
        "from lambdapic_b200.comm import default_comm\n"
        "dist.init_process_group('gloo')\n"
        "c = default_comm(); r = c.Get_rank()\n"
        "assert c.Get_size() == 2\n"
        "assert c.bcast({'a': r}, root=0) == {'a': 0}\n"
        "assert c.scatter([['p0'], ['p1']] if r == 0 else None, root=0) == [f'p{r}']\n"
        "assert c.allgather(r) == [0, 1]\n"
        "assert c.allreduce(r + 1) == 3\n"
        "g = c.gather(r * 10, root=0); assert (g == [0, 10]) if r == 0 else g is None\n"
        "c.Barrier(); dist.destroy_process_group(); print('ok', r)\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29517", str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.count("ok") == 2


def test_mirror_selection_of_callback_hints():
    """reads / writes hints of @callback (mirror elision): names -> (field mask, psi, particles)."""
    from lambdapic_b200 import callback
    from lambdapic_b200._lib import FIELD_ATTRS
    from lambdapic_b200.device import DeviceBridge
    from lambdapic_b200.engine import ALL_FIELDS
    sel = DeviceBridge._selection
    assert sel(None) == (ALL_FIELDS, True, True)
    assert sel(()) == (0, False, False)
    assert sel(("ex", "rho")) == ((1 << FIELD_ATTRS.index("ex")) | (1 << FIELD_ATTRS.index("rho")), False, False)
    assert sel(("fields", "psi")) == (ALL_FIELDS, True, False)
    assert sel({"particles"}) == (0, False, True)
    with pytest.raises(ValueError):
        sel(("exx",))

    @callback("end", interval=2, reads=("ex",), writes=())
    def diag(sim):
        pass
    assert diag.reads == ("ex",) and diag.writes == () and diag.needs_host is True


@pytest.mark.parametrize("dim,npatch,periodic", [(2, (4, 3, 1), (True, False, True)), (2, (3, 2, 1), (False, True, True)),
                                                  (3, (2, 3, 2), (True, True, False)), (3, (3, 1, 2), (False, False, True))])
def test_neighbour_tables_rebuilt_from_patch_positions_match_the_grid_builder(dim, npatch, periodic):
    """Patches.init_rect_neighbor_index_* (used again after a MovingWindow shift, core/patch/patch.py:446-592) gives the
    same tables as the block-partition builder for an unshifted grid, and follows the patches when a column rotates."""
    from lambdapic_b200.patch import Patch2D, Patch3D, Patches
    npx, npy, npz = npatch
    bc = {f"{a}{s}": ("periodic" if periodic[i] else "pml") for i, a in enumerate("xyz"[:dim]) for s in ("min", "max")}
    pg = make_patch_grid(dim, npx, npy, npz, 8, 8, 8 if dim == 3 else 1, 1.0, 1.0, 1.0, 3, periodic)
    ps = Patches(dim)
    for k, gidx in enumerate(pg.index):
        ix, iy, iz = int(gidx % npx), int((gidx // npx) % npy), int(gidx // (npx * npy))
        p = (Patch3D(0, int(gidx), ix, iy, iz, 0.0, 0.0, 0.0, 8, 8, 8, 1.0, 1.0, 1.0) if dim == 3
             else Patch2D(0, int(gidx), ix, iy, 0.0, 0.0, 8, 8, 1.0, 1.0))
        ps.append(p)
    if dim == 3:
        ps.init_rect_neighbor_index_3d(npx, npy, npz, boundary_conditions=bc)
        ps.init_neighbor_ipatch_3d()
        ps.init_neighbor_rank_3d()
    else:
        ps.init_rect_neighbor_index_2d(npx, npy, boundary_conditions=bc)
        ps.init_neighbor_ipatch_2d()
        ps.init_neighbor_rank_2d()
    assert np.array_equal(np.stack([p.neighbor_index for p in ps]), pg.neighbor_index)
    assert np.array_equal(np.stack([p.neighbor_ipatch for p in ps]), pg.neighbor_ipatch)
    assert all((p.neighbor_rank == -1).all() for p in ps)
    # rotate the columns as MovingWindow._shift does: column 0 becomes the last one
    for p in ps:
        p.ipatch_x = npx - 1 if p.ipatch_x == 0 else p.ipatch_x - 1
    axes = "xyz"[:dim]
    index_map = {tuple(getattr(p, f"ipatch_{a}") for a in axes): p.index for p in ps}
    (ps.init_rect_neighbor_index_3d if dim == 3 else ps.init_rect_neighbor_index_2d)(
        *npatch[:dim], boundary_conditions=bc, patch_index_map=index_map)
    where = {p.index: p for p in ps}
    for p in ps:
        right = p.neighbor_index[1]  # xmax
        if p.ipatch_x == npx - 1 and not periodic[0]:
            assert right == -1
        else:
            q = where[right]
            assert q.ipatch_x == (p.ipatch_x + 1) % npx and q.ipatch_y == p.ipatch_y


def test_stage_mirror_traffic_follows_the_callback_hints():
    """Simulation._stage: device-side callbacks cause no mirror traffic; hinted callbacks move only the named arrays and keep
    the device authoritative (resident stays True, so sim.energies() in the same stage works on device state); one unhinted
    callback in the stage falls back to the full download / upload with the host authoritative in between."""
    from lambdapic_b200.comm import default_comm
    from lambdapic_b200.operators import SingleRankMPI

    class FakeBridge:
        def __init__(self):
            self.resident, self.log, self.seen_resident = True, [], []

        def download(self, names=None):
            self.log.append(("down", None if names is None else frozenset(names)))

        def upload(self, names=None):
            self.log.append(("up", None if names is None else frozenset(names)))
    d = 0.8e-6 / 20
    sim = Simulation(nx=16, ny=16, dx=d, dy=d, npatch_x=1, npatch_y=1,
                     boundary_conditions={k: "periodic" for k in ("xmin", "xmax", "ymin", "ymax")})
    sim.mpi = SingleRankMPI(default_comm())
    br = sim.bridge = FakeBridge()

    def note(sim):
        br.seen_resident.append(br.resident)
    dev = callback("end", needs_host=False)(note)
    hinted = callback("end", reads=("ex",), writes=("bz",))(note)
    ro = callback("end", reads=("rho", "ex"), writes=())(note)
    plain = callback("end")(note)
    skipped = callback("end", interval=lambda s: False)(note)

    def run(cbs):
        br.log.clear(); br.seen_resident.clear(); br.resident = True
        sim._stage(SimulationCallbacks(cbs, sim), "end", "end")
        return list(br.log), list(br.seen_resident), br.resident
    assert run([dev]) == ([], [True], True)
    assert run([skipped, plain][:1]) == ([], [], True)  # nothing triggered: nothing moves, nothing runs
    log, seen, after = run([dev, hinted, ro])
    assert log == [("down", frozenset({"ex", "bz", "rho"})), ("up", frozenset({"bz"}))] and seen == [True, True, True] and after
    log, seen, after = run([ro])
    assert log == [("down", frozenset({"rho", "ex"})), ("up", frozenset())] and after
    log, seen, after = run([hinted, plain])
    assert log == [("down", None), ("up", None)] and seen == [False, False] and after


def test_timer_records_have_the_format_timer_stat_parses(tmp_path):
    """enable_timer writes `<log>.timer.txt` with the reference's TIMER records (core/utils/timer.py:84-96); the pattern is
    the one `lambdapic timer-stat` uses (cli/stat.py:12)."""
    import re
    import time
    from lambdapic_b200.simulation import Timer
    Timer.enabled, Timer.sync = True, None
    path = Timer.open_sink(str(tmp_path / "log.txt"))
    assert path.endswith("log.timer.txt")
    with Timer("unified pusher for electron"):
        time.sleep(0.002)
    with Timer("too short to log"):
        pass
    Timer.sink.close()
    Timer.sink, Timer.enabled = None, False
    lines = [ln for ln in open(path) if "TIMER" in ln]
    assert len(lines) == 1
    m = re.search(r"Rank \d+ (.*?) took ([\d.]+)ms", lines[0].split("|")[-1].strip())
    assert m and m.group(1) == "unified pusher for electron" and float(m.group(2)) >= 2.0


def test_particle_transfers_batch_the_record_attributes():
    """Engine.upload_particles / download_particles move x y z w ux uy uz inv_gamma (the device's 64-byte records) through ONE
    lpic_*_particle_records call with the right mask and pointer table, and everything else attribute by attribute."""
    import ctypes as C
    import types
    from lambdapic_b200 import engine as E
    calls = []

    class FakeLib:
        def __getattr__(self, name):
            def f(*args):
                calls.append((name, args))
                return 0
            return f

    attrs = ["x", "y", "z", "w", "ux", "uy", "uz", "inv_gamma", "_id"]
    host = {a: np.zeros(8) for a in attrs}
    host["is_dead"] = np.zeros(8, dtype=np.uint8)
    eng = types.SimpleNamespace(L=FakeLib(), ctx=None, species=[types.SimpleNamespace(attrs=attrs, host=host)])
    eng._record_table = types.MethodType(E.DeviceEngine._record_table, eng)
    E.DeviceEngine.upload_particles(eng, 0)
    names = [c[0] for c in calls]
    assert names == ["lpic_upload_particle_records", "lpic_upload_particles", "lpic_upload_particles"]
    _, (_, ispec, mask, tab) = calls[0]
    assert ispec == 0 and mask == 0xFF
    assert [tab[i] for i in range(8)] == [host[a].ctypes.data for a in attrs[:8]]
    assert sorted(c[1][2] for c in calls[1:]) == sorted([E.PART_ATTRS.index("_id"), E.P_IS_DEAD])
    calls.clear()
    E.DeviceEngine.download_particles(eng, 0, attrs=["ux", "is_dead"])
    assert [c[0] for c in calls] == ["lpic_download_particle_records", "lpic_download_particles"]
    assert calls[0][1][2] == 1 << E.PART_ATTRS.index("ux") and calls[1][1][2] == E.P_IS_DEAD
