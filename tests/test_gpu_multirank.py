"""Multi-rank path on ONE GPU: N emulated ranks (one DeviceEngine each) advanced in lockstep by the in-process driver,
which hands the packed staging buffers over by copy.  The same RankProgram runs under NCCL in bench.py / Simulation.
Bar: the N-rank result equals the 1-rank reference result on the same global patch grid -- particle
sets per patch identical by _id, fields / currents / momenta <= 1e-11 after 3 steps (summation order at rank
boundaries differs from the single-rank order)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["golden3d", "golden2d"])
@pytest.mark.parametrize("nranks", [2, 4])
def test_emulated_ranks_match_single_rank_reference(case, nranks, request):
    import torch
    from lambdapic_b200.multigpu import RankProgram, drive_in_process, torch_alloc
    from tests import gpu_harness as h
    g = request.getfixturevalue(case)
    if case == "golden2d" and nranks == 4:
        pytest.skip("2x3 patch grid does not split into 4 blocks")
    engines, grids, meta = h.split_engines_from_golden(g, "t0", nranks)
    alloc = torch_alloc(torch.device("cuda", 0))
    progs = [RankProgram(e, pg, alloc) for e, pg in zip(engines, grids)]
    rev = [bool(int(g[f"t1/reverse_x/{s}"])) for s in range(engines[0].nspec)]
    moved = 0
    for k in range(3):
        res = drive_in_process([p.step(meta["dt"], meta["q"], meta["m"], rev) for p in progs])
        moved += sum(r[1][s]["sent"] for r in res for s in range(engines[0].nspec))
        h.compare_split_state_with_golden(engines, grids, g, f"t{k + 1}", rtol=1e-11)
    assert moved > 0, "the case must exercise inter-rank migration"
    assert sum(p.bytes_sent for p in progs) > 0
    for e in engines:
        e.close()


def test_emulated_ranks_with_arena_relayout_during_remote_migration(golden3d):
    """Zero physical slack: the first remote arrivals force lpic_species_extend to re-lay-out the arena in the middle of
    the inter-rank migration (extend -> relist -> pack -> unpack)."""
    import torch
    from lambdapic_b200.multigpu import RankProgram, drive_in_process, torch_alloc
    from tests import gpu_harness as h
    g = golden3d
    engines, grids, meta = h.split_engines_from_golden(g, "t0", 2, slack=1.0, min_extra=0)
    progs = [RankProgram(e, pg, torch_alloc(torch.device("cuda", 0))) for e, pg in zip(engines, grids)]
    rev = [bool(int(g[f"t1/reverse_x/{s}"])) for s in range(engines[0].nspec)]
    grew = False
    for k in range(3):
        res = drive_in_process([p.step(meta["dt"], meta["q"], meta["m"], rev) for p in progs])
        grew = grew or any(r[1][s]["extended"] or r[1][s]["local"]["moved"] for r in res for s in range(engines[0].nspec))
        h.compare_split_state_with_golden(engines, grids, g, f"t{k + 1}", rtol=1e-11)
    assert grew
    for e in engines:
        e.close()


def test_exchange_plan_lines_up_between_ranks():
    """Sender and receiver derive the same entry order and sizes without negotiation (runs on the GPU box because the
    plan is registered with the device library, which also checks it)."""
    import torch
    from lambdapic_b200.engine import DeviceEngine
    from lambdapic_b200.multigpu import RankProgram, torch_alloc
    from lambdapic_b200.workloads import make_patch_grid
    progs = []
    for r in range(8):
        pg = make_patch_grid(3, 4, 4, 2, 6, 5, 7, 1.0, 1.0, 1.0, rank=r, nranks=8)
        eng = DeviceEngine(3, pg.npatch, 6, 5, 7, 3, 1.0, 1.0, 1.0, 1)
        eng.set_geometry(pg.x0, pg.y0, pg.z0, pg.neighbor_ipatch, pg.boxes, pg.glob, r, pg.index)
        progs.append(RankProgram(eng, pg, torch_alloc(torch.device("cuda", 0))))
    for a, pa in enumerate(progs):
        for i, b in enumerate(pa.peers):
            pb = progs[b]
            j = pb.peers.index(a)
            assert pa.send_words[i] == pb.recv_words[j] and pa.nsend[i] == pb.nrecv[j]
            for (p, bd), (q, bq) in zip(pa.send_entries[b], pb.recv_entries[a]):
                assert pa.grid.neighbor_index[p, bd] == pb.grid.index[q]
                assert pb.grid.neighbor_index[q, bq] == pa.grid.index[p]
    for p in progs:
        p.eng.close()


def test_emulated_ranks_stay_equal_to_one_rank_over_twenty_steps():
    """Longer than the golden snapshots reach: a 4x4x2 grid of 8^3 patches, 20 keV, 20 steps on 4 emulated ranks against the
    same 20 steps on ONE engine (no oracle involved: N ranks = 1 rank).  Particles cross rank boundaries every step, patches
    grow, the sorters move slots; fields must stay <= 1e-10, particle sets per patch identical, their attributes <= 1e-10."""
    import torch
    from lambdapic_b200.multigpu import RankProgram, drive_in_process, torch_alloc
    from lambdapic_b200.workloads import ThermalPlasma, build_engine
    from lambdapic_b200._lib import FIELD_ATTRS
    wl = ThermalPlasma(dim=3, cells=(32, 32, 16), patch=(8, 8, 8), ppc=(4, 2), temperature_eV=2.0e4)
    nranks, nsteps, rev = 4, 20, [False, False]
    one = build_engine(wl, rank=0, nranks=1)
    engines = [build_engine(wl, rank=r, nranks=nranks) for r in range(nranks)]
    progs = [RankProgram(e, e.grid, torch_alloc(torch.device("cuda", 0))) for e in engines]
    sent = 0
    for _ in range(nsteps):
        one.step(wl.dt, wl.q, wl.m, rev)
        res = drive_in_process([p.step(wl.dt, wl.q, wl.m, rev) for p in progs])
        sent += sum(r[1][s]["sent"] for r in res for s in range(2))
    assert sent > 0
    one.download_all()
    where = {int(gp): k for k, gp in enumerate(one.grid.index)}
    low50 = np.uint64((1 << 50) - 1)  # the ids carry the rank in their top bits
    worst = 0.0
    for e in engines:
        e.download_all()
        for k, gp in enumerate(e.grid.index):
            k1 = where[int(gp)]
            for a in FIELD_ATTRS:
                ref, got = one.field_view(a, k1), e.field_view(a, k)
                scale = float(np.abs(ref).max())
                err = float(np.abs(got - ref).max()) / scale if scale > 0 else 0.0
                worst = max(worst, err)
                assert err <= 1e-10, (a, int(gp), err)
            for s in range(2):
                m, m1 = e.species[s], one.species[s]
                alive, alive1 = ~m.view("is_dead", k), ~m1.view("is_dead", k1)
                ids, ids1 = m.view("_id", k).view(np.uint64)[alive] & low50, m1.view("_id", k1).view(np.uint64)[alive1] & low50
                assert np.array_equal(np.sort(ids), np.sort(ids1)), ("particle set", int(gp), s)
                o, o1 = np.argsort(ids), np.argsort(ids1)
                for a in ("x", "y", "z", "ux", "uy", "uz"):
                    ref, got = m1.view(a, k1)[alive1][o1], m.view(a, k)[alive][o]
                    scale = float(np.abs(ref).max()) if ref.size else 0.0
                    err = float(np.abs(got - ref).max()) / scale if scale > 0 else 0.0
                    worst = max(worst, err)
                    assert err <= 1e-10, (a, int(gp), s, err)
    for e in engines + [one]:
        e.close()
