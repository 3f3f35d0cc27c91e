"""torchrun worker: MovingWindow on N real ranks.  The 2D moving-window case of tests/golden/ref_mw_2d.npz
(x open with CPML until the window starts, window at c from t = 0, 4 x 2 patches) is started from the golden t0 state on
N ranks (static block partition, NcclExchange) and, on rank 0's GPU, on ONE rank; 24 steps with three column recycles.
Per global patch the N-rank fields / particle sets must equal the 1-rank ones (<= 1e-10, ids exact).  The reference draws
the recycled patches' new particles from each rank's own generator stream (simulation.py:700-716), so a 1-rank and an
N-rank run of the reference itself differ there; to compare, both runs here seed the loader of shift k with (seed, k)
-- the draw order inside a shift is the reference's.  Exercised: rotated origins, neighbour tables allgathered over the
ranks, the exchange plan rebuilt after every shift with the communicator kept, fields and psi cleared on the device,
only the recycled patches' particles uploaded."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(comm, device, g):
    from lambdapic_b200 import Electron, MovingWindow, Proton, Simulation, callback
    from lambdapic_b200._lib import FIELD_ATTRS, PART_ATTRS
    d, n0 = 0.8e-6 / 20, 1.742e27
    sim = Simulation(nx=32, ny=16, dx=d, dy=d * 1.25, npatch_x=4, npatch_y=2, dt_cfl=0.95,
                     boundary_conditions=dict(xmin="pml", xmax="pml", ymin="periodic", ymax="periodic"),
                     cpml_thickness=6, random_seed=4322, comm=comm, device=device)
    dens = lambda x, y: n0 * (1.0 + x * 2.0e5)  # noqa: E731
    sim.add_species([Electron(density=dens, ppc=2), Proton(density=dens, ppc=1)])
    sim.initialize()

    @callback("init")
    def load_golden(sim):
        for p in sim.patches:
            gp = p.index
            for a in FIELD_ATTRS:
                getattr(p.fields, a)[...] = g[f"t0/f/{gp}/{a}"]
            for s, part in enumerate(p.particles):
                assert part.npart == g[f"t0/p/{gp}/{s}/x"].size
                for a in PART_ATTRS:
                    getattr(part, a)[...] = g[f"t0/p/{gp}/{s}/{a}"]
                part.is_dead[...] = g[f"t0/p/{gp}/{s}/is_dead"].astype(bool)
    class SeededWindow(MovingWindow):
        def _fill_particles(self, sim, new_patches):
            keep = sim.rand_gen
            sim.rand_gen = np.random.default_rng([4322, self.num_shifts])  # same stream whatever the rank count
            try:
                super()._fill_particles(sim, new_patches)
            finally:
                sim.rand_gen = keep
    return sim, load_golden, SeededWindow(velocity=299792458.0, start_time=0.0)


def snapshot(sim):
    from lambdapic_b200._lib import FIELD_ATTRS
    out = {}
    for p in sim.patches:
        d = {a: np.array(getattr(p.fields, a)) for a in FIELD_ATTRS}
        d["origin"], d["column"] = float(p.x0), int(p.ipatch_x)
        for s, part in enumerate(p.particles):
            alive = ~np.asarray(part.is_dead)
            ids = part._id.view(np.uint64)[alive] & np.uint64((1 << 50) - 1)
            o = np.argsort(ids)
            d[f"ids{s}"] = ids[o]
            for a in ("x", "y", "ux", "uy", "uz", "w"):
                d[f"{a}{s}"] = np.asarray(getattr(part, a))[alive][o]
        out[int(p.index)] = d
    return out


def main():
    import torch
    import torch.distributed as dist
    from lambdapic_b200.comm import SingleComm, TorchComm
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_mw_2d.npz"))
    nsteps = int(g["meta/nsteps"])
    sim, load, mw = build(TorchComm(), local, g)
    use_mw = not os.environ.get("LPIC_TEST_NO_WINDOW")  # debugging aid: the same case without the window
    nsteps = int(os.environ.get("LPIC_TEST_STEPS", nsteps))
    shifts = 0
    for it in range(nsteps):
        before = [p.x0 for p in sim.patches]
        sim.run(nsteps=1, callbacks=([load] if it == 0 else []) + ([mw] if use_mw else []))
        shifts += before != [p.x0 for p in sim.patches]
    mine = snapshot(sim)
    recycles = sim.bridge.stats.get("recycles", 0)
    sim.bridge.close()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    ok = [True]
    if rank == 0:
        ref_sim, load1, mw1 = build(SingleComm(), local, g)
        for it in range(nsteps):
            ref_sim.run(nsteps=1, callbacks=([load1] if it == 0 else []) + ([mw1] if use_mw else []))
        ref = snapshot(ref_sim)
        ref_sim.bridge.close()
        worst, n, bad = 0.0, 0, []
        try:
            for part in gathered:
                for gp, d in part.items():
                    n += 1
                    r = ref[gp]
                    assert d["origin"] == r["origin"] and d["column"] == r["column"], ("origin", gp, d["origin"], r["origin"])
                    for k, v in d.items():
                        if k in ("origin", "column"):
                            continue
                        if k.startswith("ids"):
                            assert np.array_equal(v, r[k]), ("particle set", gp, k)
                            continue
                        scale = float(np.abs(r[k]).max()) if r[k].size else 0.0
                        e = float(np.abs(v - r[k]).max()) / scale if scale > 0 else 0.0
                        worst = max(worst, e)
                        if e > 1e-10:
                            bad.append((gp, k, float(f"{e:.2e}")))
            assert not bad, bad[:12]
            assert n == len(ref) and (shifts >= 2 or not use_mw or nsteps < 20)
            print(f"nccl-mw-parity ok {world} ranks vs 1 rank: {n} patches, {shifts} shifts, worst {worst:.2e}", flush=True)
        except AssertionError as exc:
            ok[0] = False
            print("nccl-mw-parity FAILED", exc, flush=True)
    dist.broadcast_object_list(ok, src=0)
    print(f"nccl-mw rank {rank}/{world}: {shifts} shifts, {recycles} device-side recycles", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok[0] else 1)


if __name__ == "__main__":
    main()
