"""torchrun worker: the PUBLIC API on N real ranks.  Each rank builds `Simulation3D` on its own GPU (static block
partition, MultiRankMPI over NCCL), an `init` callback overwrites its patches with the reference's golden t0 state (by
global patch index), three steps run, and every rank compares its patches with the 1-rank reference golden vectors
(particle sets by _id, fields <= 1e-11).  Launched by tests/test_gpu_nccl.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from lambdapic_b200 import Electron, Proton, Simulation3D, callback
    from lambdapic_b200._lib import FIELD_ATTRS, PART_ATTRS
    from tests.parity import rel_err
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_step_3d.npz"))
    d, n0 = 0.8e-6 / 20, 1.742e27
    sim = Simulation3D(nx=10, ny=8, nz=12, dx=d, dy=d * 1.25, dz=d * 0.8, npatch_x=2, npatch_y=2, npatch_z=2, dt_cfl=0.95,
                       boundary_conditions={k: "periodic" for k in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")},
                       random_seed=1234, device=local)
    sim.add_species([Electron(density=lambda x, y, z: n0, ppc=3), Proton(density=lambda x, y, z: n0, ppc=2)])

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
    worst = 0.0
    for it in range(3):
        sim.run(nsteps=1, callbacks=[load_golden] if it == 0 else [])
        for s in range(2):
            assert sim.sorter[s].reverse_x == bool(int(g[f"t1/reverse_x/{s}"])), "global drift decision"
        for p in sim.patches:
            gp = p.index
            for a in FIELD_ATTRS:
                e = rel_err(getattr(p.fields, a), g[f"t{it + 1}/f/{gp}/{a}"])
                worst = max(worst, e)
                assert e <= 1e-11, (it, gp, a, e)
            for s, part in enumerate(p.particles):
                alive, ralive = ~np.asarray(part.is_dead), ~g[f"t{it + 1}/p/{gp}/{s}/is_dead"].astype(bool)
                ids, rids = part._id.view(np.uint64)[alive], g[f"t{it + 1}/p/{gp}/{s}/_id"].view(np.uint64)[ralive]
                assert np.array_equal(np.sort(ids), np.sort(rids)), (it, gp, s)
                o, ro = np.argsort(ids), np.argsort(rids)
                for a in ("x", "y", "z", "ux", "uy", "uz", "w"):
                    e = rel_err(np.asarray(getattr(part, a))[alive][o], g[f"t{it + 1}/p/{gp}/{s}/{a}"][ralive][ro])
                    worst = max(worst, e)
                    assert e <= 1e-11, (it, gp, s, a, e)
    dist.barrier()
    print(f"nccl-sim-parity ok rank {rank}/{world} npatch {sim.patches.npatches} worst {worst:.2e}", flush=True)
    sim.bridge.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
