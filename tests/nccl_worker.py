"""torchrun worker: N real ranks (one GPU each, NCCL) advance the golden 3D case three steps; every rank checks its
own patches against the 1-rank reference golden vectors.  Launched by tests/test_gpu_nccl.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from lambdapic_b200.multigpu import RankProgram, drive_nccl, torch_alloc
    from tests import gpu_harness as h
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_step_3d.npz"))
    import lambdapic_b200.engine as E
    _orig = E.DeviceEngine.__init__

    def _init(self, *a, **k):  # every engine of this process lives on the process's GPU
        k["device"] = local
        _orig(self, *a, **k)
    E.DeviceEngine.__init__ = _init
    from lambdapic_b200.multigpu import NcclExchange, _torch_bcast
    rev = None
    # both transports: the library's own NCCL path (csrc/comm.cu, asynchronous start/wait) and the host-driven one
    for mode in ("library", "host"):
        engines, grids, meta = h.split_engines_from_golden(g, "t0", world)
        # keep only this rank's engine (the helper builds all of them; the others are closed at once)
        for r, e in enumerate(engines):
            if r != rank:
                e.close()
        eng, pg = engines[rank], grids[rank]
        rev = [bool(int(g[f"t1/reverse_x/{s}"])) for s in range(eng.nspec)]
        if mode == "library":
            prog = NcclExchange(eng, pg, _torch_bcast)
        else:
            prog = RankProgram(eng, pg, torch_alloc(torch.device("cuda", local)))
        for k in range(3):
            if mode == "library":
                prog.step(meta["dt"], meta["q"], meta["m"], rev)
            else:
                drive_nccl(prog.step(meta["dt"], meta["q"], meta["m"], rev), rank)
            worst = h.compare_split_state_with_golden([eng], [pg], g, f"t{k + 1}", rtol=1e-11)
        dist.barrier()
        print(f"nccl-parity-{mode} ok rank {rank}/{world} worst {worst:.2e} bytes_sent {prog.bytes_sent}", flush=True)
        eng.close()
    print(f"nccl-parity ok rank {rank}/{world}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
