import sys; sys.path.insert(0,'.')
import numpy as np
from lambdapic_b200.workloads import ThermalPlasma, build_engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
wl = ThermalPlasma(dim=2, cells=(n, n, 1), patch=(16, 16, 1), ppc=(32, 32))
for mode in ("cell", "slot"):
    eng = build_engine(wl)
    eng.slot_order = mode == "slot"
    for _ in range(4):
        eng.step(wl.dt, wl.q, wl.m, [False, False])
    eng.sync()
    bd = eng.step_profiled(wl.dt, wl.q, wl.m, [False, False])
    tot = sum(t for _, t in bd)
    npart = sum(eng.count_alive(s) for s in range(eng.nspec))
    print(mode, f"particles {npart:.3e} step {tot:.2f} ms -> {npart/tot*1e3:.3e} updates/s", {k: round(v, 2) for k, v in bd if 'push' in k or 'sort' in k or 'sync_particles' in k})
    eng.close()
