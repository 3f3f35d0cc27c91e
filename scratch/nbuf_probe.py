import sys; sys.path.insert(0,'.')
import numpy as np
from lambdapic_b200.workloads import ThermalPlasma, build_engine
wl = ThermalPlasma(cells=(64,64,64))
eng = build_engine(wl)
pg = eng.grid
def analyse(s, p=5):
    m = eng.species[s]; m.refresh_layout()
    eng.download_particles(s, ["x"] ); eng.download_particles(s, ["is_dead"]); eng.sync()
    x = m.view("x", p).copy(); dead = m.view("is_dead", p).copy()
    n = len(x)
    x0 = pg.x0[p] - pg.dx/2
    key = np.floor((x - x0)/pg.dx)
    key = np.where((key>=0)&(key<pg.nx), key, pg.nx-1).astype(int)
    run = 0; keys = np.zeros(n, int)
    for i in range(n):
        if not dead[i]: run = key[i]
        keys[i] = run
    cnt = np.bincount(keys, minlength=pg.nx); bmax = np.cumsum(cnt)
    owner = np.searchsorted(bmax, np.arange(n), side='right')
    mis = keys != owner
    print(f"spec {s} patch {p}: npart {n} dead {int(dead.sum())} misplaced {int(mis.sum())} counts {cnt.tolist()}")
    idx = np.nonzero(mis)[0]
    if len(idx): print("  first misplaced slots", idx[:10].tolist(), "keys", keys[idx[:10]].tolist(), "owner", owner[idx[:10]].tolist(), " last", idx[-5:].tolist())
for it in range(3):
    r = eng.step(wl.dt, wl.q, wl.m, [False, False])
    print(it, "nbuf", r[0], flush=True)
    for s in (0,1): analyse(s)
