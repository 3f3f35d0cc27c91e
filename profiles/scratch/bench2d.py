"""2D A/B of the particle kernels: ThermalPlasma(dim=2, 2048 x 2048, 32+32 ppc), per-operator CUDA-event times of step N."""
import os, sys
sys.path.insert(0, "/root/repo")
from lambdapic_b200.workloads import ThermalPlasma, build_engine
wl = ThermalPlasma(dim=2, cells=(2048, 2048, 1), patch=(16, 16, 1), ppc=(32, 32))
eng = build_engine(wl)
rev = [False, False]
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    eng.step(wl.dt, wl.q, wl.m, rev)
bd = eng.step_profiled(wl.dt, wl.q, wl.m, rev)
tot = sum(t for _, t in bd)
for name, t in bd:
    if "push" in name or "sort" in name or "sync_part" in name:
        print(f"  {name:28s} {t:8.3f} ms")
print(f"  TOTAL {tot:8.3f} ms  {wl.n_particles() / tot * 1e3:.3e} particle-updates/s  mode={'sorted' if os.environ.get('LPIC_PUSH_SORTED') else 'tile'}")
eng.close()
