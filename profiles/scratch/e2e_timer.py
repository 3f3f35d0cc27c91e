"""Where does a step of Simulation3D.run go?  Per-operator device times (enable_timer serialises the step) and the host-side
profile of an untimed run, on the 128^3 thermal box of the bench's e2e leg."""
import cProfile, pstats, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import lambdapic_b200 as lp
from lambdapic_b200.simulation import Timer
from lambdapic_b200.workloads import ThermalPlasma

def make(timer):
    wl = ThermalPlasma(dim=3, cells=(128, 128, 128), patch=(16, 16, 16), ppc=(32, 32))
    per = {k: "periodic" for k in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")}
    sim = lp.Simulation3D(nx=128, ny=128, nz=128, dx=wl.d, dy=wl.d, dz=wl.d, npatch_x=8, npatch_y=8, npatch_z=8, dt_cfl=wl.dt_cfl,
                          boundary_conditions=per, random_seed=wl.seed, store_part_fields=False, enable_timer=timer)
    sim.add_species([lp.Electron(density=wl.density, ppc=wl.ppc[0]), lp.Proton(density=wl.density, ppc=wl.ppc[1])])
    sim.initialize()
    for sp in sim.species:
        lp.SetTemperature(sp, wl.temperature_eV)(sim)
    return sim

hist = []
@lp.callback("end", needs_host=False)
def diag(sim):
    hist.append(sim.energies())

sim = make(True)
sim.run(nsteps=10, callbacks=[diag])
tot = sorted(Timer.totals.items(), key=lambda kv: -kv[1])
print("== per-operator wall time with the timer on (10 steps, device synchronised per interval), ms per step")
for k, v in tot[:25]:
    print(f"  {k:45s} {1e2 * v:9.2f}")
sim.bridge.close()
sim = make(False)
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
sim.run(nsteps=10, callbacks=[diag])
pr.disable()
print("== untimed run: %.1f ms per step incl. entry/exit copies" % (1e2 * (time.perf_counter() - t0)))
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
