#!/usr/bin/env python
"""BASELINE.json configs[3] -- the reference's `example/laser-target-3d.py` with the package import switched:

    512 x 256 x 256 cells, dx = lambda/20, dy = dz = lambda/10, CPML on all six sides, n_c plasma for x > 1 um
    (electrons + protons, 2 ppc each, ~1.28e8 particles), Gaussian laser a0 = 10 from the xmin antenna, 1001 steps.

Same constructor calls as the reference script; its HDF5 / plotting callbacks are replaced by a device-side energy
diagnostic and a read-only probe of `ey` (`reads=` hint).

    python examples/laser_target_3d.py [--nsteps 1001]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lambdapic_b200 import Electron, GaussianLaser3D, Proton, Simulation3D, c, callback, e, epsilon_0, m_e, pi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nsteps", type=int, default=1001)
ap.add_argument("--timer", action="store_true", help="per-operator device times (serialises the step)")
args = ap.parse_args()

um = 1e-6
l0 = 0.8 * um
omega0 = 2 * pi * c / l0
nc = epsilon_0 * m_e * omega0**2 / e**2

nx, ny, nz = 512, 256, 256
dx, dy, dz = l0 / 20, l0 / 10, l0 / 10
Lx, Ly, Lz = nx * dx, ny * dy, nz * dz


def density(n0):
    def _density(x, y, z):
        if x > 1 * um:
            return n0
        return 0.0
    return _density


laser = GaussianLaser3D(a0=10, w0=2e-6, l0=0.8e-6, ctau=5e-6, focus_position=Lx / 2, x0=10e-6)
sim = Simulation3D(nx=nx, ny=ny, nz=nz, dx=dx, dy=dy, dz=dz, nsteps=args.nsteps, random_seed=3, store_part_fields=False,
                   enable_timer=args.timer)
ele = Electron(density=density(1 * nc), ppc=2)
proton = Proton(density=density(1 * nc), ppc=2)
sim.add_species([ele, proton])

history = []


@callback("end", interval=100, needs_host=False)
def energies(sim):
    history.append((sim.itime, sim.energies()))


@callback("end", interval=250, reads=("ey",), writes=())
def probe(sim):
    a0 = max(float(np.abs(p.fields.ey).max()) for p in sim.patches) * e / (m_e * c * omega0)
    print(f"step {sim.itime:5d}  max |a_y| = {a0:.3f}", flush=True)


if __name__ == "__main__":
    t0 = time.perf_counter()
    sim.initialize()
    npart = sum(int((~pt.is_dead).sum()) for p in sim.patches for pt in p.particles)
    t1 = time.perf_counter()
    print(f"initialised {sim.patches.npatches} patches, {npart} particles in {t1 - t0:.1f} s", flush=True)
    sim.run(callbacks=[laser, energies, probe])
    t2 = time.perf_counter()
    n_end = sum(int((~pt.is_dead).sum()) for p in sim.patches for pt in p.particles)
    steps = sim.itime
    if args.timer:
        from lambdapic_b200.simulation import Timer
        for name, sec in sorted(Timer.totals.items(), key=lambda kv: -kv[1])[:14]:
            print(f"  {name:44s} {1e3 * sec / steps:8.3f} ms/step", flush=True)
    st = sim.bridge.stats
    print(f"laser-target-3d {nx}x{ny}x{nz}, {sim.patches.npatches} patches, {npart} particles ({n_end} at the end): "
          f"{steps} steps in {t2 - t1:.2f} s = {steps / (t2 - t1):.2f} steps/s, {0.5 * (npart + n_end) * steps / (t2 - t1):.3e} particle-updates/s "
          f"(run() entry/exit copies included: {st['h2d_bytes'] / 1e9:.1f} GB up, {st['d2h_bytes'] / 1e9:.1f} GB down)")
    it, en = history[-1]
    print(f"energies at step {it}: " + ", ".join(f"{k} {v:.3e} J" for k, v in en.items()))
    sim.bridge.close()
