#!/usr/bin/env python
"""BASELINE.json configs[2] -- the reference's `example/lwfa.py` with the package import switched:

    500 x 800 cells, dx = dy = lambda/20, 10 x 10 patches, dt_cfl = 0.99, 100 fs (~1070 steps), underdense plasma
    (0.01 n_c: electrons 10 ppc, C6+ 1 ppc, protons 2 ppc), SimpleLaser2D a0 = 2, moving window at c from t = Lx/c.

Same constructor calls as the reference script; its HDF5 / plotting callbacks are replaced by a device-side energy
diagnostic and a read-only probe of `ey` and `rho` (`reads=` hint).  The MovingWindow recycles a column of patches about
every 71 steps once it starts (step ~714); all other steps cause no PCIe traffic.

    python examples/lwfa.py [--sim-time 100e-15]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lambdapic_b200 import (Electron, MovingWindow, Proton, SimpleLaser2D, Simulation, Species, c, callback, e, epsilon_0,  # noqa: E402
                            m_e, pi)

ap = argparse.ArgumentParser()
ap.add_argument("--sim-time", type=float, default=100e-15)
args = ap.parse_args()

um = 1e-6
l0 = 0.8 * um
omega0 = 2 * pi * c / l0
nc = epsilon_0 * m_e * omega0**2 / e**2

nx, ny = 500, 800
dx = dy = l0 / 20
Lx, Ly = nx * dx, ny * dy


def density(n0):
    def _density(x, y):
        ne = 0.0
        if x > 1 * um:
            ne = n0
        if abs(y - Ly / 2) > Ly / 2 - 1 * um:
            ne = 0
        return ne
    return _density


class TimedMovingWindow(MovingWindow):
    """MovingWindow that reports what a patch-recycling step costs (download, host re-load, upload)."""

    def __call__(self, sim):
        x0 = min(p.x0 for p in sim.patches)
        t = time.perf_counter()
        super().__call__(sim)
        if min(p.x0 for p in sim.patches) != x0:
            print(f"step {sim.itime:5d}  window shift took {time.perf_counter() - t:.2f} s", flush=True)


movingwindow = TimedMovingWindow(velocity=lambda t: c + (t - Lx / c) * 0)
laser = SimpleLaser2D(a0=2, w0=5e-6, l0=0.8e-6, ctau=5e-6)
ne = 0.01 * nc
sim = Simulation(nx=nx, ny=ny, dx=dx, dy=dy, npatch_x=10, npatch_y=10, dt_cfl=0.99, sim_time=args.sim_time, random_seed=2)
ele = Electron(density=density(ne), ppc=10)
proton = Proton(density=density(ne / 8 * 2), ppc=2)
carbon = Species(name="C", charge=6, mass=12 * 1800, density=density(ne / 8), ppc=1)
sim.add_species([ele, carbon, proton])

history = []


@callback("end", interval=100, needs_host=False)
def energies(sim):
    history.append((sim.itime, sim.energies()))


@callback("end", interval=10e-15, reads=("ey", "rho"), writes=())
def probe(sim):
    a0 = max(float(np.abs(p.fields.ey).max()) for p in sim.patches) * e / (m_e * c * omega0)
    x0 = min(p.x0 for p in sim.patches)
    print(f"step {sim.itime:5d}  t = {sim.time * 1e15:6.1f} fs  window xmin = {x0 * 1e6:6.2f} um  max |a_y| = {a0:.3f}  "
          f"shifts so far = {movingwindow.num_shifts}", flush=True)


if __name__ == "__main__":
    t0 = time.perf_counter()
    sim.initialize()
    npart = sum(int((~pt.is_dead).sum()) for p in sim.patches for pt in p.particles)
    t1 = time.perf_counter()
    sim.run(callbacks=[movingwindow, laser, energies, probe])
    t2 = time.perf_counter()
    n_end = sum(int((~pt.is_dead).sum()) for p in sim.patches for pt in p.particles)
    steps = sim.itime
    st = sim.bridge.stats
    print(f"lwfa {nx}x{ny}, {sim.patches.npatches} patches, {npart} particles ({n_end} at the end): initialize {t1 - t0:.1f} s, "
          f"{steps} steps in {t2 - t1:.2f} s = {steps / (t2 - t1):.1f} steps/s, {0.5 * (npart + n_end) * steps / (t2 - t1):.3e} particle-updates/s; "
          f"{st['downloads']} downloads / {st['uploads']} uploads, {st['d2h_bytes'] / 1e9:.2f} / {st['h2d_bytes'] / 1e9:.2f} GB")
    it, en = history[-1]
    print(f"energies at step {it}: " + ", ".join(f"{k} {v:.3e} J/m" for k, v in en.items()))
    sim.bridge.close()
