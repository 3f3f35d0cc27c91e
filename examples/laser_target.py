#!/usr/bin/env python
"""BASELINE.json configs[1] -- the reference's `example/laser-target.py` with the package import switched:

    1024 x 1024 cells, dx = dy = lambda/50, CPML on all four sides, a 1 um slab of 10 n_c plasma (electrons, C6+, protons,
    10 particles per cell each, ~1.9e6 particles), Gaussian laser a0 = 10 from the xmin antenna, 2001 steps.

Same constructor calls as the reference script; its HDF5 / plotting callbacks (outside the accelerated path) are replaced
by `ExtractSpeciesDensity`, a device-side energy diagnostic and one read-only field probe that uses the `reads=` hint, so the
only per-step PCIe traffic is a handful of doubles.  Prints steps/s and particle-updates/s.

    python examples/laser_target.py [--nsteps 2001] [--nx 1024]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lambdapic_b200 import (Electron, ExtractSpeciesDensity, GaussianLaser2D, Proton, Simulation, Species, c, callback, e,  # noqa: E402
                            epsilon_0, m_e, pi)

ap = argparse.ArgumentParser()
ap.add_argument("--nsteps", type=int, default=2001)
ap.add_argument("--nx", type=int, default=1024)
args = ap.parse_args()

um = 1e-6
l0 = 0.8 * um
omega0 = 2 * pi * c / l0
nc = epsilon_0 * m_e * omega0**2 / e**2

nx = ny = args.nx
dx = dy = l0 / 50
Lx, Ly = nx * dx, ny * dy


def density(n0):
    def _density(x, y):
        ne = 0.0
        if x > Lx / 2 and x < Lx / 2 + 1 * um:
            ne = n0
        return ne
    return _density


laser = GaussianLaser2D(a0=10, w0=2e-6, l0=0.8e-6, ctau=5e-6, focus_position=Lx / 2, x0=10e-6, ellipticity=1)
sim = Simulation(nx=nx, ny=ny, dx=dx, dy=dy, nsteps=args.nsteps, random_seed=1)
ele = Electron(density=density(10 * nc), ppc=10)
proton = Proton(density=density(10 * nc / 8 * 2), ppc=10)
carbon = Species(name="C", charge=6, mass=12 * 1800, density=density(10 * nc / 8), ppc=10)
sim.add_species([ele, carbon, proton])

n_ele = ExtractSpeciesDensity(sim, ele, 500)  # as in the reference script; the guard reduce runs on the device, only rho is fetched
history = []


@callback("end", interval=100, needs_host=False)
def energies(sim):  # device-side reductions, no mirror traffic
    history.append((sim.itime, sim.energies()))


@callback("end", interval=500, reads=("ey",), writes=())
def probe(sim):  # what PlotFields / SaveFieldsToHDF5 would read: one field array crosses PCIe, nothing goes back
    a0 = max(float(np.abs(p.fields.ey).max()) for p in sim.patches) * e / (m_e * c * omega0)
    print(f"step {sim.itime:5d}  max |a_y| = {a0:.3f}  max n_e = {n_ele.density.max() / nc:.2f} n_c", flush=True)


if __name__ == "__main__":
    t0 = time.perf_counter()
    sim.initialize()
    npart = sum(int((~pt.is_dead).sum()) for p in sim.patches for pt in p.particles)
    t1 = time.perf_counter()
    sim.run(callbacks=[laser, n_ele, energies, probe])
    t2 = time.perf_counter()
    n_end = sum(int((~pt.is_dead).sum()) for p in sim.patches for pt in p.particles)
    steps = sim.itime
    print(f"laser-target {nx}x{ny}, {sim.patches.npatches} patches, {npart} particles ({n_end} at the end): initialize {t1 - t0:.1f} s, "
          f"{steps} steps in {t2 - t1:.2f} s = {steps / (t2 - t1):.1f} steps/s, "
          f"{0.5 * (npart + n_end) * steps / (t2 - t1):.3e} particle-updates/s, {nx * ny * steps / (t2 - t1):.3e} cell-updates/s")
    it, en = history[-1]
    print(f"energies at step {it}: " + ", ".join(f"{k} {v:.3e} J/m" for k, v in en.items()))
    sim.bridge.close()
