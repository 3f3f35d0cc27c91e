import os, sys, types, numpy as np
sys.path.insert(0, "/root/repo")
from tests.parity import FIELD_ATTRS, rel_err
g = np.load("/root/repo/tests/golden/ref_mw_3d.npz")
def run(mode):
    if mode == "sorted": os.environ["LPIC_PUSH_SORTED"] = "1"
    else: os.environ.pop("LPIC_PUSH_SORTED", None)
    from lambdapic_b200 import Electron, MovingWindow, Proton, Simulation3D, callback
    d, n0 = 0.8e-6 / 20, 1.742e27
    sim = Simulation3D(nx=24, ny=8, nz=8, dx=d, dy=d * 1.25, dz=d * 0.8, npatch_x=3, npatch_y=1, npatch_z=1, dt_cfl=0.95,
                       boundary_conditions=dict(xmin="pml", xmax="pml", ymin="periodic", ymax="periodic", zmin="periodic", zmax="periodic"),
                       cpml_thickness=6, random_seed=4321)
    dens = lambda x, y, z: n0 * (1.0 + x * 2.0e5)
    sim.add_species([Electron(density=dens, ppc=2), Proton(density=dens, ppc=1)])
    mw = MovingWindow(velocity=299792458.0, start_time=0.0)
    @callback("init")
    def seed(sim):
        rng = np.random.default_rng(6)
        for p in sim.patches:
            for isp, part in enumerate(p.particles):
                n = part.npart
                sig = 0.3 if isp == 0 else 0.02
                part.ux[:] = rng.normal(0.05 if isp == 0 else -0.01, sig, n)
                part.uy[:] = rng.normal(0.0, sig, n)
                part.uz[:] = rng.normal(0.0, sig, n)
                part.inv_gamma[:] = 1.0 / np.sqrt(1 + part.ux**2 + part.uy**2 + part.uz**2)
            f = p.fields
            for a, amp in (("ex", 3e11), ("ey", -2e11), ("ez", 1e11), ("bx", 500.0), ("by", -800.0), ("bz", 300.0)):
                arr = getattr(f, a)
                arr[...] = amp * rng.standard_normal(arr.shape)
    sim.initialize()
    sim.run(nsteps=1, callbacks=[seed, mw])
    for ip, p in enumerate(sim.patches):
        out = {}
        for a in ("jx", "jy", "jz", "rho", "ex"):
            ref = g[f"t1/f/{ip}/{a}"]; got = np.asarray(getattr(p.fields, a))
            i = np.unravel_index(np.abs(got - ref).argmax(), ref.shape)
            out[a] = f"{rel_err(got, ref):.1e} max {np.abs(ref).max():.2e} at {i} ref {ref[i]:.3e}"
        print(mode, ip, out)
    sim.bridge.close()
run("tile"); run("sorted")
