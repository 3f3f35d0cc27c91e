"""CPU test of the host-side laser source formulas (lambdapic_b200/laser.py) against the source planes the unmodified
reference produced (tests/golden/ref_laser_*.npz, written by oracle/make_golden.py:run_pml_case(with_laser=True))."""
import types

import numpy as np
import pytest

from lambdapic_b200.laser import GaussianLaser2D, GaussianLaser3D, SimpleLaser2D, SimpleLaser3D


def golden_laser(dim):
    l0 = 0.8e-6
    if dim == 3:
        return SimpleLaser3D(a0=2.0, w0=0.3e-6, ctau=0.05e-6, l0=l0, pol_angle=0.3, ellipticity=0.5, cep=0.7) + \
            GaussianLaser3D(a0=1.5, l0=l0, w0=0.4e-6, ctau=0.04e-6, x0=0.01e-6, focus_position=0.5e-6, pol_angle=-0.2,
                            ellipticity=-0.4, cep=0.2, l=1, p=1)
    return SimpleLaser2D(a0=2.0, w0=0.4e-6, ctau=0.05e-6, l0=l0, angle_y=0.15, pol_angle=0.3, ellipticity=0.5, cep=0.7) + \
        GaussianLaser2D(a0=1.5, l0=l0, w0=0.5e-6, ctau=0.04e-6, x0=0.01e-6, focus_position=0.5e-6, pol_angle=-0.2,
                        ellipticity=-0.4, cep=0.2)


def _axis(n, ng, d, o):
    ax = np.arange(n + 2 * ng, dtype=float)
    ax[-ng:] = np.arange(-ng, 0)
    return ax * d + o


@pytest.mark.parametrize("dim,case,nsteps", [(3, "golden_laser3d", 2), (2, "golden_laser2d", 3)])
def test_source_planes_match_reference(dim, case, nsteps, request):
    g = request.getfixturevalue(case)
    nx, ny, nz, ng = (int(g[f"meta/{k}"]) for k in ("nx", "ny", "nz", "n_guard"))
    dx, dy, dz, dt = (float(g[f"meta/{k}"]) for k in ("dx", "dy", "dz", "dt"))
    npy, npz = int(g["meta/npatch_y"]), int(g["meta/npatch_z"])
    sim = types.SimpleNamespace(dx=dx, dy=dy, dz=dz, Ly=ny * npy * dy, Lz=nz * npz * dz, cpml_thickness=int(g["meta/cpml_thickness"]), time=0.0)
    laser = golden_laser(dim)
    checked = 0
    for k in range(nsteps):
        sim.time = k * dt if k else 0.0
        sim.time = sum([dt] * k)  # the reference accumulates time += dt
        for ip in range(len(g["meta/x0"])):
            if f"src/{k}/{ip}/ey" not in g.files:
                continue
            f = types.SimpleNamespace()
            if dim == 3:
                f.yaxis = _axis(ny, ng, dy, g["meta/y0"][ip])[None, :, None]
                f.zaxis = _axis(nz, ng, dz, g["meta/z0"][ip])[None, None, :]
            else:
                f.yaxis = _axis(ny, ng, dy, g["meta/y0"][ip])[None, :]
            ey, ez = laser._calculate_bound_fields(sim, types.SimpleNamespace(fields=f))
            for got, ref in ((ey, g[f"src/{k}/{ip}/ey"]), (ez, g[f"src/{k}/{ip}/ez"])):
                got = np.broadcast_to(got, ref.shape)
                assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
                checked += 1
    assert checked >= 8


def test_parameter_validation_and_combination():
    with pytest.raises(ValueError):
        SimpleLaser2D(a0=-1, w0=1e-6, ctau=1e-6)
    with pytest.raises(NotImplementedError):
        SimpleLaser2D(a0=1, w0=1e-6, ctau=1e-6, angle_z=0.1)
    with pytest.raises(ValueError):
        GaussianLaser3D(a0=1, l0=0.8e-6, w0=1e-6, ctau=1e-6, p=-1)
    with pytest.raises(TypeError):
        SimpleLaser2D(a0=1, w0=1e-6, ctau=1e-6) + SimpleLaser3D(a0=1, w0=1e-6, ctau=1e-6)
    both = SimpleLaser2D(a0=1, w0=1e-6, ctau=1e-6) + GaussianLaser2D(a0=1, l0=0.8e-6, w0=1e-6, ctau=2e-6)
    assert both.stage == "_laser" and both.needs_host is False and both.tstop == 12e-6
