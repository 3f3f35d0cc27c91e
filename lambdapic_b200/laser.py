"""Laser antennas at the xmin boundary: same classes, constructor arguments and source-field formulas as the reference
(callback/laser.py:79-561).  The source planes (a few kB per edge patch) are evaluated on the host with numpy exactly
as the reference does; the rewrite of B at the antenna plane runs on the GPU (csrc/fields.cu:k_laser) -- the callback is
declared ``needs_host = False``, so the per-step stage `_laser` causes no field round trip over PCIe."""
from __future__ import annotations

import numpy as np

c, e, m_e, pi = 299792458.0, 1.602176634e-19, 9.1093837139e-31, 3.141592653589793


class Laser:
    DEFAULT_STAGE = "_laser"
    needs_host = False
    interval = 1

    def __init__(self) -> None:
        self.stage = self.DEFAULT_STAGE
        self.disabled = False
        self.side = "xmin"
        self.tstop = np.inf
        self.y0 = None
        self.z0 = None

    # subclasses ------------------------------------------------------------------------------------------------
    def _calculate_bound_fields(self, sim, patch):
        raise NotImplementedError

    def _ranges(self, patch):
        f = patch.fields
        r = [0, f.ny, 0, getattr(f, "nz", 1)]
        for m in patch.pml_boundary:  # transverse PML cells are left alone (laser.py:171-186, 218-241)
            if m.face == "ymin": r[0] = m.thickness
            if m.face == "ymax": r[1] = f.ny - m.thickness
            if m.face == "zmin": r[2] = m.thickness
            if m.face == "zmax": r[3] = f.nz - m.thickness
        return r

    def _stacked_edge(self, edge):
        """Virtual patch holding the transverse node coordinates of all antenna patches, stacked along y."""
        if not edge:
            return None
        f0 = edge[0][1].fields
        shape = tuple(f0.shape[1:])  # (NY,) or (NY, NZ)

        class _V:  # duck-typed like Patch / Fields for _calculate_bound_fields
            pass
        vp, vf = _V(), _V()
        vf.yaxis = np.concatenate([np.broadcast_to(p.fields.yaxis[0], shape) for _, p in edge])[None]
        if len(shape) == 2:
            vf.zaxis = np.concatenate([np.broadcast_to(p.fields.zaxis[0], shape) for _, p in edge])[None]
        vp.fields = vf
        vp.flat_shape = vf.yaxis.shape[1:]
        return vp, shape, [ip for ip, _ in edge], [self._ranges(p) for _, p in edge]

    def __call__(self, sim):
        """laser.py:109-137"""
        if self.disabled:
            return
        if c * sim.time >= self.tstop:
            self.disabled = True
            return
        if self.side != "xmin":
            raise ValueError("Invalid side: only 'xmin' is supported.")
        laserpos = sim.cpml_thickness + 2
        ps = sim.patches
        version = getattr(ps, "geometry_version", 0)  # bumped by MovingWindow when the columns rotate
        if getattr(self, "_edge_cache", (None, None))[:2] != (id(ps), version):
            edge = [(ip, p) for ip, p in enumerate(ps) if p.ipatch_x == 0]
            self._edge_cache = (id(ps), version, edge, self._stacked_edge(edge))
        edge, stacked = self._edge_cache[2], self._edge_cache[3]
        if sum(m.face == "xmin" for _, p in edge for m in p.pml_boundary) < len(edge):
            self.disabled = True  # no PML at xmin (e.g. a moving window has started)
            return
        if not edge:
            sim.mpi.comm.Barrier()
            return
        # ONE evaluation of the source formulas for all antenna patches: their transverse coordinates are stacked into a
        # single virtual patch (the formulas are elementwise), instead of ~10 small numpy calls per patch and step
        vpatch, shape, patches, ranges = stacked
        ey_s, ez_s = self._calculate_bound_fields(sim, vpatch)
        if ey_s is not None:
            full = (len(edge),) + shape
            eys = np.ascontiguousarray(np.broadcast_to(np.asarray(ey_s, dtype=np.float64), vpatch.flat_shape)).reshape(full)
            ezs = np.ascontiguousarray(np.broadcast_to(np.asarray(ez_s, dtype=np.float64), vpatch.flat_shape)).reshape(full)
            sim.bridge.laser_bfields(laserpos, patches, ranges, eys, ezs, sim.dt)
        sim.mpi.comm.Barrier()

    def __add__(self, other):
        if not isinstance(other, Laser):
            raise TypeError(f"Cannot add Laser with {type(other)}")
        if self.side != other.side:
            raise TypeError(f"Cannot add lasers from different sides: {self.side} and {other.side}")
        if isinstance(self, Laser2D) and isinstance(other, Laser2D):
            return _CombinedLaser2D(self, other)
        if isinstance(self, Laser3D) and isinstance(other, Laser3D):
            return _CombinedLaser3D(self, other)
        raise TypeError("Cannot add 2D and 3D laser")


class Laser2D(Laser):
    def _get_r(self, sim, patch):
        return abs(patch.fields.yaxis[0, :] - sim.dy / 2 - (self.y0 or sim.Ly / 2))

    def _get_phi(self, sim, patch):
        return np.arctan2(0.0, patch.fields.yaxis[0, :] - sim.dy / 2 - (self.y0 or sim.Ly / 2))

    def _get_boundary_coordinates(self, sim, patch):
        y = patch.fields.yaxis[0, :] - sim.dy / 2 - (self.y0 or sim.Ly / 2)
        return y, 0.0, abs(y)


class Laser3D(Laser):
    def _yz(self, sim, patch):
        f = patch.fields
        return (f.yaxis[0, :, :] - sim.dy / 2 - (self.y0 or sim.Ly / 2), f.zaxis[0, :, :] - sim.dz / 2 - (self.z0 or sim.Lz / 2))

    def _get_r(self, sim, patch):
        y, z = self._yz(sim, patch)
        return (y**2 + z**2)**0.5

    def _get_phi(self, sim, patch):
        y, z = self._yz(sim, patch)
        return np.arctan2(z, y)

    def _get_boundary_coordinates(self, sim, patch):
        y, z = self._yz(sim, patch)
        return y, z, np.sqrt(y**2 + z**2)


class _CombinedLaser(Laser):
    def __init__(self, laser1, laser2):
        super().__init__()
        self.laser1, self.laser2 = laser1, laser2
        self.side = laser1.side
        self.tstop = max(laser1.tstop, laser2.tstop)

    def _calculate_bound_fields(self, sim, patch):
        ey1, ez1 = self.laser1._calculate_bound_fields(sim, patch)
        ey2, ez2 = self.laser2._calculate_bound_fields(sim, patch)
        if ey1 is None and ey2 is None:
            return None, None
        if ey1 is None:
            return ey2, ez2
        if ey2 is None:
            return ey1, ez1
        return ey1 + ey2, ez1 + ez2


class _CombinedLaser2D(Laser2D, _CombinedLaser):
    pass


class _CombinedLaser3D(Laser3D, _CombinedLaser):
    pass


def _polarise(amp, phase, pol_angle, ellipticity):
    """Major/minor axis decomposition with cycle-averaged intensity conserved (laser.py:383-394, 541-552)."""
    norm = np.sqrt(1 + ellipticity**2)
    major, minor = 1.0 / norm, ellipticity / norm
    cos_pol, sin_pol = np.cos(pol_angle), np.sin(pol_angle)
    ey = amp * (major * cos_pol * np.sin(phase) - minor * sin_pol * np.cos(phase))
    ez = amp * (major * sin_pol * np.sin(phase) + minor * cos_pol * np.cos(phase))
    return ey, ez


class SimpleLaser(Laser):
    """Gaussian transverse profile, sin^2 temporal envelope, optional incidence angle (laser.py:267-396)."""

    def __init__(self, a0, w0, ctau, y0=None, z0=None, angle_y=0, angle_z=0, tstop=None, pol_angle=0.0, ellipticity=0.0,
                 cep=0.0, l0=0.8e-6, side="xmin"):
        super().__init__()
        if any(p <= 0 for p in [a0, l0, w0, ctau]):
            raise ValueError("All parameters (a0, l0, w0, ctau) must be positive")
        if side not in ["xmin"]:
            raise NotImplementedError("Invalid side: only 'xmin' is supported.")
        if abs(angle_y) >= pi / 2:
            raise ValueError("Angle_y must be in range (-pi/2, pi/2)")
        if angle_z != 0:
            raise NotImplementedError("Angle_z is not implemented")
        if abs(ellipticity) > 1:
            raise ValueError("Ellipticity must be in range [-1, 1]")
        self.a0, self.l0, self.w0, self.ctau = a0, l0, w0, ctau
        self.omega0 = 2 * pi * c / l0
        self.y0, self.z0, self.angle_y, self.angle_z = y0, z0, angle_y, angle_z
        self.tstop = 2 * ctau if tstop is None else c * tstop
        self.E0 = a0 * m_e * c * self.omega0 / e
        self.pol_angle, self.ellipticity, self.cep, self.side = pol_angle, ellipticity, cep, side
        self.k0 = self.omega0 / c
        self.ky = self.k0 * np.sin(self.angle_y)
        self.kz = 0

    def _calculate_bound_fields(self, sim, patch):
        time = sim.time
        if c * time >= self.tstop:
            return None, None
        y, z, r = self._get_boundary_coordinates(sim, patch)
        r_rot = np.sqrt((y / np.cos(self.angle_y))**2 + z**2)
        transverse_phase = -(self.ky * y + self.kz * z)
        t_rot = c * time - y * np.sin(self.angle_y)
        tprof = np.sin(t_rot / (2 * self.ctau) * pi)**2 * (t_rot < 2 * self.ctau)
        amp = self.E0 * np.exp(-r_rot**2 / self.w0**2) * tprof
        phase = self.omega0 * time + self.cep + transverse_phase
        ey, ez = _polarise(amp, phase, self.pol_angle, self.ellipticity)
        return ey * np.cos(self.angle_y), ez * np.cos(self.angle_z)


class SimpleLaser2D(Laser2D, SimpleLaser):
    pass


class SimpleLaser3D(Laser3D, SimpleLaser):
    pass


class GaussianLaser(Laser):
    """Paraxial Gaussian / Laguerre-Gaussian beam: waist evolution, Gouy phase, wavefront curvature (laser.py:405-554)."""

    def __init__(self, a0, l0, w0, ctau, x0=None, y0=None, z0=None, tstop=None, pol_angle=0.0, ellipticity=0.0, cep=0.0,
                 focus_position=0.0, side="xmin", l=0, p=0):  # noqa: E741
        super().__init__()
        if any(par <= 0 for par in [a0, l0, w0, ctau]):
            raise ValueError("All parameters (a0, l0, w0, ctau) must be positive")
        if side not in ["xmin"]:
            raise ValueError("Invalid side: only 'xmin' is implemented.")
        if abs(ellipticity) > 1:
            raise ValueError("Ellipticity must be in range [-1, 1]")
        if not isinstance(p, int) or p < 0:
            raise ValueError("Number of radial nodes p must be a non-negative integer")
        if not isinstance(l, int):
            raise ValueError("Azimuthal index l must be an integer")
        self.a0, self.l0, self.w0, self.ctau = a0, l0, w0, ctau
        self.omega0 = 2 * pi * c / l0
        self.k0 = self.omega0 / c
        self.x0 = 3 * ctau if x0 is None else x0
        self.y0, self.z0 = y0, z0
        self.tstop = 6 * ctau if tstop is None else c * tstop
        self.E0 = a0 * m_e * c * self.omega0 / e
        self.pol_angle, self.ellipticity, self.cep = pol_angle, ellipticity, cep
        self.focus_position, self.side = focus_position, side
        self.zR = pi * w0**2 / l0
        self._is_lg, self.l, self.p = False, l, p
        if l != 0 or p > 0:
            from scipy.special import factorial, genlaguerre
            self._is_lg = True
            self.lg_norm = np.sqrt(2 * factorial(p) / (pi * factorial(p + abs(l))))
            self.lg_norm /= np.sqrt(2 / pi)
            self.laguerre = genlaguerre(self.p, abs(self.l))

    def _gaussian_beam_params(self, z):
        z = z - self.focus_position
        w = self.w0 * np.sqrt(1 + (z / self.zR)**2)
        R = z * (1 + (self.zR / z)**2) if abs(z) > 1e-10 else np.inf
        return w, R, np.arctan(z / self.zR)

    def _calculate_bound_fields(self, sim, patch):
        time = sim.time
        if c * time >= self.tstop:
            return None, None
        tprof = np.exp(-(c * time - self.x0)**2 / self.ctau**2)
        x_rel = sim.cpml_thickness * sim.dx
        bw, bR, bpsi = self._gaussian_beam_params(x_rel)
        r = self._get_r(sim, patch)
        if self._is_lg:
            phi = self._get_phi(sim, patch)
            amp_lg = self.lg_norm * (np.sqrt(2) * r / bw)**abs(self.l) * self.laguerre((np.sqrt(2) * r / bw)**2)
            phase_lg = self.l * phi
        else:
            amp_lg, phase_lg = 1.0, 0.0
        amp = self.E0 * (self.w0 / bw) * np.exp(-r**2 / bw**2) * amp_lg
        phase_curv = self.k0 * r**2 / (2 * bR)
        phase = (self.omega0 * time + self.cep - self.k0 * x_rel - phase_curv - (2 * self.p + abs(self.l) + 1) * bpsi - phase_lg)
        return _polarise(amp * tprof, phase, self.pol_angle, self.ellipticity)


class GaussianLaser2D(Laser2D, GaussianLaser):
    pass


class GaussianLaser3D(Laser3D, GaussianLaser):
    pass
