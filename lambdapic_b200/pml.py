"""CPML open-boundary layers: host-side description with the reference's class names and attributes
(core/boundary/cpml.py:11-340).  The coefficient profiles are computed here exactly as the reference does; the field
update and the psi auxiliary currents run on the GPU (csrc/fields.cu: k_update_*field CPML branch, k_pml_psi).
The psi_* arrays are views into the engine's pinned psi mirror, so callbacks and restart code see the reference's names."""
from __future__ import annotations

import numpy as np

C_LIGHT = 299792458.0
# reference psi attribute names per axis, in the device order E1, E2, B1, B2
PSI_NAMES = {0: ("psi_ey_x", "psi_ez_x", "psi_by_x", "psi_bz_x"),
             1: ("psi_ex_y", "psi_ez_y", "psi_bx_y", "psi_bz_y"),
             2: ("psi_ex_z", "psi_ey_z", "psi_bx_z", "psi_by_z")}


class PML:
    axis = 0
    side = "min"

    def __init__(self, fields, thickness: int = 6, kappa_max: float = 20.0, a_max: float = 0.15, sigma_max: float = 0.7):
        dims = (fields.nx, fields.ny) + ((fields.nz,) if hasattr(fields, "nz") else ())
        for n in dims:
            if n <= thickness:
                raise ValueError(f"PML thickness must be smaller than patch size. {thickness = }, patch = {dims}")
        self.fields = fields
        self.dimensions = dims
        self.thickness, self.kappa_max, self.a_max, self.sigma_max = thickness, kappa_max, a_max, sigma_max
        self.cpml_m, self.cpml_ma = 3, 1
        # the reference scales the conductivity of EVERY face with dx (cpml.py:60)
        self.sigma_maxval = sigma_max * C_LIGHT * 0.8 * (self.cpml_m + 1.0) / fields.dx
        n = dims[self.axis]
        self.n = n
        self.kappa_e, self.kappa_b = np.ones(n), np.ones(n)
        self.sigma_e, self.sigma_b = np.zeros(n), np.zeros(n)
        self.a_e, self.a_b = np.zeros(n), np.zeros(n)
        t = thickness
        if self.side == "min":  # cpml.py:247-262
            self._coeff(1.0 - np.arange(t, dtype=float) / t, np.s_[:t], "e")
            self._coeff(1.0 - (np.arange(t, dtype=float) + 0.5) / t, np.s_[:t], "b")
            self.efield_start, self.efield_end, self.bfield_start, self.bfield_end = 0, t, 0, t
        else:                   # cpml.py:265-282
            self._coeff(1.0 - np.arange(t, dtype=float)[::-1] / t, np.s_[n - t:n], "e")
            self._coeff(1.0 - (np.arange(t, dtype=float) + 0.5)[::-1] / t, np.s_[n - t - 1:n - 1], "b")
            self.efield_start, self.efield_end, self.bfield_start, self.bfield_end = n - t, n, n - t - 1, n - 1
        ax = "xyz"[self.axis]
        for w in ("e", "b"):  # reference attribute names: kappa_ex, sigma_bx, a_ey ...
            for q in ("kappa", "sigma", "a"):
                setattr(self, f"{q}_{w}{ax}", getattr(self, f"{q}_{w}"))
        for nm in PSI_NAMES[self.axis]:
            setattr(self, nm, np.zeros(dims))

    def _coeff(self, pos, sl, which):  # init_coefficents, cpml.py:118-126
        getattr(self, f"kappa_{which}")[sl] = 1 + (self.kappa_max - 1) * pos**self.cpml_m
        getattr(self, f"sigma_{which}")[sl] = self.sigma_maxval * pos**self.cpml_m
        getattr(self, f"a_{which}")[sl] = self.a_max * (1 - pos)**self.cpml_ma

    @property
    def face(self):
        return "xyz"[self.axis] + self.side

    def profiles(self, nmax):
        """(6, nmax): kappa_e, sigma_e, a_e, kappa_b, sigma_b, a_b padded with the neutral values."""
        out = np.zeros((6, nmax))
        out[0], out[3] = 1.0, 1.0
        for r, a in enumerate((self.kappa_e, self.sigma_e, self.a_e, self.kappa_b, self.sigma_b, self.a_b)):
            out[r, :self.n] = a
        return out


class PMLX(PML):
    axis = 0


class PMLY(PML):
    axis = 1


class PMLZ(PML):
    axis = 2


class PMLXmin(PMLX):
    side = "min"


class PMLXmax(PMLX):
    side = "max"


class PMLYmin(PMLY):
    side = "min"


class PMLYmax(PMLY):
    side = "max"


class PMLZmin(PMLZ):
    side = "min"


class PMLZmax(PMLZ):
    side = "max"


FACE_CLASS = {"xmin": PMLXmin, "xmax": PMLXmax, "ymin": PMLYmin, "ymax": PMLYmax, "zmin": PMLZmin, "zmax": PMLZmax}
