"""Per-patch field container with the reference's attribute names and wrapped-guard layout
(core/fields.py:24-26,71-170).  The arrays are views into the DeviceEngine's pinned host mirror, so object
identity is kept across steps (callbacks may cache ``p.fields.ex``)."""
from __future__ import annotations

import numpy as np

from ._lib import FIELD_ATTRS


class Fields:
    attrs = list(FIELD_ATTRS)

    def __init__(self, engine, ipatch: int, dim: int, x0, y0, z0=0.0):
        self.nx, self.ny, self.n_guard = engine.nx, engine.ny, engine.ng
        self.dx, self.dy = engine.dx, engine.dy
        self.x0, self.y0 = x0, y0
        self.shape = engine.shape
        ng = self.n_guard
        for a in self.attrs:
            setattr(self, a, engine.field_view(a, ipatch))

        def axis(n, d, o):
            ax = np.arange(n + 2 * ng, dtype=float)
            if ng:
                ax[-ng:] = np.arange(-ng, 0)
            return ax * d + o
        if dim == 3:
            self.nz, self.dz, self.z0 = engine.nz, engine.dz, z0
            self.xaxis = axis(self.nx, self.dx, x0)[:, None, None]
            self.yaxis = axis(self.ny, self.dy, y0)[None, :, None]
            self.zaxis = axis(self.nz, self.dz, z0)[None, None, :]
        else:
            self.xaxis = axis(self.nx, self.dx, x0)[:, None]
            self.yaxis = axis(self.ny, self.dy, y0)[None, :]


class Fields2D(Fields):
    pass


class Fields3D(Fields):
    pass
