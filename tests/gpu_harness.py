"""Helpers shared by the GPU parity tests: golden snapshot -> DeviceEngine, engine -> host-state view."""
import types

import numpy as np

from lambdapic_b200._lib import FIELD_ATTRS, PART_ATTRS
from lambdapic_b200.engine import DeviceEngine


def boxes(x0, y0, z0, nx, ny, nz, dx, dy, dz):
    """Migration boxes widened by half a cell (core/patch/sync_particles_3d.c:402-411, patch.py:105-148 without PML)."""
    b = np.zeros((len(x0), 6))
    b[:, 0], b[:, 1] = x0 - 0.5 * dx, x0 + (nx - 1) * dx + 0.5 * dx
    b[:, 2], b[:, 3] = y0 - 0.5 * dy, y0 + (ny - 1) * dy + 0.5 * dy
    b[:, 4], b[:, 5] = z0 - 0.5 * dz, z0 + (nz - 1) * dz + 0.5 * dz
    return b


def engine_from_golden(g, tag, with_part=True, slack=1.5, min_extra=64):
    dim = int(g["meta/dim"])
    nx, ny, nz, ng = (int(g[f"meta/{k}"]) for k in ("nx", "ny", "nz", "n_guard"))
    dx, dy, dz = (float(g[f"meta/{k}"]) for k in ("dx", "dy", "dz"))
    x0, y0, z0 = g["meta/x0"], g["meta/y0"], g["meta/z0"]
    npatch, nspec = len(x0), int(g["meta/nspec"])
    eng = DeviceEngine(dim, npatch, nx, ny, nz, ng, dx, dy, dz, nspec)
    glob = g["meta/bounds_global"]
    eng.set_geometry(x0, y0, z0, g["meta/neighbor_ipatch"], boxes(x0, y0, z0, nx, ny, nz if dim == 3 else 1, dx, dy, dz if dim == 3 else 0.0), glob)
    for ip in range(npatch):
        for a in FIELD_ATTRS:
            eng.field_view(a, ip)[...] = g[f"{tag}/f/{ip}/{a}"]
    for s in range(nspec):
        npart = [g[f"{tag}/p/{ip}/{s}/x"].size for ip in range(npatch)]
        m = eng.alloc_species(s, npart, slack=slack, min_extra=min_extra, with_part=with_part)
        for ip in range(npatch):
            for a in m.attrs:
                m.view(a, ip)[...] = g[f"{tag}/p/{ip}/{s}/{a}"]
            m.view("is_dead", ip)[...] = g[f"{tag}/p/{ip}/{s}/is_dead"].astype(bool)
        eng.configure_sort(s, nx, 1, 1, dx, glob[3] - glob[2], (glob[5] - glob[4]) if dim == 3 else 1.0,
                           x0 - dx / 2, y0 - dy / 2, z0 - dz / 2)
    eng.upload_all()
    meta = dict(dt=float(g["meta/dt"]), q=[float(v) for v in g["meta/q"]], m=[float(v) for v in g["meta/m"]], dim=dim)
    return eng, meta


def host_view(eng, nbuf=None, with_sorter=True):
    """Download everything and expose it with the attribute names tests/parity.py expects."""
    eng.download_all()
    st = types.SimpleNamespace(patches=[], sorters=None)
    for ip in range(eng.npatch):
        f = types.SimpleNamespace(**{a: eng.field_view(a, ip) for a in FIELD_ATTRS})
        parts = []
        for s in range(eng.nspec):
            m = eng.species[s]
            d = {a: m.view(a, ip) for a in m.attrs}
            d["is_dead"] = m.view("is_dead", ip)
            for a in PART_ATTRS:
                d.setdefault(a, None)
            parts.append(types.SimpleNamespace(**d))
        st.patches.append(types.SimpleNamespace(fields=f, particles=parts))
    if with_sorter:
        st.sorters = []
        for s in range(eng.nspec):
            arr = eng.sort_arrays(s)
            st.sorters.append(types.SimpleNamespace(bucket_count=arr["bucket_count"], bound_min=arr["bound_min"],
                                                    bound_max=arr["bound_max"], pidx=arr["particle_index"],
                                                    nbuf_last=nbuf[s] if nbuf is not None else 0))
    return st
