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


def split_engines_from_golden(g, tag, nranks, with_part=False, slack=1.5, min_extra=64):
    """One DeviceEngine per emulated rank (all on cuda:0), each owning its block of the golden state's patches."""
    from lambdapic_b200.workloads import make_patch_grid
    dim = int(g["meta/dim"])
    nx, ny, nz, ng = (int(g[f"meta/{k}"]) for k in ("nx", "ny", "nz", "n_guard"))
    dx, dy, dz = (float(g[f"meta/{k}"]) for k in ("dx", "dy", "dz"))
    npx, npy, npz = (int(g[f"meta/npatch_{a}"]) for a in "xyz")
    nspec = int(g["meta/nspec"])
    glob = g["meta/bounds_global"]
    engines, grids = [], []
    for r in range(nranks):
        pg = make_patch_grid(dim, npx, npy, npz, nx, ny, nz, dx, dy, dz, ng, (True, True, True), r, nranks)
        eng = DeviceEngine(dim, pg.npatch, nx, ny, nz, ng, dx, dy, dz, nspec)
        eng.set_geometry(pg.x0, pg.y0, pg.z0, pg.neighbor_ipatch, pg.boxes, pg.glob, r, pg.index)
        assert np.allclose(pg.glob, glob)
        for k, gp in enumerate(pg.index):
            for a in FIELD_ATTRS:
                eng.field_view(a, k)[...] = g[f"{tag}/f/{gp}/{a}"]
        for s in range(nspec):
            npart = [g[f"{tag}/p/{gp}/{s}/x"].size for gp in pg.index]
            m = eng.alloc_species(s, npart, slack=slack, min_extra=min_extra, with_part=with_part)
            for k, gp in enumerate(pg.index):
                for a in m.attrs:
                    m.view(a, k)[...] = g[f"{tag}/p/{gp}/{s}/{a}"]
                m.view("is_dead", k)[...] = g[f"{tag}/p/{gp}/{s}/is_dead"].astype(bool)
            eng.configure_sort(s, nx, 1, 1, dx, glob[3] - glob[2], (glob[5] - glob[4]) if dim == 3 else 1.0,
                               pg.x0 - dx / 2, pg.y0 - dy / 2, pg.z0 - dz / 2)
        eng.upload_all()
        engines.append(eng)
        grids.append(pg)
    meta = dict(dt=float(g["meta/dt"]), q=[float(v) for v in g["meta/q"]], m=[float(v) for v in g["meta/m"]], dim=dim)
    return engines, grids, meta


def compare_split_state_with_golden(engines, grids, g, tag, rtol):
    """N-rank result vs the 1-rank reference on the same global patch grid: fields <= rtol; per global patch the SET
    of alive particles (by _id) must be identical and their attributes agree to rtol (slot order may differ because
    arrivals from other ranks are placed before local ones)."""
    from tests.parity import rel_err
    worst = 0.0
    nspec = engines[0].nspec
    dim = engines[0].dim
    for eng, pg in zip(engines, grids):
        eng.download_all()
        for k, gp in enumerate(pg.index):
            for a in FIELD_ATTRS:
                e = rel_err(eng.field_view(a, k), g[f"{tag}/f/{gp}/{a}"])
                worst = max(worst, e)
                assert e <= rtol, f"field {a} global patch {gp}: {e:.3e}"
            for s in range(nspec):
                m = eng.species[s]
                alive = ~m.view("is_dead", k)
                ralive = ~g[f"{tag}/p/{gp}/{s}/is_dead"].astype(bool)
                ids = m.view("_id", k).view(np.uint64)[alive]
                rids = g[f"{tag}/p/{gp}/{s}/_id"].view(np.uint64)[ralive]
                assert ids.size == rids.size and np.array_equal(np.sort(ids), np.sort(rids)), f"particle set of patch {gp} spec {s}"
                o, ro = np.argsort(ids), np.argsort(rids)
                for a in ("x", "y", "z", "w", "ux", "uy", "uz", "inv_gamma"):
                    if dim == 2 and a == "z":
                        continue
                    e = rel_err(m.view(a, k)[alive][o], g[f"{tag}/p/{gp}/{s}/{a}"][ralive][ro])
                    worst = max(worst, e)
                    assert e <= rtol, f"particle {a} global patch {gp} spec {s}: {e:.3e}"
    return worst


def engine_from_pml_golden(g, tag, with_part=True):
    """Engine for the open-boundary golden cases (tests/golden/ref_pml_*.npz): neighbour table and shrunk particle boxes
    from the reference, CPML faces rebuilt with lambdapic_b200.pml from the recorded class names, psi loaded."""
    import types as _t
    from lambdapic_b200 import pml as pmlmod
    dim = int(g["meta/dim"])
    nx, ny, nz, ng = (int(g[f"meta/{k}"]) for k in ("nx", "ny", "nz", "n_guard"))
    dx, dy, dz = (float(g[f"meta/{k}"]) for k in ("dx", "dy", "dz"))
    x0, y0, z0 = g["meta/x0"], g["meta/y0"], g["meta/z0"]
    npatch, nspec = len(x0), int(g["meta/nspec"])
    eng = DeviceEngine(dim, npatch, nx, ny, nz, ng, dx, dy, dz, nspec)
    box = g["meta/boxes"].astype(float).copy()
    half = np.array([dx, dx, dy, dy, dz if dim == 3 else 0.0, dz if dim == 3 else 0.0]) / 2
    box += np.array([-1, 1, -1, 1, -1, 1]) * half
    glob = g["meta/bounds_global"]
    eng.set_geometry(x0, y0, z0, g["meta/neighbor_ipatch"], box, glob)
    cls2face = {"PMLXmin": "xmin", "PMLXmax": "xmax", "PMLYmin": "ymin", "PMLYmax": "ymax", "PMLZmin": "zmin", "PMLZmax": "zmax"}
    fields = _t.SimpleNamespace(nx=nx, ny=ny, dx=dx, dy=dy)
    if dim == 3:
        fields.nz, fields.dz = nz, dz
    inst, names = [], []
    nmax = max(nx, ny, nz if dim == 3 else 1)
    for ip in range(npatch):
        for slot, cname in enumerate(str(g["meta/pml_faces"][ip]).split(",")):
            if not cname:
                continue
            m = pmlmod.FACE_CLASS[cls2face[cname]](fields, thickness=int(g["meta/cpml_thickness"]))
            inst.append((ip, m.axis, slot, (m.efield_start, m.efield_end, m.bfield_start, m.bfield_end), m.profiles(nmax)))
            names.append((ip, slot, pmlmod.PSI_NAMES[m.axis]))
    eng.configure_pml(inst)
    for e, (ip, slot, nms) in enumerate(names):
        for r, nm in enumerate(nms):
            eng.psi_host[e, r] = g[f"{tag}/pml/{ip}/{slot}/{nm}"]
    eng.psi_names = names
    for ip in range(npatch):
        for a in FIELD_ATTRS:
            eng.field_view(a, ip)[...] = g[f"{tag}/f/{ip}/{a}"]
    for s in range(nspec):
        npart = [g[f"{tag}/p/{ip}/{s}/x"].size for ip in range(npatch)]
        m = eng.alloc_species(s, npart, slack=1.5, min_extra=64, with_part=with_part)
        for ip in range(npatch):
            for a in m.attrs:
                m.view(a, ip)[...] = g[f"{tag}/p/{ip}/{s}/{a}"]
            m.view("is_dead", ip)[...] = g[f"{tag}/p/{ip}/{s}/is_dead"].astype(bool)
        eng.configure_sort(s, nx, 1, 1, dx, glob[3] - glob[2], (glob[5] - glob[4]) if dim == 3 else 1.0,
                           x0 - dx / 2, y0 - dy / 2, z0 - dz / 2)
    eng.upload_all()
    meta = dict(dt=float(g["meta/dt"]), q=[float(v) for v in g["meta/q"]], m=[float(v) for v in g["meta/m"]], dim=dim)
    return eng, meta


class _Synthetic(dict):
    """A golden-format snapshot built on the fly (same keys as tests/golden/ref_step_*.npz)."""

    @property
    def files(self):
        return list(self.keys())


def synthetic_snapshot(wl, field_amp=1.0, seed=5):
    """A periodic thermal plasma of `wl` (lambdapic_b200.workloads.ThermalPlasma) as a golden-format snapshot `t0`: particles
    loaded cell by cell as the reference loader does, Maxwellian momenta, smooth random E/B including guards.  Feeds both
    `engine_from_golden` and `oracle.OState.from_golden`, so the CUDA path and the oracle start from identical bits."""
    from lambdapic_b200._lib import PART_ATTRS as PA
    pg = wl.grid()
    dim = wl.dim
    g = _Synthetic()
    for k, v in dict(dim=dim, nx=pg.nx, ny=pg.ny, nz=pg.nz, n_guard=pg.n_guard, dx=pg.dx, dy=pg.dy, dz=pg.dz if dim == 3 else 0.0,
                     dt=wl.dt, nspec=len(wl.ppc), npatch_x=pg.npx, npatch_y=pg.npy, npatch_z=pg.npz).items():
        g[f"meta/{k}"] = np.asarray(v)
    g["meta/q"], g["meta/m"] = np.array(wl.q), np.array(wl.m)
    g["meta/x0"], g["meta/y0"], g["meta/z0"] = pg.x0.astype(float), pg.y0.astype(float), pg.z0.astype(float)
    g["meta/neighbor_ipatch"] = pg.neighbor_ipatch
    g["meta/bounds_global"] = pg.glob
    rng = np.random.default_rng(seed)
    shape = (pg.nx + 2 * pg.n_guard, pg.ny + 2 * pg.n_guard) + ((pg.nz + 2 * pg.n_guard,) if dim == 3 else ())
    amps = dict(ex=3e10, ey=-2e10, ez=1e10, bx=50.0, by=-80.0, bz=30.0)
    idx = np.meshgrid(*[np.arange(n) for n in ((pg.nx, pg.ny, pg.nz) if dim == 3 else (pg.nx, pg.ny))], indexing="ij")
    for ip in range(pg.npatch):
        for a in FIELD_ATTRS:
            g[f"t0/f/{ip}/{a}"] = field_amp * amps[a] * rng.standard_normal(shape) if a in amps else np.zeros(shape)
        for s, ppc in enumerate(wl.ppc):
            n = idx[0].size * ppc
            d = {a: np.zeros(n) for a in PA}
            d["x"] = pg.x0[ip] + (np.repeat(idx[0].ravel(), ppc) + rng.random(n) - 0.5) * pg.dx
            d["y"] = pg.y0[ip] + (np.repeat(idx[1].ravel(), ppc) + rng.random(n) - 0.5) * pg.dy
            if dim == 3:
                d["z"] = pg.z0[ip] + (np.repeat(idx[2].ravel(), ppc) + rng.random(n) - 0.5) * pg.dz
            for a in ("ux", "uy", "uz"):
                d[a] = rng.normal(0.0, wl.uth[s], n)
            d["inv_gamma"] = 1.0 / np.sqrt(1 + d["ux"]**2 + d["uy"]**2 + d["uz"]**2)
            d["w"] = np.full(n, wl.weights[s])
            d["_id"] = ((np.uint64(ip) << np.uint64(32)) | np.arange(n, dtype=np.uint64)).view(np.float64)
            for a in PA:
                g[f"t0/p/{ip}/{s}/{a}"] = d[a]
            g[f"t0/p/{ip}/{s}/is_dead"] = np.zeros(n, dtype=bool)
    return g


def compare_with_oracle(eng, ost, nbuf=None, rtol=1e-12, part_fields=True):
    """Device state (through host_view) against an oracle OState: integers bit-exact, floats <= rtol of each array's max-abs."""
    from tests.parity import PART_FLOAT_ATTRS, rel_err
    st = host_view(eng, nbuf, with_sorter=False)
    worst = {}
    for ip, p in enumerate(ost.patches):
        for a in FIELD_ATTRS:
            e = rel_err(getattr(st.patches[ip].fields, a), getattr(p.fields, a))
            worst[a] = max(worst.get(a, 0.0), e)
            assert e <= rtol, f"field {a} patch {ip}: rel err {e:.3e}"
        for s in range(ost.nspec):
            ref, got = p.particles[s], st.patches[ip].particles[s]
            assert np.array_equal(np.asarray(got.is_dead).astype(bool), ref.is_dead), f"is_dead / capacity patch {ip} spec {s}"
            assert np.array_equal(np.asarray(got._id).view(np.uint64), ref._id.view(np.uint64)), f"slot permutation patch {ip} spec {s}"
            alive = ~ref.is_dead
            for a in PART_FLOAT_ATTRS:
                if (eng.dim == 2 and a == "z") or (a.endswith("_part") and (not part_fields or getattr(got, a) is None)):
                    continue
                e = rel_err(np.asarray(getattr(got, a))[alive], getattr(ref, a)[alive])
                worst[a] = max(worst.get(a, 0.0), e)
                assert e <= rtol, f"particle {a} patch {ip} spec {s}: rel err {e:.3e}"
    return worst
