"""Edge cases of the integer paths against the CPU oracle on the same inputs (bit-exact): empty and ragged patches,
all-dead arrays, fine bucket grids (shared-memory and global-memory histogram paths), open boundaries where leavers
have no neighbour, capacity growth at the array end."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pair(g, tag="t1"):
    from oracle import oracle as orc
    from tests import gpu_harness as h
    eng, meta = h.engine_from_golden(g, tag)
    ost = orc.OState.from_golden(g, tag)
    return eng, ost, meta, orc, h


def _assert_particles_equal(eng, ost, h, check_float=True):
    st = h.host_view(eng, with_sorter=False)
    for ip, p in enumerate(ost.patches):
        for s in range(ost.nspec):
            a, b = st.patches[ip].particles[s], p.particles[s]
            assert np.array_equal(a.is_dead, b.is_dead), (ip, s)
            assert np.array_equal(a._id.view(np.uint64), b._id.view(np.uint64)), (ip, s)
            if check_float:
                for at in ("x", "y", "z", "ux", "w"):
                    assert np.array_equal(np.asarray(getattr(a, at)).view(np.uint64), np.asarray(getattr(b, at)).view(np.uint64)), (ip, s, at)


@pytest.mark.parametrize("buckets", ["fine", "huge"])
def test_sort_with_fine_bucket_grids_bit_exact(golden3d, buckets):
    """ny/nz buckets > 1 (the collision configuration of the reference, particle_sort.py) through the shared-memory
    histogram ('fine': 5x4x6 = 120 buckets) and the global-memory fallback ('huge': 20x16x24 = 7680 > 2048)."""
    import ctypes
    eng, ost, meta, orc, h = _pair(golden3d)
    f = 1 if buckets == "fine" else 4
    nxb, nyb, nzb = eng.nx * f, eng.ny * f, eng.nz * f
    dxb, dyb, dzb = eng.dx / f, eng.dy / f, eng.dz / f
    x0s, y0s, z0s = eng.x0 - eng.dx / 2, eng.y0 - eng.dy / 2, eng.z0 - eng.dz / 2
    for s in range(eng.nspec):
        for rev in (False, True):
            eng.configure_sort(s, nxb, nyb, nzb, dxb, dyb, dzb, x0s, y0s, z0s)
            nbuf = eng.sort(s, rev)
            arr = eng.sort_arrays(s)
            ref_nbuf = 0
            for ip, p in enumerate(ost.patches):
                pt = p.particles[s]
                attrs = [getattr(pt, a) for a in orc.PART_ATTRS]
                n, nbin = pt.npart, nxb * nyb * nzb
                bc, bmin, bmax = (np.zeros(nbin, dtype=np.int64) for _ in range(3))
                pidx, pref, ptg = (np.zeros(n, dtype=np.int64) for _ in range(3))
                P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
                ref_nbuf += orc.lib().orc_sort_patch(
                    P(pt.x), P(pt.y), P(pt.z), P(pt.is_dead.view(np.uint8)), orc._pp(attrs), ctypes.c_int64(len(attrs)), ctypes.c_int64(n),
                    ctypes.c_int64(nxb), ctypes.c_int64(nyb), ctypes.c_int64(nzb), ctypes.c_double(dxb), ctypes.c_double(dyb), ctypes.c_double(dzb),
                    ctypes.c_double(x0s[ip]), ctypes.c_double(y0s[ip]), ctypes.c_double(z0s[ip]), ctypes.c_int(int(rev)),
                    P(bc), P(bmin), P(bmax), P(pidx), P(pref), P(ptg))
                assert np.array_equal(arr["bucket_count"][ip], bc) and np.array_equal(arr["bound_min"][ip], bmin)
                assert np.array_equal(arr["bound_max"][ip], bmax) and np.array_equal(arr["particle_index"][ip], pidx)
            assert nbuf == ref_nbuf
    _assert_particles_equal(eng, ost, h)
    eng.close()


def test_empty_ragged_and_all_dead_patches(golden3d):
    """Patch 0 has no slots at all for species 0, patch 1 is all dead, patch 2 is truncated: sort, push and migration
    must neither crash nor touch the others' results (reference tests: dead particle untouched / excluded)."""
    from oracle import oracle as orc
    from tests import gpu_harness as h
    from lambdapic_b200._lib import FIELD_ATTRS
    g = golden3d
    ost = orc.OState.from_golden(g, "t1")
    for a in orc.PART_ATTRS:
        setattr(ost.patches[0].particles[0], a, getattr(ost.patches[0].particles[0], a)[:0].copy())
        setattr(ost.patches[2].particles[0], a, getattr(ost.patches[2].particles[0], a)[:37].copy())
    ost.patches[0].particles[0].is_dead = ost.patches[0].particles[0].is_dead[:0].copy()
    ost.patches[0].particles[0].npart = 0
    ost.patches[2].particles[0].is_dead = ost.patches[2].particles[0].is_dead[:37].copy()
    ost.patches[2].particles[0].npart = 37
    ost.patches[1].particles[0].is_dead[:] = True
    ost.sorters = [orc.OSorter(ost, s) for s in range(ost.nspec)]
    rev = [bool(int(g[f"t1/reverse_x/{s}"])) for s in range(ost.nspec)]
    ost.set_reverse_x(rev)
    # same state on the device
    eng, meta = h.engine_from_golden(g, "t1")
    for s in range(eng.nspec):
        npart = [p.particles[s].npart for p in ost.patches]
        m = eng.alloc_species(s, npart, slack=1.5, min_extra=64, with_part=True,
                              npart_created=[g[f"t1/p/{ip}/{s}/x"].size for ip in range(eng.npatch)])
        for ip, p in enumerate(ost.patches):
            for a in m.attrs:
                m.view(a, ip)[...] = getattr(p.particles[s], a)
            m.view("is_dead", ip)[...] = p.particles[s].is_dead
    eng.upload_all()
    for p in ost.patches:
        for s in range(ost.nspec):
            p.particles[s]._npart_created = g[f"t1/p/{p.index}/{s}/x"].size
    for _ in range(2):
        orc.step(ost, "port")
        eng.step(meta["dt"], meta["q"], meta["m"], rev, write_part=True)
    _assert_particles_equal(eng, ost, h, check_float=False)
    eng.download_fields()
    for ip, p in enumerate(ost.patches):
        for a in FIELD_ATTRS:
            ref = getattr(p.fields, a)
            assert np.abs(eng.field_view(a, ip) - ref).max() <= 1e-11 * max(np.abs(ref).max(), 1e-300), (ip, a)
    eng.close()


def test_open_boundaries_kill_leavers_without_neighbour(golden3d):
    """neighbor_ipatch = -1 on every xmax face (non-periodic edge): leavers through it are not transferred, just marked
    dead with NaN positions, and the guard cells there are left alone (core/patch/sync_particles_3d.c:326-346)."""
    from oracle import oracle as orc
    from tests import gpu_harness as h
    g = golden3d
    nbr = g["meta/neighbor_ipatch"].copy()
    from lambdapic_b200.workloads import DIR3
    for b, d in enumerate(DIR3):
        if d[0] == 1:
            nbr[1::2, b] = -1   # patches with ipatch_x = 1 lose their +x neighbours
        if d[0] == -1:
            nbr[0::2, b] = -1   # and their partners lose the matching -x ones
    eng, meta = h.engine_from_golden(g, "t1")
    x0, y0, z0 = g["meta/x0"], g["meta/y0"], g["meta/z0"]
    eng.set_geometry(x0, y0, z0, nbr, h.boxes(x0, y0, z0, eng.nx, eng.ny, eng.nz, eng.dx, eng.dy, eng.dz), g["meta/bounds_global"])
    ost = orc.OState.from_golden(g, "t1")
    for ip, p in enumerate(ost.patches):
        p.neighbor_ipatch = np.ascontiguousarray(nbr[ip])
    rev = [bool(int(g[f"t1/reverse_x/{s}"])) for s in range(ost.nspec)]
    ost.set_reverse_x(rev)
    for _ in range(2):
        orc.step(ost, "port")
        eng.step(meta["dt"], meta["q"], meta["m"], rev, write_part=True)
    _assert_particles_equal(eng, ost, h, check_float=False)
    n_gpu = sum(eng.count_alive(s) for s in range(eng.nspec))
    assert n_gpu == ost.n_alive() and n_gpu < sum(int((~g[f"t1/p/{ip}/{s}/is_dead"].astype(bool)).sum()) for ip in range(8) for s in range(2))
    eng.close()
