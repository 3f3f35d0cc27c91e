"""Shared parity checks: compare a host-side state with a golden snapshot / another state.

Parity bar (BASELINE.json north_star): integer state bit-exact (is_dead, _id slot permutation,
capacities, sorter bucket tables, particle_index, nbuf); floating point <= FLOAT_RTOL of the max-abs of each array.
"""
import numpy as np

FIELD_ATTRS = ["ex", "ey", "ez", "bx", "by", "bz", "jx", "jy", "jz", "rho"]
PART_FLOAT_ATTRS = ["x", "y", "z", "w", "ux", "uy", "uz", "inv_gamma",
                    "ex_part", "ey_part", "ez_part", "bx_part", "by_part", "bz_part"]
FLOAT_RTOL = 1e-12  # single-step tolerance stated by north_star


def rel_err(got, ref):
    scale = float(np.abs(ref).max()) if ref.size else 0.0
    if scale == 0.0:
        scale = 1.0
    return float(np.abs(got - ref).max()) / scale if ref.size else 0.0


def check_state_against_golden(st, g, tag, rtol=FLOAT_RTOL, check_sorter=True, check_part_fields=True):
    """st: object with .patches[i].fields.<attr>, .patches[i].particles[s].<attr>/is_dead and optional .sorters."""
    nspec = int(g["meta/nspec"])
    dim = int(g["meta/dim"])
    worst = 0.0
    for ip, p in enumerate(st.patches):
        for a in FIELD_ATTRS:
            ref = g[f"{tag}/f/{ip}/{a}"]
            got = np.asarray(getattr(p.fields, a))
            assert got.shape == ref.shape, (a, got.shape, ref.shape)
            e = rel_err(got, ref)
            worst = max(worst, e)
            assert e <= rtol, f"field {a} patch {ip}: rel err {e:.3e}"
        for s in range(nspec):
            pt = p.particles[s]
            rd = g[f"{tag}/p/{ip}/{s}/is_dead"].astype(bool)
            gd = np.asarray(pt.is_dead).astype(bool)
            assert gd.shape == rd.shape, f"capacity patch {ip} spec {s}: {gd.shape} vs {rd.shape}"
            assert np.array_equal(gd, rd), f"is_dead patch {ip} spec {s}"
            rid = g[f"{tag}/p/{ip}/{s}/_id"].view(np.uint64)
            gid = np.asarray(pt._id).view(np.uint64)
            assert np.array_equal(gid, rid), f"_id (slot permutation) patch {ip} spec {s}"
            alive = ~rd
            for a in PART_FLOAT_ATTRS:
                if dim == 2 and a == "z":
                    continue
                if not check_part_fields and a.endswith("_part"):
                    continue
                ref = g[f"{tag}/p/{ip}/{s}/{a}"]
                got = np.asarray(getattr(pt, a))
                e = rel_err(got[alive], ref[alive])
                worst = max(worst, e)
                assert e <= rtol, f"particle {a} patch {ip} spec {s}: rel err {e:.3e}"
                if a in ("x", "y", "z"):
                    assert np.array_equal(np.isnan(got[rd]), np.isnan(ref[rd])), f"NaN pattern of dead {a}"
            if check_sorter and getattr(st, "sorters", None) is not None:
                srt = st.sorters[s]
                for nm, arr in (("bucket_count", srt.bucket_count), ("bucket_bound_min", srt.bound_min),
                                ("bucket_bound_max", srt.bound_max)):
                    assert np.array_equal(np.asarray(arr[ip]).ravel(), g[f"{tag}/s/{ip}/{s}/{nm}"].ravel()), nm
                rp = g[f"{tag}/s/{ip}/{s}/particle_index"]
                assert np.array_equal(np.asarray(srt.pidx[ip]), rp), "particle_index"
    if check_sorter and getattr(st, "sorters", None) is not None:
        for s in range(nspec):
            assert int(st.sorters[s].nbuf_last) == int(g[f"{tag}/nbuf/{s}"]), "nbuf"
    return worst
