import os, sys, numpy as np
sys.path.insert(0, "/root/repo")
from tests import gpu_harness as h
from tests.parity import FIELD_ATTRS, rel_err
g = np.load("/root/repo/tests/golden/ref_step_3d.npz")
for mode in ("tile", "sorted"):
    if mode == "sorted": os.environ["LPIC_PUSH_SORTED"] = "1"
    else: os.environ.pop("LPIC_PUSH_SORTED", None)
    for k in range(3):
        eng, meta = h.engine_from_golden(g, f"t{k}")
        rev = [bool(int(g[f"t1/reverse_x/{s}"])) for s in range(eng.nspec)]
        nbuf, mig = eng.step(meta["dt"], meta["q"], meta["m"], rev, write_part=True)
        st = h.host_view(eng, nbuf)
        worst = {}
        for ip, p in enumerate(st.patches):
            for a in FIELD_ATTRS:
                worst[a] = max(worst.get(a, 0), rel_err(getattr(p.fields, a), g[f"t{k+1}/f/{ip}/{a}"]))
            for s in range(eng.nspec):
                alive = ~g[f"t{k+1}/p/{ip}/{s}/is_dead"].astype(bool)
                for a in ("x", "ux", "ex_part", "bz_part"):
                    worst[a] = max(worst.get(a, 0), rel_err(getattr(p.particles[s], a)[alive], g[f"t{k+1}/p/{ip}/{s}/{a}"][alive]))
        print(mode, k, {a: f"{v:.1e}" for a, v in worst.items() if v > 0})
        eng.close()
