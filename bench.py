#!/usr/bin/env python
"""bench.py -- particle-updates/s of the full PIC step (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path on the host cores

Workload (config.workload): BASELINE.json configs[4], "3D uniform thermal plasma weak/strong scaling sweep (512^3 cells,
8-64 ppc)": periodic electron-proton plasma, n = n_c(0.8 um), d = lambda/20, 1 keV, 16^3-cell patches, fp64.
  --scaling weak (default): 256^3 cells and 32+32 particles per cell PER GPU = 1.07e9 particles per GPU (SURVEY.md 8(d)-5's
                            weak shape; N = 8 is the full 512^3 box at 64 ppc)
  --scaling strong:         the 512^3 box at 4+4 ppc (1.07e9 particles in total) split over the N GPUs
configs[1..3] (laser / CPML / moving-window scripts) run through the public API in examples/ and are parity cases;
configs[0] is the reference's own CPU-sized test, used by the parity tests.  Defaults: 20 timed steps after 5 warm-up steps.
For N > 1 a multi-rank parity self-check runs before the timed region (key "multirank_parity").

One "step" = everything simulation/simulation.py:937-1130 does between stage `start` and stage `end` for the
periodic unified-pusher case: 4 FDTD half steps, 4 guard syncs, per-species sort, J/rho reset, fused
gather+Boris+Esirkepov per species, current reduce, per-species migration.

Prints ONE JSON line (see DESIGN.md "Measurement" for every key).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-updates/sec (full PIC step)"
UNIT = "particle-updates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cells", type=int, nargs=3, default=None, help="weak: cells per GPU (default 256^3); strong: global cells (default 512^3)")
    ap.add_argument("--ppc", type=int, nargs=2, default=None, help="default 32 32 (weak) / 4 4 (strong)")
    ap.add_argument("--patch", type=int, default=16)
    ap.add_argument("--temperature", type=float, default=1.0e3, help="eV (default 1 keV; 1e6 makes most particles change cell every step)")
    ap.add_argument("--slot-order", action="store_true", help="use the v1 particle kernel (memory order)")
    ap.add_argument("--breakdown", action="store_true", help="print per-operator CUDA-event times of one extra step to stderr")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-cells", type=int, default=64, help="edge of the CPU sample box (cells)")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the multi-rank parity self-check (N > 1)")
    ap.add_argument("--cell-sort", action="store_true", help="experiment: the reference's FULL-bucket sorter configuration (one bucket per "
                    "cell, what ParticleSort3D uses for species with collisions, simulation.py:1373) instead of the default x-column "
                    "buckets; the slot order then stays cell order, but it is not the order of the reference's default path")
    args = ap.parse_args()
    if args.ppc is None:
        args.ppc = [32, 32] if args.scaling == "weak" else [4, 4]
    return args


def workload(args, nranks):
    from lambdapic_b200.workloads import ThermalPlasma
    if args.scaling == "strong":  # fixed global box, block-partitioned over the ranks
        return ThermalPlasma(dim=3, cells=tuple(args.cells) if args.cells else (512, 512, 512), patch=(args.patch,) * 3, ppc=tuple(args.ppc),
                             temperature_eV=args.temperature)
    per_gpu = tuple(args.cells) if args.cells else (256, 256, 256)
    # weak scaling: the global box doubles along z, then y, then x as ranks double (8 ranks: 2x2x2 blocks)
    mult = [1, 1, 1]
    r, ax = nranks, 2
    while r > 1:
        mult[ax] *= 2
        ax = (ax - 1) % 3
        r //= 2
    cells = tuple(c * m for c, m in zip(per_gpu, mult))
    return ThermalPlasma(dim=3, cells=cells, patch=(args.patch,) * 3, ppc=tuple(args.ppc), temperature_eV=args.temperature)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ---------------------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own C extensions (oracle/_ref) or the oracle port, on the host cores
# ---------------------------------------------------------------------------------------------------------------
def cpu_state(wl_small):
    """Host-side OState with the same distribution as the device loader (uniform in cell, thermal momenta)."""
    from oracle import oracle as orc
    pg = wl_small.grid()
    st = orc.OState(3, pg.nx, pg.ny, pg.nz, pg.n_guard, pg.dx, pg.dy, pg.dz, wl_small.dt, wl_small.q, wl_small.m,
                    pg.x0, pg.y0, pg.z0, pg.neighbor_ipatch, pg.glob)
    rng = np.random.default_rng(wl_small.seed)
    ii, jj, kk = np.meshgrid(np.arange(pg.nx), np.arange(pg.ny), np.arange(pg.nz), indexing="ij")
    for ip, p in enumerate(st.patches):
        for s, ppc in enumerate(wl_small.ppc):
            n = pg.nx * pg.ny * pg.nz * ppc
            pt = p.particles[s]
            pt.x = pg.x0[ip] + (np.repeat(ii.ravel(), ppc) + rng.random(n) - 0.5) * pg.dx
            pt.y = pg.y0[ip] + (np.repeat(jj.ravel(), ppc) + rng.random(n) - 0.5) * pg.dy
            pt.z = pg.z0[ip] + (np.repeat(kk.ravel(), ppc) + rng.random(n) - 0.5) * pg.dz
            pt.ux, pt.uy, pt.uz = (rng.normal(0.0, wl_small.uth[s], n) for _ in range(3))
            pt.inv_gamma = 1.0 / np.sqrt(1 + pt.ux**2 + pt.uy**2 + pt.uz**2)
            pt.w = np.full(n, wl_small.weights[s])
            for a in orc.PART_ATTRS[8:14]:
                setattr(pt, a, np.zeros(n))
            pt.npart = n
            pt._id = pt._ids(0, n)
            pt._npart_created = n
            pt.is_dead = np.zeros(n, dtype=bool)
    st.sorters = [orc.OSorter(st, s) for s in range(st.nspec)]
    return st


def cpu_reference_run(args, steps, warmup):
    """Times `steps` full PIC steps of the reference's CPU path on a bounded sample of the workload."""
    from lambdapic_b200.workloads import ThermalPlasma
    from oracle import oracle as orc
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # torchrun exports OMP_NUM_THREADS=1 to every rank: override it (the reference arm uses all host threads it can), and
    # also tell an OpenMP runtime that was already initialised by an earlier import
    os.environ["OMP_NUM_THREADS"] = str(cores)
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(cores))
    except OSError:
        pass
    orc.lib()
    kind = "reference" if orc.have_ref() else "port"
    backend = "ref" if kind == "reference" else "port"
    e = args.cpu_cells
    wl_small = ThermalPlasma(dim=3, cells=(e, e, e), patch=(args.patch,) * 3, ppc=tuple(args.ppc))
    st = cpu_state(wl_small)
    for _ in range(max(warmup, 1)):
        orc.step(st, backend)
    t0 = time.perf_counter()
    n_upd = 0
    for _ in range(steps):
        n_upd += st.n_alive()
        orc.step(st, backend)
    dt = time.perf_counter() - t0
    sample = (f"{e}^3 cells, {args.ppc[0]}+{args.ppc[1]} ppc, {args.patch}^3 patches ({wl_small.n_particles()} particles), "
              f"{steps} steps after {max(warmup, 1)} warm-up; pusher/sort/sync = the reference's C extensions "
              f"(OpenMP, {cores} threads), FDTD = C restatement, OpenMP over patches" if kind == "reference" else
              f"{e}^3 cells, oracle C port, serial")
    return {"value": n_upd / dt, "unit": UNIT, "cores": cores if kind == "reference" else 1, "omp_threads": cores, "kind": kind,
            "sample": sample, "ms_per_step": 1e3 * dt / steps}


# ---------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference_run(args, args.steps, args.warmup)
    wl = workload(args, args.gpus)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, wl, args.gpus), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def config_dict(args, wl, nranks):
    return {"workload": f"BASELINE.json configs[4]: 3D uniform thermal e-/p+ plasma, periodic, {args.scaling} scaling "
                        f"({wl.cells[0]}x{wl.cells[1]}x{wl.cells[2]} cells on {nranks} GPU(s), "
                        f"{args.ppc[0]}+{args.ppc[1]} ppc, {args.patch}^3-cell patches, {args.temperature / 1e3:g} keV, fp64)",
            "cells_global": list(wl.cells), "particles_global": wl.n_particles(), "patch_cells": args.patch,
            "ppc": list(args.ppc), "n_guard": 3, "sorter": "one bucket per cell (--cell-sort experiment)" if args.cell_sort else "x-column buckets (reference default)", "parallelism": f"patch blocks over {nranks} GPU(s)",
            "l2_policy": "inputs larger than L2 (particle arenas are tens of GB; no flush needed)"}


def multirank_parity_check(args, rank, world, local, nsteps=3, rtol=1e-11):
    """N ranks against 1 rank on the same global patch grid, before the timed region: 64^3 cells in 16^3 patches, 2+2 ppc,
    20 keV (particles cross patch and rank boundaries every step).  Every rank runs `nsteps` steps of its block through
    the NCCL path; rank 0 also runs the whole box on its own GPU through the single-rank path.  Compared per global patch:
    E, B, J, rho (<= rtol of each array's max-abs), the SET of alive particle ids (exact) and their positions / momenta
    (<= rtol).  A mismatch makes every rank exit non-zero."""
    import torch
    import torch.distributed as dist
    from lambdapic_b200._lib import FIELD_ATTRS
    from lambdapic_b200.multigpu import HaloExchanger
    from lambdapic_b200.workloads import ThermalPlasma, build_engine
    wl = ThermalPlasma(dim=3, cells=(64, 64, 64), patch=(16, 16, 16), ppc=(2, 2), temperature_eV=2.0e4)
    rev = [False, False]

    def snapshot(eng):
        eng.download_all()
        out = {}
        for k, gp in enumerate(eng.grid.index):
            d = {a: eng.field_view(a, k).copy() for a in FIELD_ATTRS}
            for s in range(eng.nspec):
                m = eng.species[s]
                alive = ~m.view("is_dead", k)
                ids = m.view("_id", k).view(np.uint64)[alive] & np.uint64((1 << 50) - 1)  # drop the rank bits
                o = np.argsort(ids)
                d[f"ids{s}"] = ids[o]
                for a in ("x", "y", "z", "ux", "uy", "uz", "w"):
                    d[f"{a}{s}"] = m.view(a, k)[alive][o]
            out[int(gp)] = d
        return out

    eng = build_engine(wl, device=local, rank=rank, nranks=world)
    halo = HaloExchanger(eng, eng.grid)
    for _ in range(nsteps):
        halo.step(wl.dt, wl.q, wl.m, rev)
    mine = snapshot(eng)
    eng.close()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    result = [None]
    if rank == 0:
        ref_eng = build_engine(wl, device=local, rank=0, nranks=1)
        for _ in range(nsteps):
            ref_eng.step(wl.dt, wl.q, wl.m, rev)
        ref = snapshot(ref_eng)
        ref_eng.close()
        worst, bad, npatch = 0.0, [], 0
        for part in gathered:
            for gp, d in part.items():
                npatch += 1
                r = ref[gp]
                for key, got in d.items():
                    if key.startswith("ids"):
                        if got.size != r[key].size or not np.array_equal(got, r[key]):
                            bad.append(f"patch {gp} {key}: particle sets differ")
                        continue
                    if got.shape != r[key].shape:
                        continue  # reported through the ids
                    scale = float(np.abs(r[key]).max()) if r[key].size else 0.0
                    e = float(np.abs(got - r[key]).max()) / scale if scale > 0 else 0.0
                    worst = max(worst, e)
                    if e > rtol:
                        bad.append(f"patch {gp} {key}: rel err {e:.3e}")
        if npatch != len(ref):
            bad.append(f"{npatch} patches gathered, {len(ref)} expected")
        result[0] = {"ok": not bad, "max_rel": worst, "rtol": rtol, "steps": nsteps, "ranks": world,
                     "case": "64^3 cells, 16^3 patches, 2+2 ppc, 20 keV: N-rank NCCL path vs 1-rank path on rank 0's GPU",
                     "errors": bad[:5]}
    dist.broadcast_object_list(result, src=0)
    if not result[0]["ok"]:
        if rank == 0:
            print(json.dumps({"multirank_parity": result[0]}), flush=True)
        dist.destroy_process_group()
        sys.exit(3)
    return result[0]


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from lambdapic_b200 import _lib
    from lambdapic_b200.workloads import build_engine
    wl = workload(args, world)
    parity = None
    if world > 1 and not args.no_parity_check:
        parity = multirank_parity_check(args, rank, world, local)
    eng = build_engine(wl, device=local, rank=rank, nranks=world, slack=1.4 if args.scaling == "weak" else 1.3)
    eng.slot_order = args.slot_order
    if args.cell_sort:
        pg0 = eng.grid
        for sidx in range(eng.nspec):
            eng.configure_sort(sidx, pg0.nx, pg0.ny, pg0.nz, pg0.dx, pg0.dy, pg0.dz, pg0.x0 - pg0.dx / 2, pg0.y0 - pg0.dy / 2, pg0.z0 - pg0.dz / 2)
    if world > 1:
        from lambdapic_b200.multigpu import HaloExchanger
        eng.halo = HaloExchanger(eng, eng.grid)
    L = _lib.lib()
    dt, q, m = wl.dt, wl.q, wl.m
    rev = [False, False]  # thermal plasma has no net drift: the reference picks the normal x order

    def barrier():
        eng.sync()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()

    def one_step(slot=None):
        if world > 1:
            return eng.halo.step(dt, q, m, rev, slot)
        return eng.step(dt, q, m, rev, event_slot=slot)

    for _ in range(max(args.warmup, 3)):
        one_step()
    barrier()
    alive0 = sum(eng.count_alive(s) for s in range(eng.nspec))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.lpic_launch_count()
    barrier()
    L.lpic_event_record(eng.ctx, 0)
    t0 = time.perf_counter()
    for k in range(args.steps):
        one_step(slot=2 + 4 * k if k < 700 else None)  # events 2+4k .. 5+4k bracket the two push_deposit launches of step k
    L.lpic_event_record(eng.ctx, 1)
    ms = _lib.C.c_double(0)
    _lib.check(L.lpic_event_elapsed_ms(eng.ctx, 0, 1, _lib.C.byref(ms)))
    barrier()
    wall = time.perf_counter() - t0
    launches = L.lpic_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    # per-launch time of the dominant kernel (fused gather+push+deposit), CUDA events on its own stream
    push_ms = []
    for k in range(min(args.steps, 700)):
        for s in range(eng.nspec):
            e = _lib.C.c_double(0)
            _lib.check(L.lpic_event_elapsed_ms(eng.ctx, 2 + 4 * k + 2 * s, 3 + 4 * k + 2 * s, _lib.C.byref(e)))
            push_ms.append(e.value)
    alive1 = sum(eng.count_alive(s) for s in range(eng.nspec))
    if args.breakdown and world > 1 and not eng.halo.host_driven:  # per-phase CUDA-event times of one extra step, every rank
        names, slot = [], [3000]

        def mark(name):
            eng.record_event(slot[0])
            names.append(name)
            slot[0] += 1
        mark("begin")
        eng.halo.step(dt, q, m, rev, marks=mark)
        eng.sync()
        tms = []
        for i in range(1, len(names)):
            e = _lib.C.c_double(0)
            _lib.check(L.lpic_event_elapsed_ms(eng.ctx, 3000 + i - 1, 3000 + i, _lib.C.byref(e)))
            tms.append(e.value)
        tt = torch.tensor(tms, dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if rank == 0:
            for name, t in zip(names[1:], tt.tolist()):
                print(f"  {name:44s} {t:9.3f} ms (max over ranks)", file=sys.stderr)
            print(f"  {'TOTAL':44s} {sum(tt.tolist()):9.3f} ms", file=sys.stderr)
    if args.breakdown and world == 1:
        bd = eng.step_profiled(dt, q, m, rev)
        tot = sum(t for _, t in bd)
        for name, t in bd:
            print(f"  {name:28s} {t:9.3f} ms {100 * t / tot:5.1f}%", file=sys.stderr)
        print(f"  {'TOTAL':28s} {tot:9.3f} ms", file=sys.stderr)
    step_ms = ms.value / args.steps
    if world > 1:
        t = torch.tensor([step_ms, float(alive0), float(alive1), float(launches)], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        step_ms = float(tmax[0])
        alive0, alive1, launches = int(t[1]), int(t[2]), int(t[3])
    value = 0.5 * (alive0 + alive1) / (step_ms * 1e-3)

    # ---- roofline of the dominant kernel ----------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    pg = eng.grid
    n_local = 0.5 * (sum(eng.count_alive(s) for s in range(eng.nspec)) + alive0 / world) / eng.nspec  # per launch
    # algorithmic bytes of ONE fused launch: 121 B per particle (8 attrs + is_dead read, 7 written)
    # + per interior cell 6x8 B gather read + 4x16 B J/rho read-modify-write
    cells_local = pg.npatch * pg.nx * pg.ny * pg.nz
    bytes_launch = 121.0 * n_local + (48.0 + 64.0) * cells_local
    avg_push_ms = float(np.mean(push_ms))
    achieved = bytes_launch / (avg_push_ms * 1e-3) / 1e9
    fp64_peak = _lib.C.c_double(0)
    _lib.check(L.lpic_fp64_peak(eng.ctx, _lib.C.byref(fp64_peak)))
    FLOP_PER_UPDATE = 1.4e3  # gather 6x27 FMA, Boris, 27..125-point Esirkepov deposit
    fp64_achieved = FLOP_PER_UPDATE * n_local / (avg_push_ms * 1e-3) / 1e12
    roofline = {"bound": "hbm", "kernel": "lpic_push_deposit = k_tile_perm + k_push_tile<4,4,16> (E/B tile staged in shared memory, gather+Boris+Esirkepov, per-cell carried sums) + k_list_particles" if not args.slot_order else "k_particles<3,FUSED> (slot order)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6.65 TB/s",
                "avg_launch_ms": avg_push_ms, "algorithmic_bytes_per_launch": bytes_launch,
                "share_of_step": eng.nspec * avg_push_ms / (ms.value / args.steps),
                "per_species": [{"species": ["electron", "proton"][sidx] if eng.nspec == 2 else str(sidx),
                                 "avg_launch_ms": float(np.mean(push_ms[sidx::eng.nspec])),
                                 "frac": bytes_launch / (float(np.mean(push_ms[sidx::eng.nspec])) * 1e-3) / 1e9 / peak,
                                 "first_launch_ms": float(push_ms[sidx]), "last_launch_ms": float(push_ms[-eng.nspec + sidx])}
                                for sidx in range(eng.nspec)],
                "fp64": {"achieved_tflops": fp64_achieved, "peak_tflops": fp64_peak.value, "frac": fp64_achieved / max(fp64_peak.value, 1e-9),
                         "flop_per_update": FLOP_PER_UPDATE, "peak_source": "measured by lpic_fp64_peak (DFMA chains, this device, this run)"}}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        roofline["traffic"] = prof.get("bytes_per_particle") * n_local if prof.get("bytes_per_particle") else None
        roofline["traffic_source"] = prof.get("source")
    except Exception:
        pass

    # ---- end to end through the public API with host buffers (every rank takes part) ---------------------------
    eng.close()
    e2e = None
    if not args.no_e2e:
        try:
            e2e = e2e_public_api(args, wl, rank, world)
        except Exception as exc:  # keep the device-resident line even if the host-side leg cannot run on this box
            e2e = {"value": None, "unit": UNIT, "note": f"end-to-end leg failed: {type(exc).__name__}: {exc}"[:300]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args, wl, world),
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "wall_ms_per_step": 1e3 * wall / args.steps}

    if parity is not None:
        line["multirank_parity"] = parity
    if e2e is not None:
        line["e2e"] = e2e
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_reference_run(args, steps=10, warmup=1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def e2e_public_api(args, wl, rank=0, world=1):
    """Same metric through the call a user makes: ``Simulation3D(...).run(nsteps=K, callbacks=[diag])`` with the whole
    state starting and ending in pinned HOST memory.  The timed region contains the H2D copy of all fields and
    particles at run() entry, K full steps each followed by a device->host read of the step's energy diagnostic
    (a `needs_host=False` callback at stage `end`), and the D2H copy of all state at run() exit."""
    import lambdapic_b200 as lp
    import psutil
    from lambdapic_b200.workloads import ThermalPlasma
    # Host byte budget: the particles exist twice while initialize() seats them (loader arrays 73 B per particle, then the
    # pinned mirrors 73 B x 1.3 slack), plus sampler temporaries: ~200 B per particle and rank.  If this host cannot hold
    # that for the device-resident workload, the end-to-end leg runs the same plasma on a box halved along z (repeatedly),
    # and says so.
    def host_limit():
        lim = psutil.virtual_memory().available
        for path in ("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory/memory.limit_in_bytes"):
            try:
                v = open(path).read().strip()
                if v.isdigit():
                    lim = min(lim, int(v))
            except OSError:
                pass
        return lim
    full_cells = tuple(wl.cells)
    cells = list(wl.cells)
    ranks_here = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    while 200.0 * np.prod(cells) * sum(wl.ppc) / world * ranks_here > 0.7 * host_limit() and cells[2] > 2 * wl.patch[2] * (world if world > 1 else 1):
        cells[2] //= 2
    note = None
    if tuple(cells) != full_cells:
        note = (f"host memory budget: {200.0 * np.prod(full_cells) * sum(wl.ppc) / 1e9:.0f} GB needed for the mirrors of the "
                f"{full_cells[0]}x{full_cells[1]}x{full_cells[2]} box, {host_limit() / 1e9:.0f} GB usable on this host; end-to-end leg on "
                f"{cells[0]}x{cells[1]}x{cells[2]} cells, same ppc / patches / temperature")
        wl = ThermalPlasma(dim=3, cells=tuple(cells), patch=wl.patch, ppc=wl.ppc)
    if world > 1:
        import torch
        import torch.distributed as dist
    t_setup = time.perf_counter()
    per = {k: "periodic" for k in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")}
    npx, npy, npz = wl.npatches
    sim = lp.Simulation3D(nx=wl.cells[0], ny=wl.cells[1], nz=wl.cells[2], dx=wl.d, dy=wl.d, dz=wl.d, npatch_x=npx, npatch_y=npy,
                          npatch_z=npz, dt_cfl=wl.dt_cfl, boundary_conditions=per, random_seed=wl.seed, store_part_fields=False,
                          device=int(os.environ.get("LOCAL_RANK", "0")))
    sim.add_species([lp.Electron(density=wl.density, ppc=wl.ppc[0]), lp.Proton(density=wl.density, ppc=wl.ppc[1])])
    sim.initialize()
    # thermal momenta on the host mirrors: the reference's own stage-`init` callback (callback/utils.py:922-1049, ported in
    # lambdapic_b200.utils; same sampler, same draw order), applied here so that its host-side sampling stays outside the timed run()
    for sp in sim.species:
        lp.SetTemperature(sp, wl.temperature_eV)(sim)
    t_setup = time.perf_counter() - t_setup
    hist, stamps = [], []

    @lp.callback("end", needs_host=False)
    def diag(sim):
        hist.append(sim.energies())   # (a D2H read: the step's device work is complete when it returns)
        stamps.append(time.perf_counter())
    steps = max(1, args.steps)
    before = dict(sim.bridge.stats)
    n0 = sum(int((~pt.is_dead).sum()) for p in sim.patches for pt in p.particles)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
    t0 = time.perf_counter()
    sim.run(nsteps=steps, callbacks=[diag])
    dt = time.perf_counter() - t0
    n1 = sum(int((~pt.is_dead).sum()) for p in sim.patches for pt in p.particles)
    st = sim.bridge.stats
    h2d, d2h = st["h2d_bytes"] - before["h2d_bytes"], st["d2h_bytes"] - before["d2h_bytes"]
    if world > 1:  # whole-job figures: particles and bytes summed over the ranks, time = slowest rank
        v = torch.tensor([float(n0), float(n1), float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
        tm = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        n0, n1, h2d, d2h, dt = int(v[0]), int(v[1]), int(v[2]), int(v[3]), float(tm[0])
    tot = [sum(h.values()) for h in hist]
    out = {"value": 0.5 * (n0 + n1) * steps / dt, "unit": UNIT,
           "h2d_bytes_per_step": int(h2d / steps),
           "d2h_bytes_per_step": int(d2h / steps + 8 * len(hist[0]) * world),
           "steps": steps, "ms_per_step": 1e3 * dt / steps, "setup_s": t_setup,
           "h2d_seconds": st.get("h2d_seconds", 0.0) - before.get("h2d_seconds", 0.0),
           "d2h_seconds": st.get("d2h_seconds", 0.0) - before.get("d2h_seconds", 0.0),
           "energy_drift_rel": abs(tot[-1] - tot[0]) / tot[0],
           # wall time of every step inside run() (the first ones carry the reference's start-up transient: the first
           # migration appends 25 % dead slots to every patch and the following sorts drag up to half of all slots along)
           "step_ms": [round(1e3 * (b - a), 1) for a, b in zip([t0 + (st.get("h2d_seconds", 0.0) - before.get("h2d_seconds", 0.0))] + stamps[:-1], stamps)],
           "workload": f"{wl.cells[0]}x{wl.cells[1]}x{wl.cells[2]} cells, {wl.ppc[0]}+{wl.ppc[1]} ppc ({wl.n_particles()} particles)",
           "mode": "Simulation3D.run(nsteps=K): H2D of all state from pinned host mirrors at entry, K steps with a per-step "
                   "D2H energy diagnostic, D2H of all state at exit"}
    if note:
        out["note"] = note
    sim.bridge.close()
    return out


if __name__ == "__main__":
    main()
