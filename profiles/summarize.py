#!/usr/bin/env python
"""Turn ncu outputs in gpurun_out/ into the text summaries committed under profiles/.
  python profiles/summarize.py launches gpurun_out/launches.csv "<command line>" > profiles/<name>.txt
  python profiles/summarize.py full gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import collections
import csv
import re
import subprocess
import sys


def launches(path, cmd):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void <unnamed>::", "").replace("<unnamed>::", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e6 if r[ui] == "ns" else (v / 1e3 if r[ui] in ("us", "usecond") else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none ; command: {cmd}")
    print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
    print(f"{'kernel':50s} {'launches':>8s} {'total_ms':>10s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:50s} {n:8d} {t:10.3f} {100 * t / tot:6.2f}%")
    print(f"{'TOTAL':50s} {sum(a[0] for a in agg.values()):8d} {tot:10.3f}")


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__inst_executed_op_global_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
        "smsp__inst_executed_op_global_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "smsp__cycles_active.avg",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:140])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:75s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
        st = [(float(r[i]), h.split("stalled_")[1].split("_per_issue")[0]) for i, h in enumerate(hdr)
              if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")]
        print("  warp stall cycles per issued instruction:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:7]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
    else:
        full(sys.argv[2])
