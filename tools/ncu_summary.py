#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into profiles/ (tracked).

    python tools/ncu_summary.py full  gpurun_out/eval.ncu-rep  profiles/r01_eval_ncu_full.txt [--traffic POINTS]
    python tools/ncu_summary.py list  gpurun_out/launches.csv  profiles/r01_launches.txt

`full`: key raw metrics of the first kernel in the report (+ opcode mix of the hot instructions from the source page);
with --traffic N also writes profiles/traffic.json (dram bytes per launch of N points) which bench.py reports as
roofline.traffic.  `list`: per-kernel launch counts / total time / share from a `--metrics gpu__time_duration.sum` CSV.
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.avg.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def full(rep, dst, traffic_points=None):
    rows = ncu_csv(rep, "raw")
    h, u, v = rows[0], rows[1], rows[2]
    col = {n: i for i, n in enumerate(h)}
    lines = [f"# {os.path.basename(rep)} (ncu --set full --clock-control none), first kernel of the report", f"kernel: {v[col['Kernel Name']]}"]
    for k in KEYS:
        if k in col:
            lines.append(f"{k} [{u[col[k]]}] {v[col[k]]}")
    for n, i in col.items():
        if "issue_stalled" in n and n.endswith("per_issue_active.ratio"):
            lines.append(f"{n} {v[i]}")
    rd = to_bytes(v[col["dram__bytes_read.sum"]], u[col["dram__bytes_read.sum"]])
    wr = to_bytes(v[col["dram__bytes_write.sum"]], u[col["dram__bytes_write.sum"]])
    lines.append(f"dram bytes per launch (read + write): {rd + wr:.0f}")
    src = ncu_csv(rep, "source")
    if len(src) > 2 and "Source" in src[1]:
        hh = src[1]
        si, ei = hh.index("Source"), hh.index("Instructions Executed")
        agg, tot = collections.Counter(), 0
        for r in src[2:]:
            if len(r) <= ei or not r[ei].isdigit():
                continue
            s = re.sub(r"^@!?U?P\d+\s+", "", r[si].strip())
            op = ".".join(s.split()[0].rstrip(";").split(".")[:2]) if s else ""
            agg[op] += int(r[ei])
            tot += int(r[ei])
        lines.append(f"warp instructions executed by opcode (total {tot}):")
        for op, n in agg.most_common(24):
            per = f"  {32.0 * n / traffic_points:6.2f} per 32 points" if traffic_points else ""
            lines.append(f"  {op:24s} {n:12d} {100.0 * n / tot:5.1f}%{per}")
    open(dst, "w").write("\n".join(lines) + "\n")
    if traffic_points:
        with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as fh:
            json.dump({"kernel": v[col["Kernel Name"]].split("(")[0], "points_per_launch": int(traffic_points), "dram_bytes_per_launch": rd + wr,
                       "dram_bytes_read": rd, "dram_bytes_write": wr, "source": os.path.basename(dst)}, fh, indent=1)
    print("\n".join(lines[:12]))


def launch_list(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    for i, r in enumerate(rows):
        if r[0] == "ID":
            hdr, rows = r, rows[i + 1:]
            break
    k, val, unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows:
        t = float(r[val].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[unit], 1e-3)
        name = re.sub(r"\(.*", "", r[k])[:110]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    lines = [f"# {os.path.basename(src)}: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: shares, not absolutes)",
             f"# {len(rows)} launches, {tot:.1f} us total", f"{'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}  kernel"]
    for name, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append(f"{c:8d} {t:12.1f} {t / c:10.1f} {100 * t / tot:6.1f}%  {name}")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:14]))


if __name__ == "__main__":
    if sys.argv[1] == "full":
        tp = int(sys.argv[sys.argv.index("--traffic") + 1]) if "--traffic" in sys.argv else None
        full(sys.argv[2], sys.argv[3], tp)
    else:
        launch_list(sys.argv[2], sys.argv[3])
