"""Compact table of the raw-page CSVs written by tools/gpu_ncu_rows.sh: one line per captured kernel launch."""
import csv
import sys

K = [('t_us', 'gpu__time_duration.sum'), ('rdMB', 'dram__bytes_read.sum'), ('wrMB', 'dram__bytes_write.sum'), ('dram%', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
     ('regs', 'launch__registers_per_thread'), ('grid', 'launch__grid_size'), ('issue%', 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
     ('alu%', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'), ('fma%', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'),
     ('fp64%', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'), ('xu%', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'),
     ('lsu%', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'), ('inst', 'smsp__inst_executed.sum'), ('L2hit%', 'lts__t_sector_hit_rate.pct'),
     ('long', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'), ('math', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio'),
     ('lg', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio'), ('short', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio'),
     ('wait', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio'), ('bar', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio'),
     ('membar', 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio'), ('mio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio')]


def f(v, unit):
    try:
        x = float(v.replace(',', ''))
    except ValueError:
        return v
    x *= {'Kbyte': 1e-3, 'Gbyte': 1e3, 'byte': 1e-6, 'ms': 1e3, 'ns': 1e-3}.get(unit, 1)
    return f"{x:.4g}"


for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    h, u = rows[0], rows[1]
    col = {n: i for i, n in enumerate(h)}
    seen = {}
    for r in rows[2:]:
        name = r[col['Kernel Name']].split('(')[0][:44]
        seen[name] = r  # keep the last (warm) launch of every kernel
    for name, r in seen.items():
        print(name.ljust(44), ' '.join(f"{n}={f(r[col[k]], u[col[k]])}" for n, k in K if k in col))
