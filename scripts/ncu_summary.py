"""Summarise an ncu --set full capture exported with `--page raw --csv` and `--page source --csv`:
key metrics, stall reasons per issue, and hot code regions (runs of SASS instructions with the same execution count)."""
import collections
import csv
import sys

raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "smsp__warps_active.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]
for i, h in enumerate(hdr):
    if h in keys:
        print(f"{h} [{units[i]}] {vals[i]}")
st = [(h, vals[i]) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
print("stall cycles per issued instruction:")
for h, v in sorted(st, key=lambda x: -float(x[1].replace(",", "") or 0))[:9]:
    print("   ", h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
rows = list(csv.reader(open(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except Exception:
        return 0.0


tot_s = sum(f(r, "# Samples") for r in data)
tot_i = sum(f(r, "Instructions Executed") for r in data)
groups, cur = [], None
names = ("stall_wait", "stall_long_sb", "stall_short_sb", "stall_branch_resolving", "stall_no_inst", "stall_not_selected", "stall_selected",
         "stall_math", "stall_membar", "stall_sleep", "stall_dispatch", "stall_mio", "stall_lg")
for k, r in enumerate(data):
    ex = f(r, "Instructions Executed")
    if cur is None or abs(ex - cur["ex"]) > 0.15 * max(ex, cur["ex"], 1):
        cur = {"ex": ex, "start": k, "n": 0, "samples": 0, "inst": 0, "thr": 0, "st": collections.Counter()}
        groups.append(cur)
    cur["n"] += 1; cur["samples"] += f(r, "# Samples"); cur["inst"] += ex; cur["thr"] += f(r, "Thread Instructions Executed")
    for s in names:
        cur["st"][s] += f(r, s)
print(f"hot regions (of {len(data)} SASS instructions, {tot_i:.4g} executed):")
for g in groups:
    if g["samples"] > 0.01 * tot_s:
        print(f"  instr[{g['start']}:{g['start'] + g['n']}] n={g['n']} exec/instr={g['ex']:.3g} samples={100 * g['samples'] / tot_s:.1f}% "
              f"inst={100 * g['inst'] / tot_i:.1f}% lanes={g['thr'] / max(g['inst'], 1):.1f}",
              {k.replace('stall_', ''): round(100 * v / max(g['samples'], 1)) for k, v in g['st'].most_common(5)})
