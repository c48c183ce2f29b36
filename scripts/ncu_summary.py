"""Add one `ncu --page raw --csv` export (first kernel row) to profiles/r01_ncu_summary.json under a name.
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; python scripts/ncu_summary.py raw.csv <entry-name> [summary.json]"""
import csv
import json
import sys

KEEP = ("gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum")

raw, name = sys.argv[1], sys.argv[2]
out = sys.argv[3] if len(sys.argv) > 3 else "profiles/r01_ncu_summary.json"
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: [v, u] for h, u, v in zip(hdr, units, vals)}
entry = {"kernel": d.get("Kernel Name", ["?"])[0]}
entry.update({k: d[k] for k in KEEP if k in d})
try:
    summary = json.load(open(out))
except (OSError, ValueError):
    summary = {}
summary[name] = entry
json.dump(summary, open(out, "w"), indent=1)
print(name, json.dumps(entry)[:300])
