"""Add one `ncu --page raw --csv` export (first kernel row) to a profiles/*_ncu_summary.json under a name, with the
derived figures the bench line quotes: measured issue utilisation = smsp__issue_active (% of peak while the SM is active)
x sm__cycles_active / sm__cycles_elapsed, and warp / thread instructions per attempt when the attempt count is given.
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; python scripts/ncu_summary.py raw.csv <entry-name> [summary.json] [attempts] [note]"""
import csv
import json
import sys

KEEP = ("gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum")


def num(d, k):
    try:
        return float(d[k][0].replace(",", ""))
    except (KeyError, ValueError):
        return None


raw, name = sys.argv[1], sys.argv[2]
out = sys.argv[3] if len(sys.argv) > 3 else "profiles/r02_ncu_summary.json"
attempts = float(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4] not in ("", "0") else None
note = sys.argv[5] if len(sys.argv) > 5 else None
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: [v, u] for h, u, v in zip(hdr, units, vals)}
entry = {"kernel": d.get("Kernel Name", ["?"])[0]}
if note:
    entry["workload"] = note
entry.update({k: d[k] for k in KEEP if k in d})
ia, ca, ce = num(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"), num(d, "sm__cycles_active.avg"), num(d, "sm__cycles_elapsed.avg")
derived = {}
if None not in (ia, ca, ce) and ce:
    derived["issue_active_pct_while_active"] = ia
    derived["sm_active_over_elapsed"] = ca / ce
    derived["issue_utilisation_of_peak"] = ia / 100.0 * ca / ce
wi, tpi = num(d, "smsp__inst_executed.sum"), num(d, "smsp__thread_inst_executed_per_inst_executed.ratio")
wf, bc = num(d, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), num(d, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
if wf and bc is not None:
    derived["shared_bank_conflicts_over_wavefronts"] = bc / wf
rd, wr = num(d, "dram__bytes_read.sum"), num(d, "dram__bytes_write.sum")
if rd is not None and wr is not None:
    derived["dram_bytes"] = [rd, wr, d["dram__bytes_read.sum"][1]]
if attempts and wi:
    derived["attempts"] = attempts
    derived["warp_inst_per_attempt"] = wi / attempts
    if tpi:
        derived["thread_inst_per_attempt"] = wi * tpi / attempts
    t = num(d, "gpu__time_duration.sum")
    if t:
        scale = {"ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}.get(d["gpu__time_duration.sum"][1], 1e-3)
        derived["attempts_per_s_under_ncu"] = attempts / (t * scale)
entry["derived"] = derived
try:
    summary = json.load(open(out))
except (OSError, ValueError):
    summary = {}
summary[name] = entry
json.dump(summary, open(out, "w"), indent=1)
print(name, json.dumps(derived))
