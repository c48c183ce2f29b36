"""Print the hot SASS region of an ncu source-page CSV export (ncu -i X.ncu-rep --page source --csv --print-source sass):
address, samples, instructions executed, dominant stall reasons, instruction.  Usage: sass_hot.py file.csv [min_samples]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ci = {h: k for k, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ci["# Samples"]]) for r in rows[2:] if len(r) > ci["# Samples"])
print("total samples", tot)
for n, r in enumerate(rows[2:]):
    if len(r) <= ci["# Samples"]:
        continue
    smp = int(r[ci["# Samples"]])
    ex = int(r[ci["Instructions Executed"]])
    if smp < thr and ex == 0:
        continue
    st = sorted(((int(r[ci[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    sts = " ".join("%s:%d" % (b, a) for a, b in st if a)
    print("%4d %6d %5.2f%% ex=%9d thr=%4s  %-26s %s" % (n, smp, 100.0 * smp / max(tot, 1), ex, r[ci["Avg. Threads Executed"]], sts, r[ci["Source"]].strip()))
