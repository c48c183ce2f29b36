"""Per-genome chain time: which genomes set a generation's makespan?  One launch per genome (5 seeds, capped chains).
usage: genome_tail.py [agg|sep|coat|loco] [steps] [parallel|serial] [only this genome]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import evosops_b200 as E
from bench import synthetic_genomes

beh_name = sys.argv[1] if len(sys.argv) > 1 else "agg"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000000
form = sys.argv[3] if len(sys.argv) > 3 else "parallel"
beh = {"agg": E.AGG, "sep": E.SEP, "coat": E.COAT, "loco": E.LOCO}[beh_name]
fl = E.F_ROUNDS_PARALLEL if form == "parallel" else E.F_ROUNDS_SERIAL
ctx = E.Context()
if beh == E.AGG:
    gs = synthetic_genomes(0, 50, 48)
else:
    gs = np.random.default_rng({E.SEP: 3, E.COAT: 5, E.LOCO: 4}[beh]).integers(0, 11, size=(50, E.GENOME_LEN[beh]), dtype=np.uint8)
size = (20, 5, 2) if beh == E.COAT else (13, 9)
w = {E.AGG: (0, 0), E.SEP: (0.65, 0.35), E.COAT: (0.65, 0.35), E.LOCO: (0.75, 0.25)}[beh]
ori = np.asarray([1, 2, 3, 1, 2], dtype=np.uint8) if beh == E.LOCO else None
only = int(sys.argv[4]) if len(sys.argv) > 4 else None
rows = []
for g in (range(50) if only is None else [only]):
    ctx.prepare(beh, gs[g:g + 1], [size] * 5, [1, 2, 3, 4, 5], rng_mode=E.RNG_FAST, steps_override=steps, flags=fl, w1=w[0], w2=w[1], orient=ori)
    ctx.launch(); ctx.sync(); ctx.launch(); ms = ctx.sync(); ctx.collect()
    c = ctx.counters()
    rows.append((ms, c["possible"] / c["attempts"], c["committed"] / c["attempts"], g))
rows.sort()
for ms, pp, pa, g in rows:
    print("genome %2d  ms=%7.2f  cycles/step=%6.1f  p_poss=%.3f p_acc=%.4f" % (g, ms, ms * 1e-3 * 1.965e9 / (c["attempts"] / 5), pp, pa))
print("%s %s: mean ms %.2f  max ms %.2f" % (beh_name, form, np.mean([r[0] for r in rows]), rows[-1][0]))
