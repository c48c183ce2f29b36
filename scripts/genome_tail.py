"""Per-genome chain time on the C2 population: which genomes set the generation's makespan?"""
import sys
import numpy as np
sys.path.insert(0, ".")
import evosops_b200 as E
from bench import synthetic_genomes

ctx = E.Context()
gs = synthetic_genomes(0, 50, 48)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
rows = []
for g in range(50):
    ctx.prepare(E.AGG, gs[g:g + 1], [(13, 9)] * 5, [1, 2, 3, 4, 5], rng_mode=E.RNG_FAST, steps_override=steps)
    ctx.launch(); ctx.sync(); ctx.launch(); ms = ctx.sync(); ctx.collect()
    c = ctx.counters()
    rows.append((ms, c["possible"] / c["attempts"], c["committed"] / c["attempts"], g))
rows.sort()
for ms, pp, pa, g in rows:
    print("genome %2d  ms=%7.2f  cycles/step=%6.1f  p_poss=%.3f p_acc=%.4f" % (g, ms, ms * 1e-3 * 1.965e9 / steps, pp, pa))
print("mean ms %.2f  max ms %.2f" % (np.mean([r[0] for r in rows]), rows[-1][0]))
