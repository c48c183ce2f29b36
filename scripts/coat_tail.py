"""Per-genome chain time for coating (26,8,4): which genomes set a generation's makespan, and why."""
import sys
import numpy as np
sys.path.insert(0, ".")
import evosops_b200 as E
ctx = E.Context()
gs = np.random.default_rng(3).integers(0, 11, size=(50, 600), dtype=np.uint8)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
rows = []
for g in range(50):
    ctx.prepare(E.COAT, gs[g:g + 1], [(26, 8, 4)] * 5, [1, 2, 3, 4, 5], w1=0.65, w2=0.35, rng_mode=E.RNG_FAST, steps_override=steps)
    ctx.launch(); ctx.sync(); ctx.launch(); ms = ctx.sync(); ctx.collect()
    c = ctx.counters()
    rows.append((ms, c["possible"] / c["attempts"], c["committed"] / c["attempts"], g))
rows.sort()
for ms, pp, pa, g in rows[:3] + rows[-8:]:
    print("genome %2d  ms=%7.2f  cycles/step=%6.1f  p_poss=%.3f p_acc=%.4f" % (g, ms, ms * 1e-3 * 1.965e9 / steps, pp, pa))
print("mean ms %.2f  max ms %.2f" % (np.mean([r[0] for r in rows]), rows[-1][0]))
