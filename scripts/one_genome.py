"""Run one C2 genome x 5 seeds on (13,9) (latency-bound single-chain timing / ncu target)."""
import sys
sys.path.insert(0, ".")
import evosops_b200 as E
from bench import synthetic_genomes
g = int(sys.argv[1]) if len(sys.argv) > 1 else 11
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000000
ctx = E.Context()
gs = synthetic_genomes(0, 50, 48)
ctx.prepare(E.AGG, gs[g:g + 1], [(13, 9)] * 5, [1, 2, 3, 4, 5], rng_mode=E.RNG_FAST, steps_override=steps)
ctx.launch(); ms = ctx.sync(); ctx.collect()
c = ctx.counters()
print("genome %d ms=%.2f cycles/step=%.1f p_acc=%.4f" % (g, ms, ms * 1e-3 * 1.965e9 / steps, c["committed"] / c["attempts"]))
