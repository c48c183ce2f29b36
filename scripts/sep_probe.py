"""Separation: pop 50 x (13,9) x seeds 1..5, chains capped (default 2e6 steps) — device time; used for ncu captures."""
import sys
import numpy as np
sys.path.insert(0, ".")
import evosops_b200 as E
import bench
ctx = E.Context()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
gs = np.random.default_rng(11).integers(0, 11, size=(50, 600), dtype=np.uint8)
tsz, tsd = bench.trial_list([(13, 9)], [1, 2, 3, 4, 5])
ctx.prepare(E.SEP, gs, tsz, tsd, w1=0.65, w2=0.35, rng_mode=E.RNG_FAST, steps_override=steps)
ctx.launch(); ctx.sync(); ctx.launch(); ms = ctx.sync(); ctx.collect(); c = ctx.counters()
print("sep pop 50 x %d steps: ms=%8.2f rate=%.3e cycles/step(max chain)=%.1f p_poss=%.3f p_acc=%.4f" % (
    steps, ms, c["attempts"] / ms * 1e3, ms * 1e-3 * 1.965e9 / steps, c["possible"] / c["attempts"], c["committed"] / c["attempts"]))
