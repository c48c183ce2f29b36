"""Late-generation population (mutants of the reference's evolved aggregation genome) on the C2 trial list."""
import sys
import numpy as np
sys.path.insert(0, ".")
import evosops_b200 as E
import bench
ctx = E.Context()
evo = np.tile(np.asarray(bench.EVOLVED_AGG, dtype=np.uint8), (50, 1))
rng = np.random.default_rng(1)
for g in range(1, 50):
    for _ in range(3):
        k = rng.integers(0, 48); evo[g, k] = np.clip(int(evo[g, k]) + rng.choice([-1, 1]), 0, 10)
tsz, tsd = bench.trial_list(bench.SIZES_C2, bench.SEEDS_C2)
ctx.prepare(E.AGG, evo, tsz, tsd, rng_mode=E.RNG_FAST)
ctx.launch(); ctx.sync(); ctx.launch(); ms = ctx.sync(); r = ctx.collect(); c = ctx.counters()
print("evolved pop: ms=%.2f rate=%.3e p_acc=%.4f fit=%.4f" % (ms, c["attempts"] / ms * 1e3, c["committed"] / c["attempts"], r["norm"].mean()))
