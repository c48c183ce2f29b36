"""C4 coating GA generation (pop 200 x (20,5,2),(26,8,4) x seeds 1..5) and a smaller one (pop 50), device time."""
import sys
import numpy as np
sys.path.insert(0, ".")
import evosops_b200 as E
import bench
ctx = E.Context()
for pop in (200, 50):
    gs = np.random.default_rng(3).integers(0, 11, size=(pop, 600), dtype=np.uint8)
    tsz, tsd = bench.trial_list([(20, 5, 2), (26, 8, 4)], [1, 2, 3, 4, 5])
    ctx.prepare(E.COAT, gs, tsz, tsd, w1=0.65, w2=0.35, rng_mode=E.RNG_FAST)
    ctx.launch(); ctx.sync(); ctx.launch(); ms = ctx.sync(); r = ctx.collect(); c = ctx.counters()
    print("coat pop %3d: ms=%8.2f rate=%.3e p_poss=%.3f p_acc=%.4f fit=%.5f" % (pop, ms, c["attempts"] / ms * 1e3,
          c["possible"] / c["attempts"], c["committed"] / c["attempts"], r["norm"].mean()), flush=True)
