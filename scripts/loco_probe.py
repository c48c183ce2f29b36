"""C4 locomotion GA generation (pop 200 x (13,9) x seeds 1..5) and a pop-50 one, device time."""
import sys
import numpy as np
sys.path.insert(0, ".")
import evosops_b200 as E
import bench
ctx = E.Context()
for pop in (200, 50):
    gs = np.random.default_rng(5).integers(0, 11, size=(pop, 144), dtype=np.uint8)
    tsz, tsd = bench.trial_list([(13, 9)], [1, 2, 3, 4, 5])
    ori = (np.arange(pop * len(tsz)) % 3 + 1).astype(np.uint8)
    ctx.prepare(E.LOCO, gs, tsz, tsd, w1=0.75, w2=0.25, rng_mode=E.RNG_FAST, orient=ori)
    ctx.launch(); ctx.sync(); ctx.launch(); ms = ctx.sync(); r = ctx.collect(); c = ctx.counters()
    print("loco pop %3d: ms=%8.2f rate=%.3e p_poss=%.3f p_acc=%.4f fit=%.5f" % (pop, ms, c["attempts"] / ms * 1e3,
          c["possible"] / c["attempts"], c["committed"] / c["attempts"], r["norm"].mean()), flush=True)
