"""Quick device-only throughput probe (not the bench): attempts/s for a few batch shapes."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import evosops_b200 as E

ctx = E.Context()


def probe(name, beh, P, sizes, seeds, mode, steps=0, reps=2):
    rng = np.random.default_rng(abs(hash((beh, P))) % (2 ** 31))
    gs = rng.integers(0, 11, size=(P, E.GENOME_LEN[beh]), dtype=np.uint8)
    tsz = [s for s in sizes for _ in seeds]
    tsd = [sd for _ in sizes for sd in seeds]
    w = (0.65, 0.35)
    t0 = time.time()
    ctx.prepare(beh, gs, tsz, tsd, w1=w[0], w2=w[1], rng_mode=mode, steps_override=steps)
    best = None
    for _ in range(reps):
        ctx.launch()
        ms = ctx.sync()
        best = ms if best is None else min(best, ms)
    ctx.collect()
    c = ctx.counters()
    print("%-34s mode=%d inst=%6d attempts=%.3e  dev_ms=%9.2f  rate=%.3e /s  p_poss=%.3f p_acc=%.4f wall=%.1fs" % (
        name, mode, P * len(tsz), c["attempts"], best, c["attempts"] / best * 1e3,
        c["possible"] / c["attempts"], c["committed"] / c["attempts"], time.time() - t0), flush=True)


which = sys.argv[1:] or ["c2", "c5", "sep", "coat", "loco"]
modes = [m for m, nm in ((1, "fast"), (0, "verify")) if nm in which] or [1, 0]
for mode in modes:
    if "c2" in which:
        probe("C2 agg pop50 x3 sizes x5 seeds", E.AGG, 50, [(6, 4), (10, 7), (13, 9)], [1, 2, 3, 4, 5], mode)
    if "c2s" in which:
        probe("agg (13,9) pop50 x5 seeds 2e6 steps", E.AGG, 50, [(13, 9)], [1, 2, 3, 4, 5], mode, steps=2000000, reps=1)
    if "c5s" in which:
        # bench.py's seed sweep in small: the reference's evolved genome, serial rounds (what the host picks after a
        # low-acceptance batch), one launch
        import bench
        ev = np.asarray(bench.EVOLVED_AGG, dtype=np.uint8).reshape(1, 48)
        ctx.prepare(E.AGG, ev, [(13, 9)] * 20000, list(range(1, 20001)), rng_mode=mode, steps_override=200000, flags=E.F_ROUNDS_SERIAL)
        ctx.launch(); ms = ctx.sync(); ctx.collect(); c = ctx.counters()
        print("sweep evolved 20000 x (13,9) x 2e5 mode=%d ms=%.2f rate=%.3e p_poss=%.3f p_acc=%.4f" % (
            mode, ms, c["attempts"] / ms * 1e3, c["possible"] / c["attempts"], c["committed"] / c["attempts"]), flush=True)
    if "evo" in which:
        # bench.py's late-generation population (+-1 mutants of the reference's evolved genome), the C2 batch shape, default form
        import bench
        evo = bench.late_generation_population(0, 10)
        tsz, tsd = bench.trial_list(bench.SIZES_C2, bench.SEEDS_C2)
        ctx.prepare(E.AGG, evo, tsz, tsd, rng_mode=mode)
        for _ in range(3):
            ctx.launch(); ms = ctx.sync()
        ctx.collect(); c = ctx.counters()
        print("late-generation population pop50 C2 shape mode=%d ms=%.2f rate=%.3e p_poss=%.3f p_acc=%.4f" % (
            mode, ms, c["attempts"] / ms * 1e3, c["possible"] / c["attempts"], c["committed"] / c["attempts"]), flush=True)
    if "c5" in which:
        probe("agg (13,9) 20000 inst x 2e5 steps", E.AGG, 1, [(13, 9)], list(range(1, 20001)), mode, steps=200000)
        probe("agg (13,9) 100000 inst x 1e5 steps", E.AGG, 1, [(13, 9)], list(range(1, 100001)), mode, steps=100000)
        probe("agg (26,18) 20000 inst x 2e5 steps", E.AGG, 1, [(26, 18)], list(range(1, 20001)), mode, steps=200000)
    if "sep" in which:
        probe("sep (13,9) pop200 x5 seeds 2e6 steps", E.SEP, 200, [(13, 9)], [1, 2, 3, 4, 5], mode, steps=2000000)
    if "coat" in which:
        probe("coat (20,5,2) pop200 x5 full", E.COAT, 200, [(20, 5, 2)], [1, 2, 3, 4, 5], mode)
    if "loco" in which:
        probe("loco (13,9) pop200 x5 2e6 steps", E.LOCO, 200, [(13, 9)], [1, 2, 3, 4, 5], mode, steps=2000000)
