"""A/B of the commit-round forms (SOPS_F_ROUNDS_*) on the shapes that matter: the C2 generation (latency bound), a
late-generation population, the worst C2 genome alone, and a seed sweep (throughput bound).  Device time, best of 2."""
import sys

import numpy as np

sys.path.insert(0, ".")
import evosops_b200 as E
import bench

ctx = E.Context()
FORMS = [("parallel", E.F_ROUNDS_PARALLEL), ("serial", E.F_ROUNDS_SERIAL), ("adaptive", E.F_ROUNDS_ADAPTIVE)]
which = sys.argv[1:] or ["c2", "evo", "worst", "sweep"]
only = [w[5:] for w in which if w.startswith("form=")]
if only:
    FORMS = [f for f in FORMS if f[0] in only]


def timed(beh, gs, tsz, tsd, flags, steps=0, mode=E.RNG_FAST, reps=2, w=(0.0, 0.0), orient=None):
    ctx.prepare(beh, gs, tsz, tsd, rng_mode=mode, steps_override=steps, flags=flags, w1=w[0], w2=w[1], orient=orient)
    best = None
    for _ in range(reps):
        ctx.launch()
        ms = ctx.sync()
        best = ms if best is None else min(best, ms)
    ctx.collect()
    c = ctx.counters()
    return best, c


def report(name, form, ms, c):
    print("%-44s %-9s ms=%9.2f  rate=%.3e/s  p_poss=%.3f p_acc=%.4f" % (
        name, form, ms, c["attempts"] / ms * 1e3, c["possible"] / c["attempts"], c["committed"] / c["attempts"]), flush=True)


tsz, tsd = bench.trial_list(bench.SIZES_C2, bench.SEEDS_C2)
gs = bench.synthetic_genomes(0, bench.POP_C2, 48)
for mode_name, mode in (("fast", E.RNG_FAST), ("verify", E.RNG_VERIFY)):
    if mode_name == "verify" and "verify" not in which:
        continue
    for form, fl in FORMS:
        if "c2" in which:
            ms, c = timed(E.AGG, gs, tsz, tsd, fl, mode=mode)
            report("C2 generation (%s)" % mode_name, form, ms, c)
        if "evo" in which:
            evo = np.tile(np.asarray(bench.EVOLVED_AGG, dtype=np.uint8), (50, 1))
            rng = np.random.default_rng(5)
            for g in range(1, 50):
                for _ in range(3):
                    k = int(rng.integers(0, 48))
                    evo[g, k] = min(10, max(0, int(evo[g, k]) + (1 if rng.integers(0, 2) else -1)))
            ms, c = timed(E.AGG, evo, tsz, tsd, fl, mode=mode)
            report("late-generation population (%s)" % mode_name, form, ms, c)
        if "worst" in which:
            ms, c = timed(E.AGG, gs[11:12], [(13, 9)] * 5, [1, 2, 3, 4, 5], fl, steps=2000000, mode=mode)
            report("genome 11 x5 seeds x 2e6 steps (%s)" % mode_name, form, ms, c)
            print("    cycles/step = %.1f" % (ms * 1e-3 * 1.965e9 / 2e6))
        if "sweep" in which:
            ev = np.asarray(bench.EVOLVED_AGG, dtype=np.uint8).reshape(1, 48)
            ms, c = timed(E.AGG, ev, [(13, 9)] * 20000, list(range(1, 20001)), fl, steps=200000, mode=mode)
            report("sweep 20000 x (13,9) x 2e5 evolved (%s)" % mode_name, form, ms, c)
            ms, c = timed(E.AGG, gs[3:4], [(13, 9)] * 20000, list(range(1, 20001)), fl, steps=200000, mode=mode)
            report("sweep 20000 x (13,9) x 2e5 random (%s)" % mode_name, form, ms, c)
if "coat" in which:
    rng = np.random.default_rng(5)
    gc = rng.integers(0, 11, size=(50, 600), dtype=np.uint8)
    csz, csd = bench.trial_list([(20, 5, 2), (26, 8, 4)], [1, 2, 3, 4, 5])
    for form, fl in (("parallel", E.F_ROUNDS_PARALLEL), ("serial", E.F_ROUNDS_SERIAL)):
        ms, c = timed(E.COAT, gc, csz, csd, fl, w=(0.65, 0.35))
        report("coat pop50 generation", form, ms, c)
if "sep" in which:
    rng = np.random.default_rng(3)
    gsep = rng.integers(0, 11, size=(50, 600), dtype=np.uint8)
    ssz, ssd = bench.trial_list([(13, 9)], [1, 2, 3, 4, 5])
    for form, fl in (("parallel", E.F_ROUNDS_PARALLEL), ("serial", E.F_ROUNDS_SERIAL)):
        ms, c = timed(E.SEP, gsep, ssz, ssd, fl, w=(0.65, 0.35))
        report("sep pop50 x (13,9) x5 generation", form, ms, c)
if "loco" in which:
    rng = np.random.default_rng(4)
    glo = rng.integers(0, 11, size=(50, 144), dtype=np.uint8)
    lsz, lsd = bench.trial_list([(13, 9)], [1, 2, 3, 4, 5])
    ori = (np.arange(250) % 3 + 1).astype(np.uint8)
    for form, fl in (("parallel", E.F_ROUNDS_PARALLEL), ("serial", E.F_ROUNDS_SERIAL)):
        ms, c = timed(E.LOCO, glo, lsz, lsd, fl, w=(0.75, 0.25), orient=ori)
        report("loco pop50 x (13,9) x5 generation", form, ms, c)
if "pop400" in which:
    big = bench.synthetic_genomes(0, 400, 48)
    for form, fl in FORMS:
        ms, c = timed(E.AGG, big, tsz, tsd, fl)
        report("agg pop400 generation (6000 chains)", form, ms, c)
