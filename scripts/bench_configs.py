"""Device-time measurements of BASELINE.json's other configs (C1, C3, C4, C5 sizes); bench.py covers C2.
Writes one JSON line per config.  Chains run their full n^3 steps unless a cap is stated."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import evosops_b200 as E
import bench

ctx = E.Context()
W = {E.AGG: (0.0, 0.0), E.SEP: (0.65, 0.35), E.COAT: (0.65, 0.35), E.LOCO: (0.75, 0.25)}


def run(name, beh, gs, sizes, seeds, steps=0, flags=0, orient=None, reps=2):
    tsz, tsd = bench.trial_list(sizes, seeds)
    P = gs.shape[0]
    out = {}
    for mode_name, mode in (("fast", E.RNG_FAST), ("verify", E.RNG_VERIFY)):
        ori = None if orient is None else np.tile(np.asarray(orient, dtype=np.uint8), (P * len(tsz) + len(orient) - 1) // len(orient))[:P * len(tsz)]
        ctx.prepare(beh, gs, tsz, tsd, w1=W[beh][0], w2=W[beh][1], rng_mode=mode, steps_override=steps, flags=flags, orient=ori)
        best = None
        for _ in range(reps):
            ctx.launch()
            ms = ctx.sync()
            best = ms if best is None else min(best, ms)
        res = ctx.collect()
        c = ctx.counters()
        out[mode_name] = {"ms_per_generation": best, "attempts_per_s": c["attempts"] / best * 1e3,
                          "p_poss": c["possible"] / c["attempts"], "p_acc": c["committed"] / c["attempts"],
                          "mean_fitness": float(res["norm"].mean())}
    line = {"config": name, "instances": P * len(tsz), "attempts": c["attempts"], "steps_cap": steps or None}
    line.update(out)
    print(json.dumps(line), flush=True)


def rnd(n, L, seed):
    return np.random.default_rng(seed).integers(0, 11, size=(n, L), dtype=np.uint8)


theory = np.asarray([[[n + j for _ in range(4)] for j in range(3)] for n in range(4)], dtype=np.uint8).reshape(1, 48)
run("C1 agg theory genome (6,3) seed 1 (theory table)", E.AGG, theory, [(6, 3)], [1], flags=E.F_THEORY_TABLE)
run("C3 sep GA generation pop200 x (13,9) x seeds1..5", E.SEP, rnd(200, 600, 3), [(13, 9)], [1, 2, 3, 4, 5])
run("C4 loco GA generation pop200 x (13,9) x seeds1..5", E.LOCO, rnd(200, 144, 4), [(13, 9)], [1, 2, 3, 4, 5], orient=[1, 2, 3])
run("C4 coat GA generation pop200 x (20,5,2),(26,8,4) x seeds1..5", E.COAT, rnd(200, 600, 5), [(20, 5, 2), (26, 8, 4)], [1, 2, 3, 4, 5])
evolved = np.asarray(bench.EVOLVED_AGG, dtype=np.uint8).reshape(1, 48)
for size in ((13, 9), (20, 14), (26, 18)):
    run("C5 sweep evolved agg genome x 20000 seeds x (%d,%d), 1e6-step cap" % size, E.AGG, evolved, [size], list(range(1, 20001)), steps=1000000, reps=1)
