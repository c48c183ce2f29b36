"""C3 / C4 generations (pop 200, random genomes, fast mode): device ms, best of 3."""
import sys
import numpy as np
sys.path.insert(0, ".")
import evosops_b200 as E
import bench
ctx = E.Context()
which = sys.argv[1:] or ["sep", "loco", "coat"]
cap = [int(w[4:]) for w in which if w.startswith("cap=")]
cap = cap[0] if cap else 0                      # cap=N: chains stop after N steps (quick ncu captures)
W = {E.SEP: (0.65, 0.35), E.COAT: (0.65, 0.35), E.LOCO: (0.75, 0.25)}
for name, beh, sizes, gseed in (("sep", E.SEP, [(13, 9)], 3), ("loco", E.LOCO, [(13, 9)], 4), ("coat", E.COAT, [(20, 5, 2), (26, 8, 4)], 5)):
    if name not in which:
        continue
    gb = np.random.default_rng(gseed).integers(0, 11, size=(200, E.GENOME_LEN[beh]), dtype=np.uint8)
    bsz, bsd = bench.trial_list(sizes, bench.SEEDS_C2)
    ori = (np.arange(200 * len(bsz)) % 3 + 1).astype(np.uint8) if beh == E.LOCO else None
    for mode in (E.RNG_FAST, E.RNG_VERIFY):
        if mode == E.RNG_VERIFY and "verify" not in which:
            continue
        ctx.prepare(beh, gb, bsz, bsd, w1=W[beh][0], w2=W[beh][1], rng_mode=mode, orient=ori, steps_override=cap)
        ctx.launch(); ctx.sync()
        best = min(ctx.launch() or ctx.sync() for _ in range(3))
        ctx.collect()
        c = ctx.counters()
        print("%-5s mode=%d pop200: ms=%8.2f rate=%.3e p_acc=%.4f" % (name, mode, best, c["attempts"] / best * 1e3, c["committed"] / c["attempts"]), flush=True)
