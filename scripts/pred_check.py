"""Host round-form predictor on generation-0 separation / locomotion / coating populations (pop 200, chains capped at 3e6 steps):
round 0 runs before the context has seen the behaviour (adaptive kernels), rounds 1-2 after a collected hot batch (parallel-rounds
kernels).  Device ms of the second launch of each round."""
import sys
import numpy as np
sys.path.insert(0, ".")
import evosops_b200 as E
import bench
ctx = E.Context()
W = {E.SEP: (0.65, 0.35), E.COAT: (0.65, 0.35), E.LOCO: (0.75, 0.25)}
for name, beh, sizes, gseed in (("sep", E.SEP, [(13, 9)], 3), ("loco", E.LOCO, [(13, 9)], 4), ("coat", E.COAT, [(20, 5, 2), (26, 8, 4)], 5)):
    gb = np.random.default_rng(gseed).integers(0, 11, size=(200, E.GENOME_LEN[beh]), dtype=np.uint8)
    bsz, bsd = bench.trial_list(sizes, bench.SEEDS_C2)
    ori = (np.arange(200 * len(bsz)) % 3 + 1).astype(np.uint8) if beh == E.LOCO else None
    for rnd in range(3):
        ctx.prepare(beh, gb, bsz, bsd, w1=W[beh][0], w2=W[beh][1], rng_mode=E.RNG_FAST, orient=ori, steps_override=3000000)
        ctx.launch(); ctx.sync(); ctx.launch(); ms = ctx.sync()
        ctx.collect()
        print("%-5s round %d (0: no predictor yet -> adaptive; then predictor): ms=%8.2f" % (name, rnd, ms), flush=True)
