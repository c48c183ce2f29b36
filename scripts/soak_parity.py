"""Soak run of the GPU-vs-oracle parity check (final lattice, particle list, integers, counters) over many random cases,
biased towards high-acceptance genomes, every commit-round form.  Not part of the test-suite: a one-off
confidence run (usage: python scripts/soak_parity.py [cases] [seed])."""
import sys
import numpy as np
sys.path.insert(0, ".")
import evosops_b200 as E
from tests.test_gpu_parity import run_pair, check_equal

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 424242)
ctx = E.Context()
done = 0
for case in range(n_cases):
    beh = int(rng.choice([E.AGG, E.AGG, E.AGG, E.SEP, E.COAT, E.LOCO]))
    a = int(rng.integers(4, 40))
    if beh == E.COAT:
        o = int(rng.integers(1, max(2, a // 4)))
        c = int(rng.integers(1, max(2, (a - o) // 2)))
        size = (a, o, c)
        try:
            E.instance_shape(beh, size)
        except E.SopsError:
            continue
    else:
        size = (a, int(rng.integers(1, max(2, int(a * 0.75)))))
    hi = int(rng.choice([1, 2, 4, 11]))                      # gene range: small genes = high acceptance
    gs = rng.integers(0, hi, size=(2, E.GENOME_LEN[beh]), dtype=np.uint8)
    seeds = [int(rng.integers(0, 2 ** 63)), int(rng.integers(0, 2 ** 20))]
    ms = rng.integers(0, 2 ** 63, size=(2, 2), dtype=np.uint64)
    orient = rng.integers(1, 4, size=(2, 2)).astype(np.uint8) if beh == E.LOCO else None
    steps = int(rng.integers(1, 60000))
    rng_mode = int(rng.integers(0, 2))
    kflag = E.F_FORCE_CTA if (rng_mode == E.RNG_FAST and a <= 29 and rng.integers(0, 4) == 0) else 0
    kflag |= int(rng.choice([0, E.F_ROUNDS_SERIAL, E.F_ROUNDS_PARALLEL, E.F_ROUNDS_ADAPTIVE]))
    got, want, cnt = run_pair(ctx, beh, gs, [size, size], seeds, rng_mode, steps=steps, move_seeds=ms, orient=orient, flags=kflag)
    check_equal(ctx, beh, gs, [size, size], seeds, got, want, cnt, rng_mode, steps, move_seeds=ms, orient=orient, flags=kflag)
    done += 1
print("soak: %d cases bit-exact" % done)
