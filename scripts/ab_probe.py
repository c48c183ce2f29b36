"""A/B probe: C2 generation, evolved population, random- and evolved-genome sweeps (device time)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import evosops_b200 as E
import bench

ctx = E.Context()


def run(name, beh, gs, tsz, tsd, steps=0, flags=0, mode=E.RNG_FAST):
    ctx.prepare(beh, gs, tsz, tsd, rng_mode=mode, steps_override=steps, flags=flags)
    ctx.launch(); ctx.sync(); ctx.launch(); ms = ctx.sync(); ctx.collect(); c = ctx.counters()
    print("%-28s ms=%8.2f rate=%.3e p_poss=%.3f p_acc=%.4f" % (name, ms, c["attempts"] / ms * 1e3,
          c["possible"] / c["attempts"], c["committed"] / c["attempts"]), flush=True)


tsz, tsd = bench.trial_list(bench.SIZES_C2, bench.SEEDS_C2)
run("C2 random pop", E.AGG, bench.synthetic_genomes(0, 50, 48), tsz, tsd)
evo = np.tile(np.asarray(bench.EVOLVED_AGG, dtype=np.uint8), (50, 1))
rng = np.random.default_rng(1)
for g in range(1, 50):
    for _ in range(3):
        k = rng.integers(0, 48); evo[g, k] = np.clip(int(evo[g, k]) + rng.choice([-1, 1]), 0, 10)
run("C2 evolved pop (warp)", E.AGG, evo, tsz, tsd)
run("C2 evolved pop (cta4)", E.AGG, evo, tsz, tsd, flags=E.F_FORCE_CTA if hasattr(E, "F_FORCE_CTA") else 0)
seeds = list(range(1, 20001))
run("sweep random genome", E.AGG, bench.synthetic_genomes(7, 1, 48), [(13, 9)] * len(seeds), seeds, steps=200000)
run("sweep evolved genome", E.AGG, evo[:1], [(13, 9)] * len(seeds), seeds, steps=200000)
run("sweep evolved (26,18)", E.AGG, evo[:1], [(26, 18)] * len(seeds), seeds, steps=200000)
if "verify" in sys.argv:
    run("C2 random pop verify", E.AGG, bench.synthetic_genomes(0, 50, 48), tsz, tsd, mode=E.RNG_VERIFY)
    run("sweep evolved verify", E.AGG, evo[:1], [(13, 9)] * len(seeds), seeds, steps=200000, mode=E.RNG_VERIFY)
