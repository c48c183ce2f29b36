"""A/B of kernel build variants (compile-time macros) in ONE GPU session.
usage: exp_ab.py [what ...] -- name=path/to/lib.so ...     (name "base" = the in-tree library)
Every variant runs in its own process (the library is loaded once per process): a short parity soak against the oracle,
then device times of the shapes named in `what`: worst (the slowest C2 genome), c2, evo / evo200 (late-generation
populations), g0 (generation-0 populations of 100 / 200), pop400, sweep, verify, vsweep, sep, sepevo, coat, loco, calm (calm
locomotion / coating populations), xover / xover2 / sepx (per-genome chain time under every round form, by acceptance rate).
Build the variants with scripts/exp_build.sh (exp/ is git-ignored; the libraries travel with gpurun)."""
import os
import subprocess
import sys

sys.path.insert(0, ".")


def child(lib, what):
    from evosops_b200 import _build, capi
    if lib != "base":
        _build.build = lambda *a, **k: lib                       # this process binds the variant library
    import numpy as np
    import evosops_b200 as E
    import bench
    from tests.test_gpu_parity import run_pair, check_equal
    ctx = E.Context()
    # parity soak (all behaviours, both modes and round forms, high-acceptance genomes included)
    rng = np.random.default_rng(777)
    done = 0
    for case in range(int(os.environ.get("EXP_SOAK", "40"))):
        beh = int(rng.choice([E.AGG, E.AGG, E.SEP, E.COAT, E.LOCO]))
        a = int(rng.integers(4, 30))
        if beh == E.COAT:
            o = int(rng.integers(1, max(2, a // 4))); c = int(rng.integers(1, max(2, (a - o) // 2)))
            size = (a, o, c)
            try:
                E.instance_shape(beh, size)
            except E.SopsError:
                continue
        else:
            size = (a, int(rng.integers(1, max(2, int(a * 0.75)))))
        hi = int(rng.choice([1, 2, 4, 11]))
        gs = rng.integers(0, hi, size=(2, E.GENOME_LEN[beh]), dtype=np.uint8)
        seeds = [int(rng.integers(0, 2 ** 63)), int(rng.integers(0, 2 ** 20))]
        ms = rng.integers(0, 2 ** 63, size=(2, 2), dtype=np.uint64)
        orient = rng.integers(1, 4, size=(2, 2)).astype(np.uint8) if beh == E.LOCO else None
        steps = int(rng.integers(1, 40000))
        rng_mode = int(rng.integers(0, 2))
        kflag = int(rng.choice([0, E.F_ROUNDS_SERIAL, E.F_ROUNDS_PARALLEL, E.F_ROUNDS_ADAPTIVE]))
        got, want, cnt = run_pair(ctx, beh, gs, [size, size], seeds, rng_mode, steps=steps, move_seeds=ms, orient=orient, flags=kflag)
        check_equal(ctx, beh, gs, [size, size], seeds, got, want, cnt, rng_mode, steps, move_seeds=ms, orient=orient, flags=kflag)
        done += 1
    print("  parity soak: %d cases bit-exact" % done, flush=True)

    def timed(beh, gs, tsz, tsd, flags=0, steps=0, mode=E.RNG_FAST, reps=2, w=(0.0, 0.0), orient=None):
        ctx.prepare(beh, gs, tsz, tsd, rng_mode=mode, steps_override=steps, flags=flags, w1=w[0], w2=w[1], orient=orient)
        best = None
        for _ in range(reps):
            ctx.launch()
            ms = ctx.sync()
            best = ms if best is None else min(best, ms)
        ctx.collect()
        return best, ctx.counters()

    def report(name, ms, c):
        print("  %-46s ms=%9.2f  rate=%.3e/s  p_poss=%.3f p_acc=%.4f" % (
            name, ms, c["attempts"] / ms * 1e3, c["possible"] / c["attempts"], c["committed"] / c["attempts"]), flush=True)

    tsz, tsd = bench.trial_list(bench.SIZES_C2, bench.SEEDS_C2)
    gs = bench.synthetic_genomes(0, bench.POP_C2, 48)
    evo = np.tile(np.asarray(bench.EVOLVED_AGG, dtype=np.uint8), (50, 1))
    r5 = np.random.default_rng(5)
    for g in range(1, 50):
        for _ in range(3):
            k = int(r5.integers(0, 48))
            evo[g, k] = min(10, max(0, int(evo[g, k]) + (1 if r5.integers(0, 2) else -1)))
    if "worst" in what:
        ms, c = timed(E.AGG, gs[11:12], [(13, 9)] * 5, [1, 2, 3, 4, 5], E.F_ROUNDS_PARALLEL, steps=2000000)
        report("genome 11 x5 seeds x 2e6 (parallel)", ms, c)
        print("      cycles/step = %.1f" % (ms * 1e-3 * 1.965e9 / 2e6))
    if "xover" in what:
        # per-genome chain time under both round forms: where is the crossover in acceptance rate?
        rows = []
        for gi in range(50):
            r = []
            for fl in (E.F_ROUNDS_PARALLEL, E.F_ROUNDS_SERIAL, E.F_ROUNDS_ADAPTIVE):
                ms, c = timed(E.AGG, gs[gi:gi + 1], [(13, 9)] * 5, [1, 2, 3, 4, 5], fl, steps=1000000)
                r.append(ms * 1e-3 * 1.965e9 / 1e6)
            rows.append((c["committed"] / c["attempts"], r[0], r[1], r[2], gi))
        rows.sort()
        for pa, cp, cs, ca, gi in rows:
            print("    genome %2d p_acc=%.4f  cycles/step parallel=%6.1f serial=%6.1f adaptive=%6.1f" % (gi, pa, cp, cs, ca), flush=True)
    if "c2" in what:
        for nm, fl in (("parallel", E.F_ROUNDS_PARALLEL), ("default", 0)):
            ms, c = timed(E.AGG, gs, tsz, tsd, fl, reps=3)
            report("C2 generation (%s)" % nm, ms, c)
    if "evo" in what:
        for nm, fl in (("serial", E.F_ROUNDS_SERIAL), ("parallel", E.F_ROUNDS_PARALLEL), ("default", 0)):
            ms, c = timed(E.AGG, evo, tsz, tsd, fl)
            report("late-generation population (%s)" % nm, ms, c)
    if "pop400" in what:
        for nm, fl in (("parallel", E.F_ROUNDS_PARALLEL), ("adaptive", E.F_ROUNDS_ADAPTIVE)):
            ms, c = timed(E.AGG, bench.synthetic_genomes(0, 400, 48), tsz, tsd, fl)
            report("agg pop400 generation (%s)" % nm, ms, c)
        big = np.tile(evo, (8, 1))
        for nm, fl in (("serial", E.F_ROUNDS_SERIAL), ("adaptive", E.F_ROUNDS_ADAPTIVE)):
            ms, c = timed(E.AGG, big, tsz, tsd, fl)
            report("late-generation pop400 (%s)" % nm, ms, c)
    if "sweep" in what:
        ev = np.asarray(bench.EVOLVED_AGG, dtype=np.uint8).reshape(1, 48)
        for nm, fl in (("serial", E.F_ROUNDS_SERIAL), ("parallel", E.F_ROUNDS_PARALLEL), ("adaptive", E.F_ROUNDS_ADAPTIVE)):
            ms, c = timed(E.AGG, ev, [(13, 9)] * 20000, list(range(1, 20001)), fl, steps=200000)
            report("sweep 20000 x (13,9) x 2e5 evolved (%s)" % nm, ms, c)
    if "evo200" in what:
        big = np.tile(evo, (4, 1))
        for nm, fl in (("serial", E.F_ROUNDS_SERIAL), ("adaptive", E.F_ROUNDS_ADAPTIVE)):
            ms, c = timed(E.AGG, big, tsz, tsd, fl)
            report("late-generation pop200 (%s)" % nm, ms, c)
        for nm, fl in (("serial", E.F_ROUNDS_SERIAL), ("adaptive", E.F_ROUNDS_ADAPTIVE)):
            ms, c = timed(E.AGG, big[:100], tsz, tsd, fl)
            report("late-generation pop100 (%s)" % nm, ms, c)
    if "calm" in what:
        # calm loco / coat populations (genes 6..10: rarely accepting), one GA generation's shape
        for lo in (5, 7):
            glo = np.random.default_rng(40 + lo).integers(lo, 11, size=(50, 144), dtype=np.uint8)
            lsz, lsd = bench.trial_list([(13, 9)], [1, 2, 3, 4, 5])
            ori = (np.arange(250) % 3 + 1).astype(np.uint8)
            for nm, fl in (("parallel", E.F_ROUNDS_PARALLEL), ("serial", E.F_ROUNDS_SERIAL), ("adaptive", E.F_ROUNDS_ADAPTIVE)):
                ms, c = timed(E.LOCO, glo, lsz, lsd, fl, w=(0.75, 0.25), orient=ori, steps=4000000)
                report("loco genes %d..10 pop50 (13,9) 4e6 (%s)" % (lo, nm), ms, c)
            gc = np.random.default_rng(50 + lo).integers(lo, 11, size=(50, 600), dtype=np.uint8)
            csz, csd = bench.trial_list([(20, 5, 2), (26, 8, 4)], [1, 2, 3, 4, 5])
            for nm, fl in (("parallel", E.F_ROUNDS_PARALLEL), ("serial", E.F_ROUNDS_SERIAL), ("adaptive", E.F_ROUNDS_ADAPTIVE)):
                ms, c = timed(E.COAT, gc, csz, csd, fl, w=(0.65, 0.35), steps=4000000)
                report("coat genes %d..10 pop50 4e6 (%s)" % (lo, nm), ms, c)
    if "xover2" in what or "sepx" in what:
        behs = ((E.LOCO, "loco", (13, 9), (0.75, 0.25), 144), (E.COAT, "coat", (20, 5, 2), (0.65, 0.35), 600))
        if "sepx" in what:
            behs = ((E.SEP, "sep", (13, 9), (0.65, 0.35), 600),)
        for beh, nm, size, w, L in behs:
            rows = []
            for lo in range(0, 9):
                for rep in range(2):
                    g = np.random.default_rng(900 + 10 * lo + rep).integers(lo, 11, size=(1, L), dtype=np.uint8)
                    ori = np.asarray([1, 2, 3, 1, 2], dtype=np.uint8) if beh == E.LOCO else None
                    r = []
                    for fl in (E.F_ROUNDS_PARALLEL, E.F_ROUNDS_SERIAL, E.F_ROUNDS_ADAPTIVE):
                        ms, c = timed(beh, g, [size] * 5, [1, 2, 3, 4, 5], fl, steps=1000000, w=w, orient=ori)
                        r.append(ms * 1e-3 * 1.965e9 / 1e6)
                    rows.append((c["committed"] / c["attempts"], r[0], r[1], r[2]))
            rows.sort()
            for pa, cp, cs, ca in rows:
                print("    %s p_acc=%.4f  cycles/step parallel=%6.1f serial=%6.1f adaptive=%6.1f" % (nm, pa, cp, cs, ca), flush=True)
    if "g0" in what:
        for P in (100, 200):
            g0 = bench.synthetic_genomes(0, P, 48)
            for nm, fl in (("parallel", E.F_ROUNDS_PARALLEL), ("adaptive", E.F_ROUNDS_ADAPTIVE)):
                ms, c = timed(E.AGG, g0, tsz, tsd, fl)
                report("generation-0 pop%d (%s)" % (P, nm), ms, c)
    if "vsweep" in what:
        ev = np.asarray(bench.EVOLVED_AGG, dtype=np.uint8).reshape(1, 48)
        for nm, fl in (("serial", E.F_ROUNDS_SERIAL), ("parallel", E.F_ROUNDS_PARALLEL)):
            ms, c = timed(E.AGG, ev, [(13, 9)] * 20000, list(range(1, 20001)), fl, steps=100000, mode=E.RNG_VERIFY)
            report("verify sweep 20000 x (13,9) x 1e5 evolved (%s)" % nm, ms, c)
    if "verify" in what:
        ms, c = timed(E.AGG, gs, tsz, tsd, E.F_ROUNDS_PARALLEL, mode=E.RNG_VERIFY, reps=1)
        report("C2 generation verify (parallel)", ms, c)
    if "sep" in what:
        gsep = np.random.default_rng(3).integers(0, 11, size=(200, 600), dtype=np.uint8)
        ssz, ssd = bench.trial_list([(13, 9)], [1, 2, 3, 4, 5])
        for nm, fl in (("parallel", E.F_ROUNDS_PARALLEL), ("serial", E.F_ROUNDS_SERIAL), ("default", 0)):
            ms, c = timed(E.SEP, gsep, ssz, ssd, fl, w=(0.65, 0.35), steps=2000000)
            report("sep pop200 x (13,9) x5, 2e6 steps (%s)" % nm, ms, c)
    if "sepevo" in what:
        import json
        fx = json.load(open("tests/golden/reference_fixtures.json"))
        se = np.tile(np.asarray(fx["sep_genome_evolved"], dtype=np.uint8).reshape(-1), (50, 1))
        r6 = np.random.default_rng(6)
        for g in range(1, 50):
            for _ in range(6):
                k = int(r6.integers(0, 600))
                se[g, k] = min(10, max(0, int(se[g, k]) + (1 if r6.integers(0, 2) else -1)))
        ssz, ssd = bench.trial_list([(13, 9)], [1, 2, 3, 4, 5])
        for nm, fl in (("parallel", E.F_ROUNDS_PARALLEL), ("serial", E.F_ROUNDS_SERIAL), ("default", 0)):
            ms, c = timed(E.SEP, se, ssz, ssd, fl, w=(0.65, 0.35), steps=4000000)
            report("sep evolved pop50 x (13,9) x5, 4e6 steps (%s)" % nm, ms, c)
    if "loco" in what:
        glo = np.random.default_rng(4).integers(0, 11, size=(200, 144), dtype=np.uint8)
        lsz, lsd = bench.trial_list([(13, 9)], [1, 2, 3, 4, 5])
        ori = (np.arange(1000) % 3 + 1).astype(np.uint8)
        for nm, fl in (("parallel", E.F_ROUNDS_PARALLEL), ("serial", E.F_ROUNDS_SERIAL)):
            ms, c = timed(E.LOCO, glo, lsz, lsd, fl, w=(0.75, 0.25), orient=ori, steps=2000000)
            report("loco pop200 x (13,9) x5, 2e6 steps (%s)" % nm, ms, c)
    if "coat" in what:
        gc = np.random.default_rng(5).integers(0, 11, size=(200, 600), dtype=np.uint8)
        csz, csd = bench.trial_list([(20, 5, 2), (26, 8, 4)], [1, 2, 3, 4, 5])
        for nm, fl in (("parallel", E.F_ROUNDS_PARALLEL), ("serial", E.F_ROUNDS_SERIAL)):
            ms, c = timed(E.COAT, gc, csz, csd, fl, w=(0.65, 0.35), steps=1000000)
            report("coat pop200, 1e6 steps (%s)" % nm, ms, c)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2], sys.argv[3:])
    else:
        args = sys.argv[1:]
        cut = args.index("--") if "--" in args else len(args)
        what = args[:cut] or ["worst", "c2", "evo", "sweep"]
        variants = args[cut + 1:] or ["base=base"]
        for v in variants:
            name, path = v.split("=", 1)
            print("== %s (%s)" % (name, path), flush=True)
            rc = subprocess.call(["timeout", os.environ.get("EXP_TIMEOUT", "240"), sys.executable, __file__, "--child",
                                  path if path == "base" else os.path.abspath(path)] + what)
            if rc:
                print("  !! variant %s failed: rc=%d" % (name, rc), flush=True)
