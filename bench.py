#!/usr/bin/env python3
"""bench.py — particle-move attempts/s (and GA generations/s) of the SOPS fitness-evaluation hot path.

A "step" is one GA generation's evaluation batch (what base_ga.rs:393-411's nested par_iter does):
config C2 of BASELINE.json / SURVEY.md §8d — aggregation, pop 50 x sizes {(6,4),(10,7),(13,9)} x seeds 1..5
= 750 independent chains of n^3 attempts each (6.24e9 attempts).  With --gpus N the population is 50*N genomes
sharded over the ranks (weak scaling: 50 genomes per GPU; instances are independent, so the chain kernels need no
collective).

  value   attempts/s, inputs resident in HBM, CUDA-event time of the chain kernel on its own stream
  e2e     the same through the public sharded call (evosops_b200.sharding.evaluate_sharded over Context.eval_batch):
          HOST buffers, validation + H2D + kernel + D2H + float fitness on every rank, then ONE NCCL all_gather of the
          24-byte results so every rank holds the population's fitness table (what base_ga.rs:401-424 reduces)
  strong_scaling   fixed populations (agg pop 400, sep C3 pop 200) over the N ranks, beside the same population on one GPU
  --impl reference   the CPU oracle (the Rust reference cannot be built here: no cargo) on all host threads,
                     each step a bounded sample (--ref-pop genomes of the 50)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIZES_C2 = [(6, 4), (10, 7), (13, 9)]
SEEDS_C2 = [1, 2, 3, 4, 5]
POP_C2 = 50
GENOME_SEED = 0x45766F534F5053          # "EvoSOPS" 
# the reference's evolved aggregation genome (misc_scripts/genome_agg_compare.py:1), data only
EVOLVED_AGG = [4, 0, 0, 1, 6, 1, 0, 0, 1, 6, 10, 0, 7, 1, 2, 2, 3, 1, 0, 1, 4, 9, 5, 0,
               6, 1, 2, 0, 10, 10, 0, 1, 10, 2, 9, 8, 8, 7, 3, 0, 9, 9, 7, 8, 7, 4, 7, 2]


def splitmix64(state):
    state = (state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return state, z ^ (z >> 31)


def synthetic_genomes(first, count, length, gran=10):
    """i.i.d. uniform genes in 0..=gran from SplitMix64; genome g depends only on its global index."""
    out = np.zeros((count, length), dtype=np.uint8)
    for g in range(count):
        st = (GENOME_SEED + 0x632BE59BD9B4E019 * (first + g)) & 0xFFFFFFFFFFFFFFFF
        for k in range(length):
            st, z = splitmix64(st)
            out[g, k] = z % (gran + 1)
    return out


def trial_list(sizes, seeds):
    # base_ga.rs:356-362: sizes-major, seeds-minor
    return [s for s in sizes for _ in seeds], [sd for _ in sizes for sd in seeds]


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 7] or [r for (_, r) in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows),
                "samples": len(rows), "reasons": reasons}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except (OSError, ValueError):
        return {"sm_max_mhz": 1965.0}, "fallback"


def cpu_leg(pop, gran, n_threads, rng_mode):
    """One bounded sample of the C2 batch on the host cores with the CPU oracle (checker code, timed as baseline)."""
    from oracle import oracle as O
    gs = synthetic_genomes(0, pop, 48, gran)
    tsz, tsd = trial_list(SIZES_C2, SEEDS_C2)
    t0 = time.perf_counter()
    _, cnt = O.eval_batch(O.AGG, gs, tsz, tsd, rng_mode=rng_mode, n_threads=n_threads)
    dt = time.perf_counter() - t0
    return float(cnt[0]), dt


def run_reference(args, rank, world):
    from oracle import oracle as O
    if rank != 0:
        return
    cores = O.lib().oracle_hardware_threads()
    rng_mode = 0                       # the reference's own move stream (fastrand WyRand)
    for _ in range(args.warmup):
        cpu_leg(args.ref_pop, args.gran, cores, rng_mode)
    attempts, secs = 0.0, 0.0
    for _ in range(args.steps):
        a, dt = cpu_leg(args.ref_pop, args.gran, cores, rng_mode)
        attempts += a
        secs += dt
    v = attempts / secs
    sample = "pop %d of %d genomes x %d sizes x %d seeds per step (oracle port, WyRand stream)" % (
        args.ref_pop, POP_C2, len(SIZES_C2), len(SEEDS_C2))
    line = {"impl": "reference", "metric": "particle-move attempts/s", "value": v, "unit": "attempts/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64/u8 integer", "data": "synthetic",
            "config": workload_config(args),
            "generations_per_s": v / attempts_per_generation(), "population": POP_C2,
            "cpu_baseline": {"value": v, "unit": "attempts/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "attempts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def issue_view(attempts_per_s, p_poss, sm_mhz):
    """SURVEY.md 8d's binding ceiling: SM issue.  ALG_INT_OPS/attempt = 50 + 90 p_poss thread-instructions,
    peak = 148 SMs x 4 warp-instructions/clk x 32 lanes x clock."""
    peak = 148 * 4 * 32 * sm_mhz * 1e6
    ops = 50.0 + 90.0 * p_poss
    ceiling = peak / ops
    return {"alg_thread_inst_per_attempt": ops, "peak_thread_inst_per_s": peak, "ceiling_attempts_per_s": ceiling,
            "frac": attempts_per_s / ceiling}


def attempts_per_generation():
    tot = 0
    for (a, p) in SIZES_C2:
        n = 3 * p * (p + 1) + 1
        tot += n ** 3 * len(SEEDS_C2) * POP_C2
    return float(tot)


def workload_config(args):
    return {"workload": "C2: aggregation GA generation (-b agg -e ga), pop %d x sizes %s x seeds 1..5, gran %d, n^3 attempts per chain"
            % (POP_C2, "".join("(%d,%d)" % s for s in SIZES_C2), args.gran),
            "instances_per_gpu": POP_C2 * len(SIZES_C2) * len(SEEDS_C2), "rng": args.rng,
            "sharding": "population sharded, %d genomes per GPU" % POP_C2,
            "l2": "flushed between timed steps (256 MiB write); working set is shared-memory resident",
            "reference_arm_sample": "--impl reference times %d of the %d genomes per step (a rate: attempts/s)" % (args.ref_pop, POP_C2)}


def ncu_measured(tag):
    """Issue-slot utilisation of the chain kernel as ncu measured it on this workload (profiles/r02_ncu_summary.json,
    written by scripts/ncu_summary.py from a `ncu --set full` capture of this same command)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_summary.json")) as f:
            return json.load(f).get(tag)
    except (OSError, ValueError):
        return None


def late_generation_population(rank, gran):
    """+-1 mutants of the reference's evolved genome (what the GA converges to): the C2 batch shape with few accepted moves."""
    evo = np.tile(np.asarray(EVOLVED_AGG, dtype=np.uint8), (POP_C2, 1))
    for g in range(1, POP_C2):
        st = (GENOME_SEED ^ (0x9E3779B97F4A7C15 * (rank * POP_C2 + g))) & 0xFFFFFFFFFFFFFFFF
        for _ in range(3):
            st, z = splitmix64(st)
            k = z % 48
            evo[g, k] = min(gran, max(0, int(evo[g, k]) + (1 if (z >> 32) & 1 else -1)))
    return evo


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rng", default="fast", choices=["fast", "verify"])
    ap.add_argument("--gran", type=int, default=10)
    ap.add_argument("--ref-pop", type=int, default=16)
    ap.add_argument("--no-extra", action="store_true", help="skip the side measurements (other mode, configs C3/C4, sweeps, strong scaling)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import evosops_b200 as E
    from evosops_b200 import sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to stdout when it creates the communicator: point fd 1 at stderr until the
        # first collective is through, so that stdout carries exactly ONE JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(val, op):
        if world == 1:
            return float(val)
        t = torch.tensor([val], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    MAX = dist.ReduceOp.MAX if world > 1 else None
    SUM = dist.ReduceOp.SUM if world > 1 else None

    ctx = E.Context(device_ids=[local])
    mode = E.RNG_FAST if args.rng == "fast" else E.RNG_VERIFY
    P_glob = POP_C2 * world
    gs_glob = synthetic_genomes(0, P_glob, 48, args.gran)       # genome g depends only on its global index
    lo, hi = sharding.shard_range(P_glob, world, rank)
    gs = gs_glob[lo:hi]
    tsz, tsd = trial_list(SIZES_C2, SEEDS_C2)
    T = len(tsz)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def l2_flush():
        flush.add_(1)
        torch.cuda.synchronize()

    def best_of(reps=3):
        """device ms of the prepared batch: one warm-up launch, then the best of `reps` (max over ranks each time)"""
        ctx.launch(); ctx.sync()
        best = None
        for _ in range(reps):
            ctx.launch()
            ms = reduce(ctx.sync(), MAX)
            best = ms if best is None else min(best, ms)
        return best

    # ---- device-resident leg: prepare once (H2D), time K launches with CUDA events on the launch stream
    ctx.prepare(E.AGG, gs, tsz, tsd, gran=args.gran, rng_mode=mode)
    for _ in range(max(args.warmup, 3)):
        ctx.launch()
        ctx.sync()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    t_wall0 = time.time()
    dev_ms = 0.0
    for _ in range(args.steps):
        l2_flush()
        ctx.launch()
        dev_ms += ctx.sync()
    barrier()
    t_wall1 = time.time()
    ctx.collect()
    cnt = ctx.counters()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    attempts_step = float(cnt["attempts"])
    p_poss = cnt["possible"] / cnt["attempts"]
    p_acc = cnt["committed"] / cnt["attempts"]
    launches_step = cnt["launches"]
    worst_ms = reduce(dev_ms, MAX)
    total_attempts = reduce(attempts_step, SUM) * args.steps
    value = total_attempts / (worst_ms * 1e-3)

    # ---- end-to-end leg: the public sharded call with host buffers; H2D, D2H and the NCCL all_gather of the 24-byte
    # results are inside the timed region (at N = 1 the gather is a no-op)
    def evaluate_local(g_shard, ms_shard, _orient):
        return ctx.eval_batch(E.AGG, g_shard, tsz, tsd, gran=args.gran, rng_mode=mode, move_seeds=ms_shard)

    res = sharding.evaluate_sharded(evaluate_local, gs_glob, T)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = sharding.evaluate_sharded(evaluate_local, gs_glob, T)
    barrier()
    e2e_s = time.perf_counter() - t0
    cnt2 = ctx.counters()
    e2e_worst = reduce(e2e_s, MAX)
    e2e_value = total_attempts / e2e_worst
    assert res.shape == (P_glob, T)
    mean_fitness = float(res["norm"].mean())
    rows = -(-P_glob // world)
    comm_bytes = 0 if world == 1 else rows * T * 24 * world       # bytes every rank receives in the all_gather

    extra = {}
    sweep_local = None
    if not args.no_extra:
        # (every side measurement: one warm-up launch, best of 3, device time, max over ranks)
        # the other RNG mode on the same batch
        other = E.RNG_VERIFY if mode == E.RNG_FAST else E.RNG_FAST
        ctx.prepare(E.AGG, gs, tsz, tsd, gran=args.gran, rng_mode=other)
        ms_o = best_of()
        ctx.collect()
        extra["verify_mode" if other == E.RNG_VERIFY else "fast_mode"] = {
            "value": reduce(attempts_step, SUM) / (ms_o * 1e-3), "unit": "attempts/s", "ms_per_step": ms_o}
        # ---- strong scaling: FIXED populations sharded over the N ranks (device time of the slowest rank), beside the
        # same population on ONE GPU in the same run (rank 0 alone); at N > 1 rank 0 also checks that the sharded
        # results are bit-identical to the single-GPU ones
        strong = {}
        for name, beh, pop, sizes, w in (("agg_pop400_c2_sizes", E.AGG, 400, SIZES_C2, (0.0, 0.0)),
                                         ("sep_pop200_c3", E.SEP, 200, [(13, 9)], (0.65, 0.35))):
            L = E.GENOME_LEN[beh]
            gfix = synthetic_genomes(0, pop, L, args.gran) if beh == E.AGG else \
                np.random.default_rng(3).integers(0, 11, size=(pop, L), dtype=np.uint8)
            ssz, ssd = trial_list(sizes, SEEDS_C2)
            slo, shi = sharding.shard_range(pop, world, rank)
            ctx.prepare(beh, gfix[slo:shi], ssz, ssd, gran=args.gran, w1=w[0], w2=w[1], rng_mode=mode)
            ms_n = best_of()
            mine = ctx.collect()
            rec = {"population": pop, "instances": pop * len(ssz), "ms_per_generation": ms_n, "n_gpus": world}
            if world > 1:
                ms_1 = 0.0
                same = 1.0
                if rank == 0:
                    ctx.prepare(beh, gfix, ssz, ssd, gran=args.gran, w1=w[0], w2=w[1], rng_mode=mode)
                    ctx.launch(); ctx.sync(); ctx.launch(); ms_1 = ctx.sync()
                    full = ctx.collect()
                    same = 1.0 if np.array_equal(full[slo:shi], mine) else 0.0
                barrier()
                ms_1 = reduce(ms_1, MAX)
                same = reduce(same, dist.ReduceOp.MIN)
                assert same == 1.0, "sharded results differ from the single-GPU evaluation of the same genomes"
                rec.update({"ms_per_generation_one_gpu_same_run": ms_1, "speedup": ms_1 / ms_n,
                            "efficiency": ms_1 / ms_n / world, "rank0_shard_equals_single_gpu": True})
            strong[name] = rec
        extra["strong_scaling"] = strong
        # a late-generation population: same batch shape as the headline, far fewer accepted moves per chain
        evo = late_generation_population(rank, args.gran)
        ctx.prepare(E.AGG, evo, tsz, tsd, gran=args.gran, rng_mode=mode)
        ms_e = best_of()
        res_e = ctx.collect()
        extra["evolved_population"] = {"value": reduce(attempts_step, SUM) / (ms_e * 1e-3), "unit": "attempts/s", "ms_per_step": ms_e,
                                       "generations_per_s": 1e3 / ms_e, "mean_fitness": float(res_e["norm"].mean())}
        # the other behaviours' GA generations (BASELINE configs 3 and 4: pop 200, random genomes), per GPU
        W = {E.SEP: (0.65, 0.35), E.COAT: (0.65, 0.35), E.LOCO: (0.75, 0.25)}
        for name, beh, sizes, gseed in (("c3_sep_pop200_13_9", E.SEP, [(13, 9)], 3), ("c4_loco_pop200_13_9", E.LOCO, [(13, 9)], 4),
                                        ("c4_coat_pop200_20_5_2_26_8_4", E.COAT, [(20, 5, 2), (26, 8, 4)], 5)):
            gb = np.random.default_rng(gseed + 1000 * rank).integers(0, 11, size=(200, E.GENOME_LEN[beh]), dtype=np.uint8)
            bsz, bsd = trial_list(sizes, SEEDS_C2)
            ori = (np.arange(200 * len(bsz)) % 3 + 1).astype(np.uint8) if beh == E.LOCO else None
            ctx.prepare(beh, gb, bsz, bsd, gran=args.gran, w1=W[beh][0], w2=W[beh][1], rng_mode=mode, orient=ori)
            ms_b = best_of()
            ctx.collect()
            cb = ctx.counters()
            extra[name] = {"value": reduce(float(cb["attempts"]), SUM) / (ms_b * 1e-3), "unit": "attempts/s", "ms_per_step": ms_b,
                           "instances_per_gpu": 200 * len(bsz), "p_poss": cb["possible"] / cb["attempts"],
                           "p_acc": cb["committed"] / cb["attempts"]}
        # throughput-bound seed sweeps (config 5 style, capped chains): one fixed genome x many seeds
        n_sweep, cap = 50000, 200000
        seeds = list(range(1 + rank * n_sweep, 1 + (rank + 1) * n_sweep))
        for nm, md in (("fast", E.RNG_FAST), ("verify", E.RNG_VERIFY)):
            ctx.prepare(E.AGG, np.asarray(EVOLVED_AGG, dtype=np.uint8).reshape(1, 48), [(13, 9)] * n_sweep, seeds, gran=args.gran,
                        rng_mode=md, steps_override=cap)
            ms_s = best_of()
            ctx.collect()
            cs = ctx.counters()
            if nm == "fast":
                sweep_local = (ms_s, cs["possible"] / cs["attempts"], cs["committed"] / cs["attempts"], float(cs["attempts"]))
            extra["sweep_13_9_%s" % nm] = {"value": reduce(float(n_sweep) * cap, SUM) / (ms_s * 1e-3),
                                           "unit": "attempts/s", "instances_per_gpu": n_sweep, "steps_cap": cap,
                                           "genome": "reference evolved agg genome"}

    if rank == 0:
        peaks, peak_src = measured_peaks()
        sm_mhz = float(peaks.get("sm_max_mhz", 1965.0))
        peak_smem = 148 * 128 * sm_mhz * 1e6 / 1e9                       # GB/s: 128 B/clk/SM
        alg_bytes = 10.0 + 22.3 * p_poss + 18.0 * p_acc
        k_ms = dev_ms / (args.steps * max(launches_step, 1))
        achieved = alg_bytes * attempts_step / max(launches_step, 1) / (k_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
        issue = issue_view(attempts_step / (k_ms * 1e-3), p_poss, sm_mhz)
        line = {"metric": "particle-move attempts/s", "value": value, "unit": "attempts/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": worst_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32/u64 integer",
                "data": "synthetic", "config": workload_config(args),
                "generations_per_s": 1e3 / (worst_ms / args.steps), "population": P_glob,
                "e2e": {"value": e2e_value, "unit": "attempts/s", "h2d_bytes_per_step": cnt2["h2d_bytes"],
                        "d2h_bytes_per_step": cnt2["d2h_bytes"], "ms_per_step": e2e_worst / args.steps * 1e3,
                        "api": "evosops_b200.sharding.evaluate_sharded(Context.eval_batch)",
                        "comm": {"collective": "all_gather" if world > 1 else None, "backend": "nccl" if world > 1 else None,
                                 "bytes_received_per_rank_per_step": comm_bytes}},
                "gpu_launches": int(launches_step * args.steps),
                "clocks": clocks,
                # the path is shared-memory resident: `bound` names the SM-side roofline.  frac/achieved/peak are the
                # shared-memory bandwidth view the contract's byte formula gives; `binding` is the ceiling that actually
                # binds (SM issue) and `measured` what ncu saw on this command.
                "roofline": {"bound": "smem", "achieved": achieved, "peak": peak_smem, "unit": "GB/s",
                             "frac": achieved / peak_smem, "traffic": traffic,
                             "kernel": "sops_chain_kernel<AGG,RW=1,%s,%s rounds>" % (args.rng, "adaptive" if args.rng == "fast" else "parallel"),
                             "kernel_ms": k_ms,
                             "alg_bytes_per_attempt": alg_bytes, "p_poss": p_poss, "p_acc": p_acc,
                             "peak_source": "%s sm_max_mhz x 148 SMs x 128 B/clk shared memory" % peak_src,
                             "binding": {"ceiling": "SM issue (thread-instructions)", "frac": issue["frac"], **issue},
                             "measured": ncu_measured("c2_bench_fast"),
                             "note": "750 chains pulled by 592 warps (one per warp scheduler): dependent-latency bound, the slowest genome's chains set the time; see DESIGN.md"},
                "mean_fitness": mean_fitness}
        if sweep_local:
            ms_s, pp_s, pa_s, att_s = sweep_local
            ab = 10.0 + 22.3 * pp_s + 18.0 * pa_s
            ach = ab * att_s / (ms_s * 1e-3) / 1e9
            iv = issue_view(att_s / (ms_s * 1e-3), pp_s, sm_mhz)
            extra["roofline_sweep"] = {"bound": "smem", "achieved": ach, "peak": peak_smem, "unit": "GB/s", "frac": ach / peak_smem,
                                       "traffic": None, "kernel": "sops_chain_kernel<AGG,RW=1,fast,serial rounds> on the 50000-chain sweep (rank 0)",
                                       "kernel_ms": ms_s, "alg_bytes_per_attempt": ab, "p_poss": pp_s, "p_acc": pa_s,
                                       "binding": {"ceiling": "SM issue (thread-instructions)", "frac": iv["frac"], **iv},
                                       "measured": ncu_measured("sweep_13_9_fast")}
        line.update(extra)
        if world == 1:
            import multiprocessing
            cores = multiprocessing.cpu_count()
            a, dt = cpu_leg(args.ref_pop, args.gran, cores, 0)
            line["cpu_baseline"] = {"value": a / dt, "unit": "attempts/s", "cores": cores, "kind": "port",
                                    "sample": "pop %d of %d genomes x 3 sizes x 5 seeds, one pass, oracle port (WyRand stream)"
                                              % (args.ref_pop, POP_C2), "seconds": dt}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
