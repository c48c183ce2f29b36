"""BASELINE config 5 as stated: ONE fixed genome (the reference's evolved aggregation genome) x 10^5 seeds x particle sizes
(13,9), (20,14), (26,18) [271 / 631 / 1027 particles], the 10^5 chains of a size sharded over the N ranks (strong scaling).
Chains are capped at --cap steps (default 1e7: a full (26,18) chain is 1.08e9 steps, 10^5 of them 1e14 attempts) — the cap is
stated in every line.  One JSON line per size: aggregate attempts/s (device time, max over ranks), p_poss, p_acc, and the
SURVEY 8d issue-ceiling fraction.  Run:  python scripts/c5_sweep.py   or   torchrun --nproc-per-node N scripts/c5_sweep.py"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=100000)
    ap.add_argument("--cap", type=int, default=10000000)
    ap.add_argument("--rng", default="fast", choices=["fast", "verify"])
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    import torch
    import torch.distributed as dist
    import evosops_b200 as E
    from evosops_b200 import sharding
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        saved = os.dup(1); os.dup2(2, 1)                      # NCCL's banner goes to stderr
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier(); torch.cuda.synchronize(); sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)

    def red(v, op):
        if world == 1:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    ctx = E.Context(device_ids=[local])
    genome = np.asarray(bench.EVOLVED_AGG, dtype=np.uint8).reshape(1, 48)
    mode = E.RNG_FAST if args.rng == "fast" else E.RNG_VERIFY
    lo, hi = sharding.shard_range(args.seeds, world, rank)
    seeds = list(range(1 + lo, 1 + hi))
    for size in ((13, 9), (20, 14), (26, 18)):
        n = 3 * size[1] * (size[1] + 1) + 1
        steps = min(args.cap, n ** 3)
        ctx.prepare(E.AGG, genome, [size] * len(seeds), seeds, rng_mode=mode, steps_override=steps, flags=E.F_ROUNDS_SERIAL)
        ctx.launch(); ctx.sync()                             # warm-up
        best = None
        for _ in range(2):
            if world > 1:
                dist.barrier()
            ctx.launch()
            ms = red(ctx.sync(), dist.ReduceOp.MAX if world > 1 else None)
            best = ms if best is None else min(best, ms)
        res = ctx.collect()
        c = ctx.counters()
        att = red(c["attempts"], dist.ReduceOp.SUM if world > 1 else None)
        poss = red(c["possible"], dist.ReduceOp.SUM if world > 1 else None)
        acc = red(c["committed"], dist.ReduceOp.SUM if world > 1 else None)
        fit = red(float(res["norm"].sum()), dist.ReduceOp.SUM if world > 1 else None) / args.seeds
        if rank == 0:
            rate = att / (best * 1e-3)
            iv = bench.issue_view(rate / world, poss / att, 1965.0)
            print(json.dumps({"config": "C5 sweep: reference evolved agg genome x %d seeds x (%d,%d)" % (args.seeds, size[0], size[1]),
                              "n_gpus": world, "particles": n, "chain_steps": steps, "steps_cap": args.cap, "full_chain_steps": n ** 3,
                              "rng": args.rng, "instances": args.seeds, "attempts": att, "ms": best, "attempts_per_s": rate,
                              "attempts_per_s_per_gpu": rate / world, "p_poss": poss / att, "p_acc": acc / att,
                              "issue_ceiling_frac_per_gpu": iv["frac"], "alg_thread_inst_per_attempt": iv["alg_thread_inst_per_attempt"],
                              "mean_fitness_at_cap": fit}), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
