"""One-process-per-GPU sharding of a population.

Instances are independent (base_ga.rs:393-411 runs them as unordered rayon tasks), so the population is
split into contiguous genome ranges, each rank evaluates its range on its own GPU with no data-path
collective, and the only exchange is one all_gather of the per-trial results (24 B each) so that every
rank holds the full fitness table for the host-side GA step.  Random streams are keyed by the global
(genome, trial) pair — never by rank or device — so results are identical for any world size.
"""
import numpy as np

from .capi import RESULT_DTYPE


def shard_range(P, world, rank):
    """Contiguous genome range [lo, hi) of `rank`; sizes differ by at most one."""
    base, extra = divmod(P, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def default_move_seeds(P, T, trial_seeds):
    """move_seeds[g, t] used when the caller passes none: the trial seed (include/evosops.h)."""
    return np.broadcast_to(np.asarray(trial_seeds, dtype=np.uint64)[None, :], (P, T)).copy()


def evaluate_sharded(evaluate, genomes, n_trials, move_seeds=None, orient=None, group=None):
    """Evaluate `genomes` [P, L] across the ranks of `group` (torch.distributed; None = default group or
    single process).  `evaluate(genomes_shard, move_seeds_shard, orient_shard) -> results[p, T]` runs the
    local shard (e.g. a closure over Context.eval_batch).  Returns results[P, T] on every rank."""
    import torch
    import torch.distributed as dist

    genomes = np.ascontiguousarray(genomes)
    P = genomes.shape[0]
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    lo, hi = shard_range(P, world, rank)
    ms = None if move_seeds is None else np.asarray(move_seeds, dtype=np.uint64).reshape(P, n_trials)[lo:hi]
    orr = None if orient is None else np.asarray(orient, dtype=np.uint8).reshape(P, n_trials)[lo:hi]
    if hi > lo:
        local = np.ascontiguousarray(evaluate(genomes[lo:hi], ms, orr))
        assert local.shape == (hi - lo, n_trials) and local.dtype == RESULT_DTYPE
    else:
        local = np.zeros((0, n_trials), dtype=RESULT_DTYPE)
    if world == 1:
        return local
    # fixed-size all_gather of raw result bytes (padded to the largest shard)
    rows = -(-P // world)
    buf = np.zeros((rows, n_trials), dtype=RESULT_DTYPE)
    buf[:hi - lo] = local
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.from_numpy(buf.view(np.uint8).reshape(-1)).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    out = np.zeros((P, n_trials), dtype=RESULT_DTYPE)
    for r in range(world):
        rlo, rhi = shard_range(P, world, r)
        got = parts[r].cpu().numpy().view(RESULT_DTYPE).reshape(rows, n_trials)
        out[rlo:rhi] = got[:rhi - rlo]
    return out
