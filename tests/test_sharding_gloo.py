"""N>1 host path on CPU: world_size-2 gloo.  The kernels cannot run here, so the local evaluator is the
oracle (tests may use it as the checker); what is under test is the partition + gather logic of
evosops_b200.sharding — results must equal the single-process evaluation for every world size."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, P, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from evosops_b200 import sharding
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    gs = rng.integers(0, 11, size=(P, 48), dtype=np.uint8)
    sizes, seeds = [(6, 3), (6, 3), (7, 4)], [1, 2, 1]
    ms = (np.arange(P * 3, dtype=np.uint64) * np.uint64(7919) + np.uint64(11)).reshape(P, 3)

    def evaluate(g_shard, ms_shard, _orient):
        res, _ = O.eval_batch(O.AGG, g_shard, sizes, seeds, rng_mode=1, move_seeds=ms_shard, n_threads=2)
        return res

    full = sharding.evaluate_sharded(evaluate, gs, 3, move_seeds=ms)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("P", [5, 2, 1])
def test_two_rank_gather_equals_single_process(tmp_path, P):
    from evosops_b200 import sharding
    from oracle import oracle as O
    port = _free_port()
    mp.spawn(_worker, args=(2, port, P, str(tmp_path)), nprocs=2, join=True)
    rng = np.random.default_rng(5)
    gs = rng.integers(0, 11, size=(P, 48), dtype=np.uint8)
    sizes, seeds = [(6, 3), (6, 3), (7, 4)], [1, 2, 1]
    ms = (np.arange(P * 3, dtype=np.uint64) * np.uint64(7919) + np.uint64(11)).reshape(P, 3)
    want, _ = O.eval_batch(O.AGG, gs, sizes, seeds, rng_mode=1, move_seeds=ms, n_threads=2)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "r%d.npy" % r))
        assert got.dtype == want.dtype and np.array_equal(got, want)


def test_shard_ranges_partition_the_population():
    from evosops_b200.sharding import shard_range
    for P in (0, 1, 7, 50, 200):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(P, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == P
            assert all(spans[k][1] == spans[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
