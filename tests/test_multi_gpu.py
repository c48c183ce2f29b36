"""N > 1 on hardware: the population sharded over two GPUs, one process per GPU, NCCL all_gather of the results
(evosops_b200.sharding), and the one-process multi-device context (sops_ctx_create with n_dev > 1, LPT split).
Both must give exactly the single-GPU results.  Skipped on a one-GPU box (run with `gpurun --gpus 2`)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    rng = np.random.default_rng(17)
    gs = rng.integers(0, 11, size=(7, 48), dtype=np.uint8)
    sizes, seeds = [(6, 3), (10, 7), (13, 9)], [1, 2, 3]
    ms = (np.arange(7 * 3, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(5)).reshape(7, 3)
    return gs, sizes, seeds, ms


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import evosops_b200 as E
    from evosops_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    gs, sizes, seeds, ms = _inputs()
    ctx = E.Context(device_ids=[rank])
    for mode in (E.RNG_VERIFY, E.RNG_FAST):
        def evaluate(g_shard, ms_shard, _orient):
            return ctx.eval_batch(E.AGG, g_shard, sizes, seeds, rng_mode=mode, move_seeds=ms_shard, steps_override=200000)
        full = sharding.evaluate_sharded(evaluate, gs, 3, move_seeds=ms)
        np.save(os.path.join(out_dir, "r%d_m%d.npy" % (rank, mode)), full)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_nccl_gather_equals_single_gpu(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import evosops_b200 as E
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    gs, sizes, seeds, ms = _inputs()
    one = E.Context(device_ids=[0])
    try:
        for mode in (E.RNG_VERIFY, E.RNG_FAST):
            want = one.eval_batch(E.AGG, gs, sizes, seeds, rng_mode=mode, move_seeds=ms, steps_override=200000)
            for r in range(2):
                got = np.load(os.path.join(str(tmp_path), "r%d_m%d.npy" % (r, mode)))
                assert got.dtype == want.dtype and np.array_equal(got, want), (mode, r)
    finally:
        one.close()


def test_multi_device_context_all_behaviours():
    # one process, two devices: cost-sorted instances LPT-assigned to the devices; same bits as one device
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import evosops_b200 as E
    rng = np.random.default_rng(23)
    one, two = E.Context(device_ids=[0]), E.Context(device_ids=[0, 1])
    W = {E.AGG: (0.0, 0.0), E.SEP: (0.65, 0.35), E.COAT: (0.65, 0.35), E.LOCO: (0.75, 0.25)}
    try:
        for beh in (E.AGG, E.SEP, E.COAT, E.LOCO):
            gs = rng.integers(0, 11, size=(5, E.GENOME_LEN[beh]), dtype=np.uint8)
            sizes = [(12, 2, 1), (20, 5, 2), (26, 8, 4)] if beh == E.COAT else [(6, 3), (13, 9), (20, 14)]
            seeds = [4, 5, 6]
            ori = rng.integers(1, 4, size=(5, 3)).astype(np.uint8) if beh == E.LOCO else None
            for mode in (E.RNG_VERIFY, E.RNG_FAST):
                kw = dict(w1=W[beh][0], w2=W[beh][1], rng_mode=mode, steps_override=150000, orient=ori, flags=E.F_DUMP)
                a = one.eval_batch(beh, gs, sizes, seeds, **kw)
                b = two.eval_batch(beh, gs, sizes, seeds, **kw)
                assert np.array_equal(a, b), (beh, mode)
                for q in range(15):
                    G, n, _ = E.instance_shape(beh, sizes[q % 3])
                    ga, pa = one.dump_state(q, G, n)
                    gb, pb = two.dump_state(q, G, n)
                    assert np.array_equal(ga, gb) and np.array_equal(pa, pb)
    finally:
        one.close()
        two.close()
