"""N > 1 path on CPU: world_size-2 gloo processes exercise the sharding / gather / max-over-ranks
plumbing bench.py uses under torchrun (gpad_b200/sharding.py).  The per-shard "solver" here is the
CPU oracle standing in for the GPU (tests may use it); the real kernels are covered by -m gpu."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total, ragged, q):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "gpu-dualgradient-mpc_b200")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import problems as P
    from gpad_b200 import sharding
    from oracle import Oracle, schedule
    pb = P.battery(3, 4)
    X0 = np.random.default_rng(0).random((total, 3)) - 0.5          # the whole job, identical on every rank
    lo, hi = sharding.shard_range(total, rank, world)
    g_P, p_D, _ = pb.instance(X0[lo:hi])
    theta, beta = schedule(100)
    sol = Oracle().solve_batch(3, 4, pb.m, pb.M_G, pb.G_L, g_P, p_D, theta, beta, nthreads=1)
    z = torch.from_numpy(sol["z"])
    u0 = sharding.gather_first_moves(z, 3, dst=0)
    slowest = sharding.max_over_ranks(10.0 + rank)
    if rank == 0:
        q.put((u0.numpy(), slowest, (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 7])
def test_two_rank_shard_and_gather_equals_single_process(total):
    import problems as P
    from oracle import Oracle, schedule
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, total % 2, q)) for r in range(2)]
    for p in procs:
        p.start()
    u0, slowest, rng0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pb = P.battery(3, 4)
    X0 = np.random.default_rng(0).random((total, 3)) - 0.5
    g_P, p_D, _ = pb.instance(X0)
    theta, beta = schedule(100)
    full = Oracle().solve_batch(3, 4, pb.m, pb.M_G, pb.G_L, g_P, p_D, theta, beta, nthreads=2)
    assert u0.shape == (total, 3)
    assert np.array_equal(u0, full["z"][:, :3])          # sharded job == single-process job, bit for bit
    assert slowest == 11.0                                # max over ranks
    assert rng0 == (0, (total + 1) // 2)


def test_shard_ranges_partition_the_batch():
    from gpad_b200 import sharding
    for total in (1, 7, 64, 65536, 1000003):
        for world in (1, 2, 4, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
