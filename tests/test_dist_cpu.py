"""world_size-2 gloo run of the multi-GPU plumbing on CPU: model broadcast, shard ranges, max-over-ranks."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from gp_emulator_b200.sharding import broadcast_model, max_over_ranks, shard_range, sum_over_ranks
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = None
        if rank == 0:
            rs = np.random.RandomState(0)
            model = {"inputs": rs.random_sample((7, 3)), "theta": rs.random_sample(5), "invQ": rs.random_sample((7, 7)),
                     "invQt": rs.random_sample(7)}
        got = broadcast_model(model, src=0)
        lo, hi = shard_range(1001, rank, world)
        checksum = float(sum(v.sum() for v in got.values()))
        tmax = max_over_ranks(1.0 + rank)
        total = sum_over_ranks(hi - lo)
        q.put((rank, lo, hi, checksum, tmax, total, {k: v.shape for k, v in got.items()}))
    finally:
        dist.destroy_process_group()


def test_broadcast_and_shard_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, c0, t0, n0, s0), (r1, lo1, hi1, c1, t1, n1, s1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 501, 501, 1001)
    assert c0 == c1 and s0 == s1 and s0["invQ"] == (7, 7)
    assert t0 == t1 == 2.0 and n0 == n1 == 1001.0
