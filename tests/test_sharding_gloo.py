"""world_size-2 gloo test of the N>1 host logic: contiguous theta ranges, final result gather,
max-over-ranks timing.  The per-rank "solver" here is the CPU oracle (tests may use it); on the
GPU box each rank calls the C ABI on its own device instead (bench.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util  # noqa: F401  (path setup)
from oracle import mpqp, ldp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    import lmpc_b200
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mpc = mpqp.invpend()
    P = ldp.OracleProblem(mpqp.qp2ldp(mpqp.mpc2mpqp(mpc), mpc.nu), mpc.nu)
    Th = mpqp.sample_theta(*mpqp.invpend_range(1.0), n_total)        # every rank sees the same stream
    lo, hi = lmpc_b200.shard_range(n_total, rank, world)
    out = P.solve_batch(Th[:, lo:hi], nthreads=1)
    U = lmpc_b200.gather_columns(torch.from_numpy(np.ascontiguousarray(out["U"])), n_total)
    tmax = lmpc_b200.max_over_ranks(1.0 + rank)
    if rank == 0:
        q.put((U.numpy().copy(), tmax, (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [1001, 64])
def test_sharded_equals_unsharded(n_total):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    U, tmax, r0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    mpc = mpqp.invpend()
    P = ldp.OracleProblem(mpqp.qp2ldp(mpqp.mpc2mpqp(mpc), mpc.nu), mpc.nu)
    ref = P.solve_batch(mpqp.sample_theta(*mpqp.invpend_range(1.0), n_total))
    assert np.array_equal(U, ref["U"])
    assert tmax == 2.0 and r0 == (0, n_total // 2)


def test_shard_ranges_partition_the_batch():
    import lmpc_b200
    for n in (0, 1, 7, 10_000_000):
        for w in (1, 2, 3, 8):
            rs = [lmpc_b200.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert max(h - l for l, h in rs) - min(h - l for l, h in rs) <= 1
    with pytest.raises(ValueError):
        lmpc_b200.shard_range(10, 2, 2)
