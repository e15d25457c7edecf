"""Multi-rank host logic on the CPU (world_size 2, gloo): region shards are dealt to the ranks, each rank computes only
its shards (from the reads overlapping the shard, as the feeder would deliver them), rank 0 gathers the rows in window
order, and the result equals the single-process run.  The per-shard compute here is the CPU oracle standing in for the
GPU (this is a test of the sharding / halo / gather logic, which has no collective on the data path)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import pbtest
from popbam_b200 import sharding


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(pbtest.ROOT / "tests"))
    fx = pbtest.Fixture(contig_len=30500, n_ingroup=6, has_outgroup=1, depth=10.0, snp_density=0.02, seed=21)
    p = fx.params()
    an = pbtest.AN["NUCDIV"]
    wb, we = pbtest.window_grid(0, fx.contig_len, 5000)
    shards = sharding.plan_shards(wb, we, 10000)
    rows = {}
    for s in sharding.shards_of_rank(len(shards), rank, world):
        w0, w1 = shards[s]
        # the oracle takes the whole batch; reads outside the shard's span are ignored by the window test exactly as
        # the halo reads of a feeder-delivered shard would be
        run = pbtest.OracleRun(p, fx.batch(), fx.ref(), an, wb[w0:w1], we[w0:w1])
        rows[s] = run.text(an, fx.print_opts())
        run.close()
    text = sharding.gather_rows(rows, rank, world)
    slow = sharding.max_over_ranks(1.0 + rank, world)
    if rank == 0:
        q.put((text, slow, len(shards)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_run_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    text, slow, n_shards = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert n_shards == 3 and slow == 2.0
    fx = pbtest.Fixture(contig_len=30500, n_ingroup=6, has_outgroup=1, depth=10.0, snp_density=0.02, seed=21)
    wb, we = pbtest.window_grid(0, fx.contig_len, 5000)
    one = pbtest.OracleRun(fx.params(), fx.batch(), fx.ref(), pbtest.AN["NUCDIV"], wb, we)
    assert text == one.text(pbtest.AN["NUCDIV"], fx.print_opts())


@pytest.mark.parametrize("shard_bp,expect", [(1, [(0, 1), (1, 2), (2, 3)]), (25000, [(0, 2), (2, 3)]), (10 ** 9, [(0, 3)])])
def test_plan_shards(shard_bp, expect):
    wb, we = np.array([0, 10000, 20000]), np.array([9999, 19999, 29999])
    assert sharding.plan_shards(wb, we, shard_bp) == expect
    assert sharding.shards_of_rank(5, 1, 2) == [1, 3]
