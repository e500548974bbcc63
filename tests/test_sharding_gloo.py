"""The N>1 host path on CPU: two gloo ranks own contiguous column blocks, draw their own shard of the
workload, and the end-of-run gather reassembles per-column summaries in column order."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lgar_c_b200.sharding import gather_summaries, reduce_totals, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ntotal, q):
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    col0, ncol = shard_range(ntotal, rank, world)
    soils = bench.column_soils(col0, ncol)
    layers = bench.column_layers(col0, ncol)
    # a stand-in for the per-column summary an engine would return: depends on the column's identity only
    summary = torch.from_numpy(np.concatenate([soils.astype(np.float64), layers, np.arange(col0, col0 + ncol, dtype=np.float64)[:, None]], axis=1))
    g = gather_summaries(summary, dist, dst=0)
    tot = reduce_totals(summary.sum(dim=0), dist)
    if rank == 0:
        q.put((g.numpy(), tot.numpy()))
    else:
        assert g is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition_the_columns():
    for n, w in ((10, 2), (10, 3), (7, 8), (1_250_000 * 8, 8), (5, 1)):
        got = [shard_range(n, r, w) for r in range(w)]
        assert got[0][0] == 0 and sum(c for _, c in got) == n
        for (a0, an), (b0, _) in zip(got, got[1:]):
            assert a0 + an == b0


def test_two_rank_gather_matches_single_rank():
    import bench
    ntotal, world = 1001, 2     # ragged last shard
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ntotal, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, totals = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    full = np.concatenate([bench.column_soils(0, ntotal).astype(np.float64), bench.column_layers(0, ntotal), np.arange(ntotal, dtype=np.float64)[:, None]], axis=1)
    assert np.array_equal(gathered, full), "gathered shards are not the single-rank result in column order"
    assert np.allclose(totals, full.sum(axis=0))
