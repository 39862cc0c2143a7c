"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row sharding + all-gather of the kNN lists
must reproduce the single-rank result bit for bit, for even and uneven shards.  The single-GPU build is
replaced by the oracle's canonical top-k here (test infrastructure); on GPUs it is ops.knn_cosine."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _local_topk(q, db, k):
    sys.path.insert(0, ROOT)
    from oracle import build_oracle as bo
    sim = torch.sigmoid(torch.nn.functional.normalize(q, dim=1) @ torch.nn.functional.normalize(db, dim=1).t())
    v, i = bo.canonical_topk(sim, k)
    gap = torch.full((q.shape[0],), float("inf")) if db.shape[0] <= k else \
        (torch.sort(sim, dim=1, descending=True, stable=True).values[:, k - 1:k + 1].diff(dim=1).abs().view(-1))
    return i, v, gap


def _worker(rank, world, port, nq, ndb, k, q_is_sharded, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from bridged_gnn_b200 import dist as bd
    g = torch.Generator().manual_seed(0)
    q, db = torch.randn(nq, 16, generator=g), torch.randn(ndb, 16, generator=g)
    if q_is_sharded:
        s, e = bd.row_shard(nq, rank, world)
        idx, val, gap = bd.sharded_topk(q[s:e].clone(), db, k, _local_topk, q_is_sharded=True, n_total=nq)
    else:
        idx, val, gap = bd.sharded_topk(q, db, k, _local_topk)
    edges = bd.edges_from_topk(idx)
    ret[rank] = (idx, val, gap, edges)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nq,ndb,k,q_is_sharded", [(10, 40, 3, False), (11, 40, 5, True), (1, 9, 2, False), (64, 64, 64, True)])
def test_two_rank_build_equals_single_rank(nq, ndb, k, q_is_sharded):
    port = 29500 + (os.getpid() % 2000) + nq
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, nq, ndb, k, q_is_sharded, ret), nprocs=2, join=True)
    g = torch.Generator().manual_seed(0)
    q, db = torch.randn(nq, 16, generator=g), torch.randn(ndb, 16, generator=g)
    i0, v0, g0 = _local_topk(q, db, k)
    for r in (0, 1):
        idx, val, gap, edges = ret[r]
        assert torch.equal(idx, i0) and torch.equal(val, v0) and torch.equal(gap, g0)
        assert edges.shape == (2, nq * k) and torch.equal(edges[1], torch.arange(nq).repeat_interleave(k))


def test_row_shard_partitions_exactly():
    from bridged_gnn_b200.dist import row_shard
    for n in (0, 1, 7, 8, 262144, 262145):
        for w in (1, 2, 3, 8):
            spans = [row_shard(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
