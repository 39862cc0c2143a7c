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


# ------------------------------------------------------------------ destination-partitioned message passing
def _mp_worker(rank, world, port, balanced, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    from bridged_gnn_b200 import dist as bd
    from bridged_gnn_b200 import ops
    from bridged_gnn_b200.data import Data
    from bridged_gnn_b200.models import KTGNN_no_complement, graph_partition
    from oracle import mp_oracle as mo

    # CUDA operators -> oracle on the CPU (test infrastructure, as in tests/test_host_logic.py).  The fakes follow the
    # partitioned calling convention: H over all nodes (None for an operand this rank never reads), result = the
    # rank's own destination rows.
    class FakeGraph:
        def __init__(self, ei, n, n_rows=None, row_off=0):
            self.edge_index, self.n, self.n_src = ei, n, n
            self.n_rows, self.row_off = (n if n_rows is None else n_rows), row_off
    ops.cached_graph = lambda ei, n, n_rows=None, row_off=0: FakeGraph(ei, n, n_rows, row_off)

    def gat(Hs, Ht, a1, a2, graph, dst_is_src, slope=0.1):
        cm = dst_is_src.bool()
        ei = graph.edge_index
        m1 = cm[ei[1]]
        if Hs is None:
            assert not bool(m1.any())          # no source-domain destination on this rank
            Hs = torch.zeros_like(Ht)
        if Ht is None:
            assert bool(m1.all())
            Ht = torch.zeros_like(Hs)
        full = mo.adapted_conv_aggregate(Hs, Ht, ei[:, m1], ei[:, ~m1], cm, a1.view(-1), a2.view(-1), slope)
        return full[graph.row_off: graph.row_off + graph.n_rows]
    ops.gat_aggregate = gat
    ops.adapted_transform = mo.adapted_transform_epilogue

    gb = dict(np.load(os.path.join(ROOT, "tests", "golden", "office_a2d_build.npz")))
    gm = dict(np.load(os.path.join(ROOT, "tests", "golden", "office_a2d_mp.npz")))
    T = torch.from_numpy
    x, y, cm, tm = T(gb["x"]), T(gb["y"]), T(gb["central_mask"]), T(gb["train_mask"])
    ei_u = T(gm["edge_index_undirected"])
    n = x.shape[0]
    sd = {k[len("ktgnn.sd."):]: T(v) for k, v in gm.items() if k.startswith("ktgnn.sd.")}
    _, _, ei_all = graph_partition(ei_u, cm)
    # equal row blocks, or blocks cut at equal incoming-edge counts (unequal sizes, point-to-point exchanges only)
    part = bd.DstPartition(n, bounds=bd.DstPartition.balanced_bounds(ei_all[1], n, world) if balanced else None)
    data_loc = Data(x=part.local_rows(x), edge_index=part.local_edges(ei_all), central_mask=part.pad_rows(cm), part=part)

    def gather_rows(t):          # this rank's rows into place, summed over ranks (blocks may differ in size)
        out = t.new_zeros((n,) + tuple(t.shape[1:]))
        out[part.r0:part.r1] = t[: part.r1 - part.r0]
        dist.all_reduce(out)
        return out

    # (1) eval forward == reference logits
    model = KTGNN_no_complement(256, 31, 2, 64, root_weight=False, use_bn=True, dim_share=256)
    model.load_state_dict(sd)
    model.eval()
    with torch.no_grad():
        lb, lt, ltt, _ = model(data_loc)
    full = [gather_rows(t) for t in (lb, lt, ltt)]

    # (2) train-mode gradients (no BatchNorm: SyncBatchNorm is CUDA-only) == single-process gradients
    torch.manual_seed(0)
    m2 = KTGNN_no_complement(256, 31, 2, 64, root_weight=False, use_bn=False, dim_share=256, dropout=0.0)
    ref = KTGNN_no_complement(256, 31, 2, 64, root_weight=False, use_bn=False, dim_share=256, dropout=0.0)
    ref.load_state_dict(m2.state_dict())
    nll = torch.nn.functional.nll_loss
    cnt = int(tm.sum())
    m2.train(), ref.train()
    m2.clf_transformer[1].eval(), ref.clf_transformer[1].eval()      # batch statistics would need SyncBatchNorm (CUDA only)
    out = m2(data_loc)
    tm_loc, y_loc = part.local_rows(tm), part.local_rows(y)
    loss = sum(nll(o[tm_loc], y_loc[tm_loc], reduction="sum") for o in out[:3]) / cnt
    loss.backward()
    part.sync_grads(m2)
    out_r = ref(Data(x=x, edge_index=ei_u, central_mask=cm))
    loss_r = sum(nll(o[tm], y[tm], reduction="sum") for o in out_r[:3]) / cnt
    loss_r.backward()
    errs = {k: float((p.grad - q.grad).abs().max() / (q.grad.abs().max() + 1e-12))
            for (k, p), q in zip(m2.named_parameters(), ref.parameters())}
    gerr = max(errs.values())
    if rank == 0 and gerr > 5e-5:
        print({k: "%.2e" % v for k, v in errs.items() if v > 5e-5}, file=sys.stderr)
    ltot = loss.detach().clone()
    dist.all_reduce(ltot)
    ret[rank] = ([f.clone() for f in full], gerr, float(ltot), float(loss_r))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("balanced", [False, True])
def test_two_rank_partitioned_ktgnn_matches_single_rank(balanced):
    import numpy as np
    port = 31500 + (os.getpid() % 2000) + (7 if balanced else 0)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_mp_worker, args=(2, port, balanced, ret), nprocs=2, join=True)
    gm = dict(np.load(os.path.join(ROOT, "tests", "golden", "office_a2d_mp.npz")))
    for r in (0, 1):
        full, gerr, ltot, lref = ret[r]
        for got, key in zip(full, ("ktgnn.eval.logp_base", "ktgnn.eval.logp_target", "ktgnn.eval.logp_trans")):
            want = torch.from_numpy(gm[key])
            assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max()) + 1e-7, key
        assert gerr < 5e-5
        assert abs(ltot - lref) < 1e-5 * max(1.0, abs(lref))


# ------------------------------------------------------------------ domain-aware halo exchange
def _halo_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from bridged_gnn_b200 import dist as bd
    n, ns, c = 12, 8, 3                      # 8 source + 4 target nodes over 3 ranks: ranks 0, 1 source-only, rank 2 target-only
    cm = torch.arange(n) < ns
    part = bd.DstPartition(n)
    needs = part.needs(cm)
    g = torch.Generator().manual_seed(5)
    h_s_all, h_t_all = torch.randn(n, c, generator=g), torch.randn(n, c, generator=g)
    w_s, w_t = torch.randn(world, n, c, generator=g), torch.randn(world, n, c, generator=g)     # per-rank loss weights
    h_s = part.local_rows(h_s_all).requires_grad_(True)
    h_t = part.local_rows(h_t_all).requires_grad_(True)
    H_s, H_t = bd.halo_exchange(h_s, h_t, part, cm)
    ok = True
    loss = h_s.sum() * 0.0
    if needs[rank][0]:
        ok &= torch.equal(H_s, h_s_all)
        loss = loss + (H_s * w_s[rank]).sum()
    else:
        ok &= H_s is None
    if needs[rank][1]:
        ok &= torch.equal(H_t, h_t_all)
        loss = loss + (H_t * w_t[rank]).sum()
    else:
        ok &= H_t is None
    loss.backward()
    want_s = sum(w_s[r] for r in range(world) if needs[r][0])[part.r0:part.r1]
    want_t = sum(w_t[r] for r in range(world) if needs[r][1])[part.r0:part.r1]
    ret[rank] = (bool(ok), needs, float((h_s.grad - want_s).abs().max()), float((h_t.grad - want_t).abs().max()))
    dist.barrier()
    dist.destroy_process_group()


def test_halo_exchange_sends_each_operand_only_where_it_is_read():
    port = 33500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_halo_worker, args=(3, port, ret), nprocs=3, join=True)
    for r in range(3):
        ok, needs, es, et = ret[r]
        assert needs == ((True, False), (True, False), (False, True))
        assert ok and es < 1e-6 and et < 1e-6
