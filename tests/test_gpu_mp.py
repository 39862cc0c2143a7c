"""GPU parity of message passing over the bridged graph (CSR build, SpMM, fused AdaptedConv
aggregation, KT-GNN / SAGE / GCN models) against the oracle and the reference-generated golden
vectors.  Integer outputs are bit-exact; fp32 features, logits and gradients within 1e-5 relative."""
import pytest
import torch

from conftest import sub_state
from oracle import build_oracle as bo
from oracle import mp_oracle as mo

pytestmark = pytest.mark.gpu
T = torch.from_numpy
RTOL = 1e-5


def relclose(a, b, tol=RTOL):
    a, b = a.detach().cpu(), b.detach().cpu()
    return float((a - b).abs().max()) <= tol * float(b.abs().max()) + 1e-7


def _ops():
    from bridged_gnn_b200 import ops
    return ops


def _rand_graph(n, e, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, n, (2, e), generator=g)


# ------------------------------------------------------------------ CSR build / coalesce (bit-exact)
@pytest.mark.parametrize("n,e", [(1, 0), (10, 0), (7, 30), (1000, 20000), (3408, 37522), (100000, 1 << 20)])
def test_edges_to_csr_exact(n, e):
    ops = _ops()
    ei = _rand_graph(n, e, 1)
    rowptr, col, perm, ne = ops.edges_to_csr(ei[0].cuda(), ei[1].cuda(), n)
    assert ne == e
    key = ei[1] * n + ei[0]
    _, order = torch.sort(key, stable=True)
    assert torch.equal(col.cpu().long(), ei[0][order])
    cnt = torch.bincount(ei[1], minlength=n)
    assert torch.equal(rowptr.cpu().long(), torch.cat((torch.zeros(1, dtype=torch.long), cnt.cumsum(0))))
    # perm maps CSR slots back to input edges (ties between duplicate edges may be permuted)
    p = perm.cpu()
    assert torch.equal(ei[0][p], ei[0][order]) and torch.equal(ei[1][p], ei[1][order])
    assert sorted(p.tolist()) == list(range(e)) if e < 50000 else p.unique().numel() == e


@pytest.mark.parametrize("n,e", [(5, 40), (300, 5000), (3408, 60000)])
def test_coalesce_matches_pyg_semantics(n, e):
    ops = _ops()
    ei = _rand_graph(n, e, 2)
    out = ops.coalesce(ei.cuda(), n).cpu()
    assert torch.equal(out, bo.coalesce(ei, n))
    from bridged_gnn_b200.data import to_undirected
    assert torch.equal(to_undirected(ei.cuda(), n).cpu(), mo.to_undirected(ei, n))


@pytest.mark.parametrize("n,e,training", [(1, 0, True), (10, 0, True), (7, 30, True), (1000, 20000, True),
                                          (1024, 30000, False), (3408, 37522, True), (100000, 1 << 20, True)])
def test_graph_prepare_equals_partition_then_csr(n, e, training):
    """bgnn_graph_prepare (self-loop rewrite inside the key construction, keys-only sort, transposed CSR by a stable
    sort on the source bits, slot map, row orders) against the path it replaces: graph_partition's
    add_self_loops(remove_self_loops(.)) on the host, then two full CSR builds -- every array bit for bit."""
    ops = _ops()
    from bridged_gnn_b200.models import graph_partition
    ei = _rand_graph(n, e, 5).cuda()
    if e:
        ei[1, : e // 8] = ei[0, : e // 8]                      # plenty of self loops in the input
        ei[:, e // 8: e // 4] = ei[:, : e // 4 - e // 8]       # and duplicate edges (kept, like the reference)
    cm = torch.zeros(n, dtype=torch.bool, device="cuda")
    cm[: n * 3 // 4] = True
    ref = ops.CSRGraph(graph_partition(ei, cm)[2], n).prepare(training=training)
    g = ops.CSRGraph.prepared(ei, n, rewrite_self_loops=True, training=training)
    assert g.e == ref.e and g.n == ref.n and g.n_rows == ref.n_rows
    assert torch.equal(g.rowptr, ref.rowptr) and torch.equal(g.col, ref.col)
    assert torch.equal(g.order(), ref.order())
    if training:
        assert torch.equal(g.t[0], ref.t[0]) and torch.equal(g.t[1], ref.t[1])
        assert torch.equal(g.csr_to_csc, ref.csr_to_csc)
        assert torch.equal(g.t_order(), ref.t_order())
    else:
        assert g._t is None
    # without the rewrite it is a plain CSR build of the list as given
    g2 = ops.CSRGraph.prepared(ei, n, rewrite_self_loops=False, training=True)
    ref2 = ops.CSRGraph(ei, n).prepare(training=True)
    assert g2.e == ref2.e and torch.equal(g2.rowptr, ref2.rowptr) and torch.equal(g2.col[: g2.e], ref2.col[: g2.e])
    assert torch.equal(g2.t[0], ref2.t[0]) and torch.equal(g2.t[1][: g2.e], ref2.t[1][: g2.e])
    if e:
        bad = ei.clone()
        bad[0, 0] = n
        with pytest.raises(IndexError):
            ops.CSRGraph.prepared(bad, n)


def test_ktgnn_fast_graph_path_matches_partitioned_lists():
    """On the GPU the model builds its graph with ONE library call from data.edge_index; the reference's cached
    edge_index1 / edge_index2 / edge_index attributes are materialised on first access, are graph_partition's, and
    map to the very graph the kernels used."""
    ops = _ops()
    from bridged_gnn_b200.data import Data
    from bridged_gnn_b200.models import KTGNN_no_complement, graph_partition
    n = 3000
    g = torch.Generator().manual_seed(9)
    ei = torch.randint(0, n, (2, 40000), generator=g).cuda()
    cm = (torch.arange(n) < 2000).cuda()
    x = torch.randn(n, 64, generator=g).cuda()
    data = Data(x=x, edge_index=ei, central_mask=cm)
    torch.manual_seed(0)
    model = KTGNN_no_complement(64, 3, 2, 32, root_weight=False, use_bn=True, dim_share=64, dropout=0.0).cuda().train()
    out = model(data)
    assert model._ei is None and model._fast is not None          # nothing was partitioned on the host
    fast = model._fast[0]
    e1, e2, e = graph_partition(ei, cm)
    assert torch.equal(model.edge_index1, e1) and torch.equal(model.edge_index2, e2) and torch.equal(model.edge_index, e)
    assert ops.cached_graph(model.edge_index, n) is fast
    sum(o.sum() for o in out[:3]).backward()
    g_fast = [p.grad.clone() for p in model.parameters()]
    # the same step over the partitioned lists (the path of a CPU-style caller): identical graph, identical numbers
    model.zero_grad()
    model.edge_index = None
    model._ei = (e1, e2, e)
    out2 = model(data)
    assert model._fast is None
    for a, b in zip(out[:3], out2[:3]):
        assert torch.equal(a, b)
    sum(o.sum() for o in out2[:3]).backward()
    for a, p in zip(g_fast, model.parameters()):
        assert torch.equal(a, p.grad)
    with pytest.raises(AttributeError):
        model.edge_index = e


def test_office_undirected_and_partition(office_build, office_mp):
    from bridged_gnn_b200.data import to_undirected
    from bridged_gnn_b200.models import graph_partition
    ei = to_undirected(T(office_build["edge_index"]).cuda(), 3408)
    assert torch.equal(ei.cpu(), T(office_mp["edge_index_undirected"]))
    e1, e2, e = graph_partition(ei, T(office_build["central_mask"]).cuda())
    assert torch.equal(e1.cpu(), T(office_mp["ktgnn.ei1"])) and torch.equal(e2.cpu(), T(office_mp["ktgnn.ei2"]))


# ------------------------------------------------------------------ SpMM
@pytest.mark.parametrize("f", [1, 2, 31, 64, 100, 128, 256, 300, 512, 1685])
@pytest.mark.parametrize("reduce", ["sum", "mean"])
def test_spmm_forward_backward_random(f, reduce):
    ops = _ops()
    n, e = 500, 6000
    ei = _rand_graph(n, e, 3)
    ei[1, ei[1] == 7] = 8                       # node 7 has no incoming edge (empty row)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(n, f, generator=g)
    w = torch.rand(e, generator=g) if reduce == "sum" else None
    xr = x.clone().requires_grad_(True)
    y_ref = mo.spmm(ei, xr, n, reduce, w)
    gout = torch.randn(n, f, generator=g)
    (y_ref * gout).sum().backward()
    graph = ops.CSRGraph(ei.cuda(), n)
    xg = x.cuda().requires_grad_(True)
    y = ops.spmm(graph, xg, reduce, None if w is None else w.cuda())
    assert relclose(y, y_ref)
    assert bool((y[7] == 0).all())
    (y * gout.cuda()).sum().backward()
    assert relclose(xg.grad, xr.grad)


def test_spmm_scales_equal_edge_weights():
    """gather_scale / out_scale factorisation used by GCNConv == explicit per-edge weights."""
    ops = _ops()
    n, e, f = 400, 5000, 64
    ei = _rand_graph(n, e, 5)
    x = torch.randn(n, f)
    a, b = torch.rand(n) + 0.5, torch.rand(n) + 0.5
    graph = ops.CSRGraph(ei.cuda(), n)
    y1 = ops.spmm(graph, x.cuda(), "sum", None, a.cuda(), b.cuda())
    y2 = mo.spmm(ei, x, n, "sum", a[ei[0]] * b[ei[1]])
    assert relclose(y1, y2)


# ------------------------------------------------------------------ fused AdaptedConv aggregation
def _agg_inputs(n, c, e, seed):
    g = torch.Generator().manual_seed(seed)
    ei = _rand_graph(n, e, seed)
    cm = torch.zeros(n, dtype=torch.bool)
    cm[: (2 * n) // 3] = True
    e1, e2, eall = mo.graph_partition(ei, cm)
    Hs, Ht = torch.randn(n, c, generator=g), torch.randn(n, c, generator=g)
    a1, a2 = torch.randn(c, generator=g) * 0.5, torch.randn(c, generator=g) * 0.5
    gout = torch.randn(n, c, generator=g)
    return eall, e1, e2, cm, Hs, Ht, a1, a2, gout


@pytest.mark.parametrize("n,c,e", [(50, 1, 300), (300, 2, 3000), (300, 6, 3000), (500, 62, 6000), (1000, 31, 12000), (1000, 64, 12000), (700, 100, 5000),
                                   (513, 128, 9000), (400, 256, 4000), (200, 300, 2000), (150, 512, 1500)])
def test_gat_aggregate_forward_backward_random(n, c, e):
    ops = _ops()
    eall, e1, e2, cm, Hs, Ht, a1, a2, gout = _agg_inputs(n, c, e, 20 + c)
    leaf = [t.clone().requires_grad_(True) for t in (Hs, Ht, a1, a2)]
    y_ref = mo.adapted_conv_aggregate(leaf[0], leaf[1], e1, e2, cm, leaf[2], leaf[3])
    (y_ref * gout).sum().backward()
    graph = ops.CSRGraph(eall.cuda(), n)
    dl = [t.clone().cuda().requires_grad_(True) for t in (Hs, Ht, a1, a2)]
    y = ops.gat_aggregate(dl[0], dl[1], dl[2], dl[3], graph, cm.to(torch.uint8).cuda(), 0.1)
    assert relclose(y, y_ref)
    (y * gout.cuda()).sum().backward()
    for got, ref, name in zip(dl, leaf, ("Hs", "Ht", "a_t2s", "a_s2t")):
        assert relclose(got.grad, ref.grad, 2e-5), name


@pytest.mark.parametrize("heads,c,n,e", [(3, 2, 1000, 12000), (2, 1, 300, 2000), (3, 4, 700, 9000), (2, 3, 64, 900), (3, 1, 2000, 30000)])
def test_gat_aggregate_heads_equals_one_aggregation_per_head(heads, c, n, e):
    """Several narrow aggregations in one pass: per head identical to the single-head op (forward and gradients)."""
    ops = _ops()
    assert ops.gat_heads_supported(heads, c) and not ops.gat_heads_supported(4, c) and not ops.gat_heads_supported(heads, 5)
    g = torch.Generator().manual_seed(31 + heads * 7 + c)
    ei = _rand_graph(n, e, 40 + c)
    ei = torch.cat((ei, torch.stack((torch.randint(0, n, (400,), generator=g), torch.zeros(400, dtype=torch.long)))), 1)   # a hub row
    cm = torch.zeros(n, dtype=torch.bool)
    cm[: (2 * n) // 3] = True
    _, _, eall = mo.graph_partition(ei, cm)
    graph = ops.CSRGraph(eall.cuda(), n)
    cm8 = cm.to(torch.uint8).cuda()
    f = heads * c
    vals = [torch.randn(n, f, generator=g), torch.randn(n, f, generator=g), torch.randn(f, generator=g) * 0.5,
            torch.randn(f, generator=g) * 0.5]
    gout = torch.randn(n, f, generator=g).cuda()
    one = [t.clone().cuda().requires_grad_(True) for t in vals]
    ys = [ops.gat_aggregate(one[0][:, h * c:(h + 1) * c].contiguous(), one[1][:, h * c:(h + 1) * c].contiguous(),
                            one[2][h * c:(h + 1) * c], one[3][h * c:(h + 1) * c], graph, cm8, 0.1) for h in range(heads)]
    y_ref = torch.cat(ys, 1)
    (y_ref * gout).sum().backward()
    mh = [t.clone().cuda().requires_grad_(True) for t in vals]
    y = ops.gat_aggregate_heads(mh[0], mh[1], mh[2], mh[3], graph, cm8, 0.1, heads)
    assert relclose(y, y_ref.detach().cpu(), 2e-6)
    (y * gout).sum().backward()
    for got, ref, name in zip(mh, one, ("Hs", "Ht", "a_t2s", "a_s2t")):
        assert relclose(got.grad, ref.grad.cpu(), 1e-5), name


def test_ktgnn_classifier_heads_batched_equals_unbatched(monkeypatch):
    """KTGNN_no_complement with a few classes runs its three classifier convs as one multi-head pass; logits and
    gradients equal the conv-by-conv path."""
    ops = _ops()
    from bridged_gnn_b200.data import Data, to_undirected
    from bridged_gnn_b200.models import KTGNN_no_complement
    g = torch.Generator().manual_seed(5)
    n, d, classes = 1500, 32, 3
    ei = to_undirected(_rand_graph(n, 14000, 8).cuda(), n)
    cm = torch.zeros(n, dtype=torch.bool)
    cm[:1000] = True
    x = torch.randn(n, d, generator=g)
    y = torch.randint(0, classes, (n,), generator=g).cuda()
    data = Data(x=x.cuda(), edge_index=ei, central_mask=cm.cuda())
    torch.manual_seed(0)
    model = KTGNN_no_complement(d, classes, 2, 16, root_weight=False, use_bn=True, dim_share=d, dropout=0.0).cuda().train()

    def run():
        model.zero_grad(set_to_none=True)
        lb, lt, ltt, _ = model(data)
        loss = sum(torch.nn.functional.nll_loss(t, y) for t in (lb, lt, ltt))
        loss.backward()
        return [t.detach().clone() for t in (lb, lt, ltt)], [p.grad.detach().clone() for p in model.parameters()]

    calls = {"n": 0}
    real = ops.gat_aggregate_heads

    def counted(*a, **k):
        calls["n"] += 1
        return real(*a, **k)
    monkeypatch.setattr(ops, "gat_aggregate_heads", counted)
    out_b, grads_b = run()
    assert calls["n"] == 1                                    # the batched path was taken
    monkeypatch.setattr(ops, "gat_heads_supported", lambda h, c: False)
    out_u, grads_u = run()
    assert calls["n"] == 1
    for a, b in zip(out_b, out_u):
        assert relclose(a, b.cpu(), 2e-6)
    for a, b in zip(grads_b, grads_u):
        assert relclose(a, b.cpu(), 2e-5)


def test_rows_by_degree_is_a_stable_descending_permutation():
    ops = _ops()
    for n, e in ((1, 0), (7, 30), (1000, 20000), (4096, 100)):
        ei = _rand_graph(n, e, 5 + n) if e else torch.zeros((2, 0), dtype=torch.long)
        graph = ops.CSRGraph(ei.cuda(), n)
        order = graph.order(64).cpu().long()
        deg = (graph.rowptr[1:] - graph.rowptr[:-1]).cpu().long()
        assert sorted(order.tolist()) == list(range(n))
        ref = torch.sort(-deg, stable=True).indices            # descending degree, ties by ascending row id
        assert torch.equal(order, ref)
        hubs = ops.rows_by_degree(graph.rowptr, n, 20).cpu().long()   # only rows with >= 20 edges move to the front
        key = torch.where(deg >= 20, -deg, torch.ones_like(deg))
        assert torch.equal(hubs, torch.sort(key, stable=True).indices)


def test_gat_aggregate_is_independent_of_the_row_order():
    """The degree order only changes which warp processes which row: forward values are bit-identical to the natural
    order (C-ABI entry points without an order), and so are the row gradients except that, with an order, the longest
    rows get a whole warp each in the backward (their edges are split over 4 groups: another fp32 summation order)."""
    from bridged_gnn_b200 import _lib
    ops = _ops()
    lib = _lib.load()
    n, c, e = 2000, 64, 30000
    eall, e1, e2, cm, Hs, Ht, a1, a2, gout = _agg_inputs(n, c, e, 77)
    graph = ops.CSRGraph(eall.cuda(), n)
    cm8 = cm.to(torch.uint8).cuda()
    dl = [t.clone().cuda().requires_grad_(True) for t in (Hs, Ht, a1, a2)]
    y = ops.gat_aggregate(dl[0], dl[1], dl[2], dl[3], graph, cm8, 0.1)
    y.backward(gout.cuda())
    out = torch.empty_like(y)
    rmax, rsum = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
    P = _lib.ptr
    _lib.check(lib.bgnn_gatv2_fwd_f32(P(graph.rowptr), P(graph.col), P(cm8), P(dl[0].detach()), P(dl[1].detach()),
                                      P(dl[2].detach()), P(dl[3].detach()), 0.1, n, c, P(out), P(rmax), P(rsum),
                                      _lib.stream(out.device)))
    assert torch.equal(out, y.detach())
    gHs, gHt = torch.empty_like(out), torch.empty_like(out)
    ga1, ga2 = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    ws = _lib.workspace(lib.bgnn_gatv2_bwd_workspace_bytes(n, graph.e, c), out.device)
    t_rowptr, t_col, _ = graph.t
    _lib.check(lib.bgnn_gatv2_bwd_f32(P(graph.rowptr), P(graph.col), P(t_rowptr), P(t_col), P(graph.csr_to_csc), graph.e,
                                      P(cm8), P(dl[0].detach()), P(dl[1].detach()), P(dl[2].detach()), P(dl[3].detach()),
                                      0.1, n, c, P(out), P(rmax), P(rsum), P(gout.cuda()), P(gHs), P(gHt), P(ga1), P(ga2),
                                      P(ws), ws.numel(), _lib.stream(out.device)))
    assert relclose(gHs, dl[0].grad, 1e-6) and relclose(gHt, dl[1].grad, 1e-6)
    assert relclose(ga1, dl[2].grad, 1e-5) and relclose(ga2, dl[3].grad, 1e-5)
    # run-to-run reproducibility of the ordered path (fixed summation orders everywhere)
    dl2 = [t.clone().cuda().requires_grad_(True) for t in (Hs, Ht, a1, a2)]
    ops.gat_aggregate(dl2[0], dl2[1], dl2[2], dl2[3], graph, cm8, 0.1).backward(gout.cuda())
    assert all(torch.equal(a.grad, b.grad) for a, b in zip(dl, dl2))


@pytest.mark.parametrize("n,c,e,cut", [(2000, 64, 30000, 1100), (900, 32, 9000, 600), (700, 2, 8000, 500), (500, 100, 5000, 333)])
def test_gat_aggregate_destination_partitioned_equals_whole_graph(n, c, e, cut):
    """The multi-GPU layout on one GPU: two 'ranks' own the destination rows [0, cut) and [cut, n) -- local row ids,
    global H / mask / gradient arrays.  Outputs are bit-identical to the whole-graph call row for row; the partial
    gradients of the two ranks add up to the whole-graph gradients.  The second rank owns target-domain rows only when
    cut >= 2n/3, in which case it is given no Hs at all."""
    ops = _ops()
    eall, e1, e2, cm, Hs, Ht, a1, a2, gout = _agg_inputs(n, c, e, 91 + c)
    cm8 = cm.to(torch.uint8).cuda()
    eall = eall.cuda()
    dl = [t.clone().cuda().requires_grad_(True) for t in (Hs, Ht, a1, a2)]
    y = ops.gat_aggregate(dl[0], dl[1], dl[2], dl[3], ops.CSRGraph(eall, n), cm8, 0.1)
    y.backward(gout.cuda())
    tot = [torch.zeros_like(t) for t in dl]
    for lo, hi in ((0, cut), (cut, n)):
        sel = (eall[1] >= lo) & (eall[1] < hi)
        g = ops.CSRGraph(eall[:, sel].contiguous(), n, n_rows=hi - lo, row_off=lo)
        assert g.n_rows == hi - lo and g.rowptr.numel() == hi - lo + 1 and int(g.col.max()) < n
        pl = [t.clone().cuda().requires_grad_(True) for t in (Hs, Ht, a1, a2)]
        only_tar = not bool(cm[lo:hi].any())
        only_src = bool(cm[lo:hi].all())
        yp = ops.gat_aggregate(None if only_tar else pl[0], None if only_src else pl[1], pl[2], pl[3], g, cm8, 0.1)
        assert torch.equal(yp, y.detach()[lo:hi])
        yp.backward(gout.cuda()[lo:hi])
        for k, t in enumerate(pl):
            if t.grad is not None:
                tot[k] += t.grad
            else:
                assert (k == 0 and only_tar) or (k == 1 and only_src)
    for k, name in enumerate(("Hs", "Ht", "a_t2s", "a_s2t")):
        assert relclose(tot[k], dl[k].grad, 2e-5), name


@pytest.mark.parametrize("heads,c,n,e,cut", [(3, 2, 1000, 12000, 400), (2, 4, 600, 7000, 450)])
def test_gat_heads_destination_partitioned_equals_whole_graph(heads, c, n, e, cut):
    ops = _ops()
    eall, e1, e2, cm, _, _, _, _, _ = _agg_inputs(n, c, e, 55 + heads)
    g0 = torch.Generator().manual_seed(heads * 7 + c)
    f = heads * c
    Hs, Ht = torch.randn(n, f, generator=g0), torch.randn(n, f, generator=g0)
    a1, a2, gout = torch.randn(f, generator=g0) * 0.5, torch.randn(f, generator=g0) * 0.5, torch.randn(n, f, generator=g0)
    cm8 = cm.to(torch.uint8).cuda()
    eall = eall.cuda()
    dl = [t.clone().cuda().requires_grad_(True) for t in (Hs, Ht, a1, a2)]
    y = ops.gat_aggregate_heads(dl[0], dl[1], dl[2], dl[3], ops.CSRGraph(eall, n), cm8, 0.1, heads)
    y.backward(gout.cuda())
    tot = [torch.zeros_like(t) for t in dl]
    for lo, hi in ((0, cut), (cut, n)):
        sel = (eall[1] >= lo) & (eall[1] < hi)
        g = ops.CSRGraph(eall[:, sel].contiguous(), n, n_rows=hi - lo, row_off=lo)
        pl = [t.clone().cuda().requires_grad_(True) for t in (Hs, Ht, a1, a2)]
        yp = ops.gat_aggregate_heads(pl[0], pl[1], pl[2], pl[3], g, cm8, 0.1, heads)
        assert torch.equal(yp, y.detach()[lo:hi])
        yp.backward(gout.cuda()[lo:hi])
        for k, t in enumerate(pl):
            tot[k] += t.grad
    for k, name in enumerate(("Hs", "Ht", "a_t2s", "a_s2t")):
        assert relclose(tot[k], dl[k].grad, 2e-5), name


def test_spmm_column_panels_in_place():
    """bgnn_spmm_csr_ld_f32: a column panel of X aggregated into a column panel of Y equals the same columns of the
    full product bit for bit (the panel-pipelined multi-GPU SpMM relies on it)."""
    ops = _ops()
    n, f = 3000, 256
    ei = _rand_graph(n, 40000, 3).cuda()
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(4)).cuda()
    g = ops.CSRGraph(ei, n)
    full = ops.spmm(g, x, "mean")
    y = torch.full((n, f), float("nan"), device="cuda")
    for lo, hi in ((0, 64), (64, 128), (128, 200), (200, 256)):
        ops._spmm_raw(g.rowptr, g.col, x[:, lo:hi], n, True, out=y[:, lo:hi])
    assert torch.equal(y, full)


def test_batch_norm_relu_dist_single_rank_equals_local():
    """The rank-combining BatchNorm path with a 1-rank process group reproduces the local two-pass kernels."""
    import torch.distributed as dist
    ops = _ops()
    if not dist.is_initialized():
        import os
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", str(38000 + os.getpid() % 2000))
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda:0"))
    try:
        g = torch.Generator().manual_seed(8)
        x = (torch.randn(5000, 64, generator=g) * 3 + 1).cuda()
        go = torch.randn(5000, 64, generator=g).cuda()
        bn1, bn2 = torch.nn.BatchNorm1d(64).cuda(), torch.nn.BatchNorm1d(64).cuda()
        with torch.no_grad():
            bn1.weight.uniform_(0.5, 1.5), bn1.bias.uniform_(-0.5, 0.5)
        bn2.load_state_dict(bn1.state_dict())
        x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        y1 = ops.batch_norm_relu(x1, bn1)
        y2 = ops.batch_norm_relu_dist(x2, bn2, None)
        y1.backward(go), y2.backward(go)
        assert relclose(y2, y1, 1e-6) and relclose(x2.grad, x1.grad, 1e-5)
        assert relclose(bn2.weight.grad, bn1.weight.grad, 1e-5) and relclose(bn2.bias.grad, bn1.bias.grad, 1e-5)
        assert relclose(bn2.running_mean, bn1.running_mean, 1e-6) and relclose(bn2.running_var, bn1.running_var, 1e-6)
    finally:
        dist.destroy_process_group()


def test_gat_aggregate_unsupported_width_fails_loudly():
    ops = _ops()
    eall, e1, e2, cm, Hs, Ht, a1, a2, gout = _agg_inputs(40, 1000, 200, 9)
    graph = ops.CSRGraph(eall.cuda(), 40)
    with pytest.raises(RuntimeError, match="unsupported"):
        ops.gat_aggregate(Hs.cuda(), Ht.cuda(), a1.cuda(), a2.cuda(), graph, cm.to(torch.uint8).cuda(), 0.1)


def test_gat_aggregate_rows_without_edges():
    """Rows with no incoming edge (other ranks' rows in the partitioned layout): zero output, zero gradient,
    and the remaining rows unaffected."""
    ops = _ops()
    n, c = 300, 64
    g = torch.Generator().manual_seed(12)
    ei = torch.randint(0, n, (2, 4000), generator=g)
    ei = ei[:, ei[1] < 150]                                  # only the first 150 rows receive edges
    cm = torch.zeros(n, dtype=torch.bool)
    cm[:200] = True
    Hs, Ht = torch.randn(n, c, generator=g), torch.randn(n, c, generator=g)
    a1, a2 = torch.randn(c, generator=g) * 0.3, torch.randn(c, generator=g) * 0.3
    gout = torch.randn(n, c, generator=g)
    leaf = [t.clone().requires_grad_(True) for t in (Hs, Ht, a1, a2)]
    m1 = cm[ei[1]]
    y_ref = mo.adapted_conv_aggregate(leaf[0], leaf[1], ei[:, m1], ei[:, ~m1], cm, leaf[2], leaf[3])
    (y_ref * gout).sum().backward()
    graph = ops.CSRGraph(ei.cuda(), n)
    dl = [t.clone().cuda().requires_grad_(True) for t in (Hs, Ht, a1, a2)]
    y = ops.gat_aggregate(dl[0], dl[1], dl[2], dl[3], graph, cm.to(torch.uint8).cuda(), 0.1)
    assert bool((y[150:] == 0).all()) and relclose(y, y_ref)
    (y * gout.cuda()).sum().backward()
    for got, ref in zip(dl, leaf):
        assert relclose(got.grad, ref.grad, 2e-5)


def test_gat_aggregate_single_edge_rows_and_large_scores():
    """Rows whose only edge is the self loop give out = H[i]; huge scores must not overflow the softmax."""
    ops = _ops()
    n, c = 64, 32
    cm = torch.zeros(n, dtype=torch.bool)
    cm[:40] = True
    e1, e2, eall = mo.graph_partition(torch.zeros((2, 0), dtype=torch.long), cm)
    H = torch.randn(n, c) * 50
    a = torch.randn(c) * 5
    graph = ops.CSRGraph(eall.cuda(), n)
    y = ops.gat_aggregate(H.cuda(), H.cuda() * 2, a.cuda(), a.cuda(), graph, cm.to(torch.uint8).cuda(), 0.1)
    assert torch.allclose(y[:40].cpu(), H[:40], rtol=1e-6, atol=0)
    assert torch.allclose(y[40:].cpu(), 2 * H[40:], rtol=1e-6, atol=0)
    assert bool(torch.isfinite(y).all())


@pytest.mark.parametrize("n,c", [(1, 1), (1000, 2), (777, 31), (5000, 32), (4096, 64), (3000, 100), (2000, 256), (300, 512)])
@pytest.mark.parametrize("with_bias", [False, True])
def test_adapted_transform_forward_backward(n, c, with_bias):
    ops = _ops()
    g = torch.Generator().manual_seed(40 + c)
    P = torch.randn(n, 2 * c + 2, generator=g)
    wd, kg = torch.randn(1, 2 * c, generator=g), torch.randn(2, generator=g)
    bias = torch.randn(2 * c, generator=g)
    cm = torch.rand(n, generator=g) < 0.7
    go_s, go_t = torch.randn(n, c, generator=g), torch.randn(n, c, generator=g)
    vals = [P, wd, kg] + ([bias] if with_bias else [])
    leaf = [t.clone().requires_grad_(True) for t in vals]
    Pb = leaf[0] + torch.cat((leaf[3], torch.zeros(2))) if with_bias else leaf[0]
    Hs_r, Ht_r = mo.adapted_transform_epilogue(Pb, leaf[1], leaf[2], cm.to(torch.uint8))
    ((Hs_r * go_s).sum() + (Ht_r * go_t).sum()).backward()
    dl = [t.clone().cuda().requires_grad_(True) for t in vals]
    Hs, Ht = ops.adapted_transform(dl[0], dl[1], dl[2], cm.to(torch.uint8).cuda(), dl[3] if with_bias else None)
    assert relclose(Hs, Hs_r, 2e-6) and relclose(Ht, Ht_r, 2e-6)
    ((Hs * go_s.cuda()).sum() + (Ht * go_t.cuda()).sum()).backward()
    for got, ref, name in zip(dl, leaf, ("P", "wd", "kg", "bias")):
        assert got.grad.shape == ref.grad.shape and relclose(got.grad, ref.grad, 2e-5), name


@pytest.mark.parametrize("n,k,no", [(1, 4, 1), (127, 32, 32), (128, 128, 130), (1000, 130, 128), (4099, 100, 96),
                                    (20000, 256, 256), (300, 36, 7), (150 * 128 + 5, 64, 64)])
def test_rowpanel_gemm_matches_float64(n, k, no):
    """3 x TF32 row-panel GEMM (tcgen05, the streamed operand split on chip) == float64 product to fp32 accuracy,
    for resident and streamed B planes, ragged n / k / no and more tiles than CTAs."""
    ops = _ops()
    g = torch.Generator().manual_seed(n + k + no)
    A = torch.randn(n, k, generator=g) * torch.exp(torch.randn(n, 1, generator=g))
    B = torch.randn(no, k, generator=g)
    ref = A.double() @ B.double().t()
    scale = (A.double().abs() @ B.double().abs().t()).clamp(min=1e-30)     # the error bound of a dot product
    Y = ops.rowpanel_gemm(A.cuda(), B.cuda())
    assert Y.shape == (n, no)
    bias = torch.randn(no, generator=g)
    assert torch.allclose(ops.rowpanel_gemm(A.cuda(), B.cuda(), bias.cuda()), Y + bias.cuda(), rtol=0, atol=1e-6 * float(Y.abs().max()))
    err = float(((Y.cpu().double() - ref).abs() / scale).max())
    assert err < 2e-6, err
    if k % 4 == 0 and n > 1:      # a row-strided view is read in place
        big = torch.zeros(n, k + 8)
        big[:, :k] = A
        Y2 = ops.rowpanel_gemm(big.cuda()[:, :k], B.cuda())
        assert torch.equal(Y2, Y)


@pytest.mark.parametrize("n,no,d", [(1, 1, 4), (31, 32, 32), (32, 130, 128), (1000, 64, 64), (4099, 130, 128), (70000, 256, 100),
                                    (5000, 7, 36), (148 * 32 * 3 + 17, 96, 128)])
def test_wgrad_gemm_matches_float64(n, no, d):
    """G^T X on tcgen05 (both operands MN-major in shared memory, split into tf32 planes on chip) == float64 product
    to fp32 accuracy; ragged n / no / d, fewer row blocks than CTAs and several per CTA; bit-reproducible."""
    ops = _ops()
    g = torch.Generator().manual_seed(n + no + d)
    G = torch.randn(n, no, generator=g) * torch.exp(torch.randn(n, 1, generator=g))
    X = torch.randn(n, d, generator=g)
    ldg, ldx = no + (-no) % 4 + 4, d + (-d) % 4
    Gd = torch.zeros(n, ldg).cuda()
    Gd[:, :no] = G.cuda()
    Xd = torch.zeros(n, ldx).cuda()
    Xd[:, :d] = X.cuda()
    Gv, Xv = Gd[:, :no], Xd[:, :d]
    assert ops.wgrad_gemm_supported(Gv, Xv)
    W = ops.wgrad_gemm(Gv, Xv)
    ref = G.double().t() @ X.double()
    scale = (G.double().abs().t() @ X.double().abs()).clamp(min=1e-30)
    assert W.shape == (no, d)
    err = float(((W.cpu().double() - ref).abs() / scale).max())
    assert err < 2e-6, err
    assert torch.equal(W, ops.wgrad_gemm(Gv, Xv))
    if d <= 96:                # column sums of G ride along on an all-ones feature
        W2, cs = ops.wgrad_gemm(Gv, Xv, colsum=True)
        assert torch.equal(W2, W)
        ref_cs = G.double().sum(0)
        assert float(((cs.cpu().double() - ref_cs).abs() / G.double().abs().sum(0).clamp(min=1e-30)).max()) < 2e-6


def test_wgrad_gemm_long_reduction_keeps_fp32_accuracy():
    """1e6-row reduction with heavy cancellation (the training shape): the tensor core's fp32 accumulation is not
    round-to-nearest, so the kernel flushes its TMEM accumulator into an fp32 partial every 512 rows; the result has
    to stay at the accuracy of an fp32 FMA GEMM (measured 1.9e-8 of sum |g||x| against cuBLAS' 1.4e-8; one
    accumulator for the whole kernel gave 2.7e-7) and within 1e-5 of max |ref|."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(0)
    n = 1 << 20
    G = torch.randn(n, 64, device="cuda", generator=g) * 1e-3
    X = torch.randn(n, 64, device="cuda", generator=g)
    W = ops.wgrad_gemm(G, X)
    ref = G.double().t() @ X.double()
    scale = G.double().abs().t() @ X.double().abs()
    err = (W.double() - ref).abs()
    assert float((err / scale).max()) < 6e-8
    assert float(err.max() / ref.abs().max()) < 1e-5


def test_wgrad_gemm_rejects_what_tma_cannot_address():
    ops = _ops()
    G, X = torch.randn(64, 10).cuda(), torch.randn(64, 16).cuda()
    assert not ops.wgrad_gemm_supported(G, X)            # 40-byte rows
    assert not ops.wgrad_gemm_supported(X, torch.randn(64, 132).cuda())     # d > 128
    with pytest.raises(ValueError):
        ops.wgrad_gemm(G, X)


@pytest.mark.parametrize("n,din,dout,bias", [(1, 4, 4, True), (1000, 64, 64, True), (4097, 100, 36, False), (20000, 128, 256, True)])
def test_node_linear_forward_backward(n, din, dout, bias):
    ops = _ops()
    g = torch.Generator().manual_seed(n + din)
    x, w, b = torch.randn(n, din, generator=g), torch.randn(dout, din, generator=g) * 0.2, torch.randn(dout, generator=g)
    go = torch.randn(n, dout, generator=g)
    vals = [x, w] + ([b] if bias else [])
    leaf = [t.clone().double().requires_grad_(True) for t in vals]
    y_r = torch.nn.functional.linear(leaf[0], leaf[1], leaf[2] if bias else None)
    (y_r * go).sum().backward()
    dl = [t.clone().cuda().requires_grad_(True) for t in vals]
    assert ops.linear_supported(dl[0], dl[1])
    y = ops.linear(dl[0], dl[1], dl[2] if bias else None)
    assert relclose(y, y_r.float(), 3e-6)
    (y * go.cuda()).sum().backward()
    for got, ref, name in zip(dl, leaf, ("x", "weight", "bias")):
        assert got.grad.shape == ref.grad.shape and relclose(got.grad, ref.grad.float(), 1e-5), name


@pytest.mark.parametrize("n,c,d,heads,bias", [(50, 1, 4, 1, True), (1000, 2, 64, 2, True), (777, 3, 100, 2, False),
                                              (5000, 4, 128, 1, True), (3000, 2, 256, 2, True)])
def test_adapted_skinny_group_forward_backward(n, c, d, heads, bias):
    """Domain means + narrow transform of 1-2 heads over the same x as one autograd node == the oracle's chain
    (means -> Delta -> wd, kg -> epilogue) differentiated by torch in float64, gradient through the means included."""
    ops = _ops()
    g = torch.Generator().manual_seed(300 + n + c)
    x = torch.randn(n, d, generator=g) + 0.5
    cm = torch.rand(n, generator=g) < 0.7
    cm[0], cm[1] = True, False
    inv = torch.tensor([1.0 / float(cm.sum()), 1.0 / float((~cm).sum())])
    hp = []
    for _ in range(heads):
        w = torch.randn(2 * c + 2, d, generator=g) * 0.3
        b = torch.cat((torch.randn(2 * c, generator=g), torch.zeros(2))) if bias else None
        a = torch.randn(2, d, generator=g) * 0.3
        hp.append((w, b, a))
    go_s, go_t = torch.randn(n, heads * c, generator=g), torch.randn(n, heads * c, generator=g)

    xr = x.double().requires_grad_(True)
    hpr = [tuple(None if t is None else t.double().requires_grad_(True) for t in h) for h in hp]
    cf = cm.double()
    means = torch.stack(((cf * inv[0].double()) @ xr, ((1 - cf) * inv[1].double()) @ xr))
    delta = means[0] - means[1]
    outs_s, outs_t = [], []
    for w, b, a in hpr:
        P = xr @ w.t() + (b if b is not None else 0.0)
        Hs_r, Ht_r = mo.adapted_transform_epilogue(P, w[: 2 * c] @ delta, a @ delta, cm.to(torch.uint8))
        outs_s.append(Hs_r)
        outs_t.append(Ht_r)
    Hs_r, Ht_r = torch.cat(outs_s, 1), torch.cat(outs_t, 1)
    ((Hs_r * go_s).sum() + (Ht_r * go_t).sum()).backward()

    xd = x.cuda().requires_grad_(True)
    hpd = [tuple(None if t is None else t.clone().cuda().requires_grad_(True) for t in h) for h in hp]
    assert ops.adapted_skinny_group_supported(xd, c, heads)
    Hs, Ht = ops.adapted_skinny_group(xd, cm.to(torch.uint8).cuda(), inv.cuda(), hpd)
    assert relclose(Hs, Hs_r.float(), 5e-6) and relclose(Ht, Ht_r.float(), 5e-6)
    ((Hs * go_s.cuda()).sum() + (Ht * go_t.cuda()).sum()).backward()
    assert relclose(xd.grad, xr.grad.float(), 2e-5)
    for h in range(heads):
        for got, ref, name in zip(hpd[h], hpr[h], ("w_cat", "b_cat", "a_tail")):
            if got is None:
                continue
            rg = ref.grad.float()
            if name == "b_cat":
                rg = rg.clone()
                rg[2 * c:] = 0.0
            assert relclose(got.grad, rg, 2e-5), (h, name)


@pytest.mark.parametrize("n,c", [(16, 4), (1000, 64), (4099, 100), (50000, 64), (3000, 256), (700, 1024)])
@pytest.mark.parametrize("relu", [True, False])
def test_batch_norm_relu_matches_torch(n, c, relu):
    """Fused BatchNorm1d (+ ReLU), training and inference mode, against nn.BatchNorm1d / F.relu in float64: outputs,
    all three gradients, running statistics; the backward recomputes the ReLU mask from x."""
    ops = _ops()
    g = torch.Generator().manual_seed(n + c)
    x = torch.randn(n, c, generator=g) * (1.0 + torch.rand(c, generator=g)) + 3.0 * torch.randn(c, generator=g)
    go = torch.randn(n, c, generator=g)
    ref = torch.nn.BatchNorm1d(c).double()
    with torch.no_grad():
        ref.weight.copy_(torch.randn(c, generator=g))
        ref.bias.copy_(torch.randn(c, generator=g))
    bn = torch.nn.BatchNorm1d(c).cuda()
    bn.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    xr = x.double().requires_grad_(True)
    yr = ref(xr)
    yr = torch.relu(yr) if relu else yr
    (yr * go).sum().backward()
    xd = x.cuda().requires_grad_(True)
    assert ops.batch_norm_relu_supported(xd, bn)
    y = ops.batch_norm_relu(xd, bn, relu=relu)
    (y * go.cuda()).sum().backward()
    assert relclose(y, yr.float(), 3e-6)
    assert relclose(xd.grad, xr.grad.float(), 2e-5)
    assert relclose(bn.weight.grad, ref.weight.grad.float(), 2e-5) and relclose(bn.bias.grad, ref.bias.grad.float(), 2e-5)
    assert relclose(bn.running_mean, ref.running_mean.float(), 2e-6) and relclose(bn.running_var, ref.running_var.float(), 2e-6)
    assert int(bn.num_batches_tracked) == 1
    y2 = ops.batch_norm_relu(xd, bn, relu=relu)          # deterministic
    assert torch.equal(y2, y)
    with torch.no_grad():
        ref(x.double())                                  # second momentum update on the reference side too
    bn.eval()
    ref.eval()
    with torch.no_grad():
        assert ops.batch_norm_relu_supported(xd, bn)
        ye = ops.batch_norm_relu(xd, bn, relu=relu)
        yre = ref(x.double())
        yre = torch.relu(yre) if relu else yre
    assert relclose(ye, yre.float(), 3e-6)
    assert not ops.batch_norm_relu_supported(xd, bn)     # inference statistics with autograd on: left to torch


@pytest.mark.parametrize("n,c,d,bias", [(1, 32, 4, True), (1000, 64, 128, True), (777, 32, 100, False), (5000, 64, 64, True),
                                        (20000, 96, 256, True), (129, 64, 36, False)])
def test_adapted_wide_forward_backward(n, c, d, bias):
    """Wide-output transform straight from x (tensor-core contraction with the node-wise epilogue fused) == dense
    contraction + node-wise epilogue of the oracle."""
    ops = _ops()
    assert ops.adapted_wide_supported(c, d) and not ops.adapted_wide_supported(c + 1, d) and not ops.adapted_wide_supported(160, d)
    g = torch.Generator().manual_seed(90 + c + d)
    x = torch.randn(n, d, generator=g)
    w_cat = torch.randn(2 * c + 2, d, generator=g) * 0.3
    b2 = torch.randn(2 * c, generator=g) if bias else None
    wd, kg = torch.randn(1, 2 * c, generator=g), torch.randn(2, generator=g)
    cm = torch.rand(n, generator=g) < 0.7
    go_s, go_t = torch.randn(n, c, generator=g), torch.randn(n, c, generator=g)
    names = ["x", "w_cat", "wd", "kg"] + (["bias"] if bias else [])
    vals = [x, w_cat, wd, kg] + ([b2] if bias else [])
    leaf = [t.clone().double().requires_grad_(True) for t in vals]
    P = leaf[0] @ leaf[1].t() + (torch.cat((leaf[4], torch.zeros(2, dtype=torch.float64))) if bias else 0.0)
    Hs_r, Ht_r = mo.adapted_transform_epilogue(P, leaf[2], leaf[3], cm.to(torch.uint8))
    ((Hs_r * go_s).sum() + (Ht_r * go_t).sum()).backward()
    dl = [t.clone().cuda().requires_grad_(True) for t in vals]
    Hs, Ht = ops.adapted_wide(dl[0], dl[1], dl[4] if bias else None, dl[2], dl[3], cm.to(torch.uint8).cuda())
    assert relclose(Hs, Hs_r.float(), 3e-6) and relclose(Ht, Ht_r.float(), 3e-6)
    ((Hs * go_s.cuda()).sum() + (Ht * go_t.cuda()).sum()).backward()
    for got, ref, name in zip(dl, leaf, names):
        assert got.grad.shape == ref.grad.shape and relclose(got.grad, ref.grad.float(), 2e-5), name
    # first-layer case: x is data (no gradient) -- dP is never materialised, the weight-gradient GEMM reads
    # [dHs | dHt | d gates] in place
    dl2 = [t.clone().cuda().requires_grad_(i > 0) for i, t in enumerate(vals)]
    Hs2, Ht2 = ops.adapted_wide(dl2[0], dl2[1], dl2[4] if bias else None, dl2[2], dl2[3], cm.to(torch.uint8).cuda())
    assert torch.equal(Hs2, Hs) and torch.equal(Ht2, Ht)
    ((Hs2 * go_s.cuda()).sum() + (Ht2 * go_t.cuda()).sum()).backward()
    assert dl2[0].grad is None
    for got, ref, name in list(zip(dl2, leaf, names))[1:]:
        assert got.grad.shape == ref.grad.shape and relclose(got.grad, ref.grad.float(), 2e-5), name


def test_tf32_planes_kernel_equals_host_restatement():
    """The one-launch operand preparation == the torch restatement (tests/test_host_logic.py pins that one), bit for
    bit, for contiguous and transposed inputs."""
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    w = torch.randn(130, 100, generator=g) * torch.exp(3 * torch.randn(130, 1, generator=g))
    for view_cpu, view_gpu in ((w, w.cuda()), (w.t(), w.cuda().t())):
        hi_r, lo_r = ops.tf32_planes(view_cpu)
        hi, lo = ops.tf32_planes(view_gpu)
        assert hi.shape == hi_r.shape and torch.equal(hi.cpu(), hi_r) and torch.equal(lo.cpu(), lo_r)


def test_wgrad_gemm_cat_equals_concatenation():
    ops = _ops()
    g = torch.Generator().manual_seed(11)
    n = 3000
    A, B, C = torch.randn(n, 64, generator=g).cuda(), torch.randn(n, 32, generator=g).cuda(), torch.randn(n, 4, generator=g).cuda()
    X = torch.randn(n, 100, generator=g).cuda()
    W = ops.wgrad_gemm_cat([A, B, C[:, :2]], X)
    ref = torch.cat((A, B, C[:, :2]), 1).double().t() @ X.double()
    assert W.shape == (98, 100) and relclose(W, ref.float(), 2e-6)
    assert relclose(ops.wgrad_gemm_cat([A, B], X), ref[:96].float(), 2e-6)
    assert not ops.wgrad_gemm_cat_supported([C[:, :2], A], X)          # only the last block may be ragged


@pytest.mark.parametrize("n,c,d,bias", [(1, 1, 4, True), (1000, 2, 64, True), (777, 3, 100, False), (5000, 4, 128, True),
                                        (3000, 2, 256, True), (257, 4, 8, False)])
def test_adapted_skinny_forward_backward(n, c, d, bias):
    """Narrow-output transform straight from x == dense contraction + node-wise epilogue of the oracle."""
    ops = _ops()
    assert ops.adapted_skinny_supported(c, d) and not ops.adapted_skinny_supported(5, d) and not ops.adapted_skinny_supported(c, 260)
    g = torch.Generator().manual_seed(70 + c + d)
    x = torch.randn(n, d, generator=g)
    w_cat = torch.randn(2 * c + 2, d, generator=g) * 0.3
    b_cat = torch.cat((torch.randn(2 * c, generator=g), torch.zeros(2))) if bias else None
    wd, kg = torch.randn(1, 2 * c, generator=g), torch.randn(2, generator=g)
    cm = torch.rand(n, generator=g) < 0.7
    go_s, go_t = torch.randn(n, c, generator=g), torch.randn(n, c, generator=g)
    names = ["x", "w_cat", "wd", "kg"] + (["b_cat"] if bias else [])
    vals = [x, w_cat, wd, kg] + ([b_cat] if bias else [])
    leaf = [t.clone().requires_grad_(True) for t in vals]
    P = leaf[0] @ leaf[1].t() + (leaf[4] if bias else 0.0)
    Hs_r, Ht_r = mo.adapted_transform_epilogue(P, leaf[2], leaf[3], cm.to(torch.uint8))
    ((Hs_r * go_s).sum() + (Ht_r * go_t).sum()).backward()
    dl = [t.clone().cuda().requires_grad_(True) for t in vals]
    Hs, Ht = ops.adapted_skinny(dl[0], dl[1], dl[4] if bias else None, dl[2], dl[3], cm.to(torch.uint8).cuda())
    assert relclose(Hs, Hs_r, 5e-6) and relclose(Ht, Ht_r, 5e-6)
    ((Hs * go_s.cuda()).sum() + (Ht * go_t.cuda()).sum()).backward()
    for got, ref, name in zip(dl, leaf, names):
        if name == "b_cat":      # the two gate columns have no bias parameter: their gradient is reported as 0
            assert relclose(got.grad[: 2 * c], ref.grad[: 2 * c], 2e-5) and bool((got.grad[2 * c:] == 0).all())
        else:
            assert got.grad.shape == ref.grad.shape and relclose(got.grad, ref.grad, 2e-5), name


@pytest.mark.parametrize("n,d", [(1, 4), (1000, 64), (4097, 256), (300, 1024)])
def test_domain_means_forward_backward(n, d):
    ops = _ops()
    g = torch.Generator().manual_seed(90 + d)
    x = torch.randn(n, d, generator=g)
    cm = torch.rand(n, generator=g) < 0.6
    if n > 1:
        cm[0], cm[1] = True, False
    ns, nt = cm.sum().clamp(min=1).float(), (n - cm.sum()).clamp(min=1).float()
    gm = torch.randn(2, d, generator=g)
    xr = x.clone().requires_grad_(True)
    ref = torch.stack(((xr * cm.float().unsqueeze(1)).sum(0) / ns, (xr * (~cm).float().unsqueeze(1)).sum(0) / nt))
    (ref * gm).sum().backward()
    xd = x.clone().cuda().requires_grad_(True)
    inv = torch.stack((1.0 / ns, 1.0 / nt)).cuda()
    got = ops.domain_means(xd, cm.to(torch.uint8).cuda(), inv)
    assert relclose(got, ref, 2e-6)
    (got * gm.cuda()).sum().backward()
    assert relclose(xd.grad, xr.grad, 2e-6)


# ------------------------------------------------------------------ diagnostics (utils.py:101-131) vs the dense restatement
@pytest.mark.parametrize("n,e,classes", [(60, 300, 2), (500, 6000, 5), (1500, 9000, 31)])
def test_homophily_diagnostics_match_the_dense_reference(n, e, classes):
    from bridged_gnn_b200.data import Data
    from bridged_gnn_b200.utils import eval_bridged_Graph, eval_homophily
    g = torch.Generator().manual_seed(3 + n)
    ei = torch.randint(0, n, (2, e), generator=g)
    ei = torch.cat((ei, ei[:, :20]), 1)                       # a few duplicate edges
    y = torch.randint(0, classes, (n,), generator=g)
    y[torch.rand(n, generator=g) < 0.2] = -1                  # unlabelled nodes
    test_mask = torch.rand(n, generator=g) < 0.4
    r_ref, _ = mo.eval_bridged_graph(ei, y, test_mask, n)
    h1_ref, h2_ref = mo.eval_homophily(ei, y, n)
    data = Data(x=torch.zeros(n, 4).cuda(), edge_index=ei.cuda(), y=y.cuda(), test_mask=test_mask.cuda())
    r = eval_bridged_Graph(data, verbose=False)
    assert abs(float(r) - r_ref) < 1e-6
    for max_pairs in (1 << 26, 500):                          # one block / many blocks of start nodes
        h1, h2 = eval_homophily(data, verbose=False, max_pairs=max_pairs)
        assert abs(h1 - h1_ref) < 1e-6 and abs(h2 - h2_ref) < 1e-6


def test_homophily_diagnostics_on_the_office_graph_match_the_reference(diag_golden, office_build):
    from bridged_gnn_b200.data import Data
    from bridged_gnn_b200.utils import eval_bridged_Graph, eval_homophily
    g = office_build
    data = Data(x=T(g["x"]).cuda(), edge_index=T(g["edge_index"]).cuda(), y=T(g["y"]).cuda(), test_mask=T(g["test_mask"]).cuda())
    assert abs(float(eval_bridged_Graph(data, verbose=False)) - float(diag_golden["office.local_ratio"])) < 1e-6
    h1, h2 = eval_homophily(data, verbose=False)
    assert abs(h1 - float(diag_golden["office.h1"])) < 1e-6 and abs(h2 - float(diag_golden["office.h2"])) < 1e-6


# ------------------------------------------------------------------ golden: reference layers on the office bridged graph
def test_adapted_conv_module_matches_reference(office_mp, office_build):
    from bridged_gnn_b200.models import AdaptedConv
    m = office_mp
    conv = AdaptedConv(64, 31, root_weight=False)
    conv.load_state_dict(sub_state(m, "conv.sd."))
    conv.cuda()
    c = T(office_build["central_mask"]).cuda()
    e1, e2 = T(m["ktgnn.ei1"]).cuda(), T(m["ktgnn.ei2"]).cuda()
    x = T(m["conv.x"]).cuda().requires_grad_(True)
    y = conv(x, torch.cat((e1, e2), 1), e1, e2, c)
    assert relclose(y, T(m["conv.y"]))
    (y * T(m["conv.gout"]).cuda()).sum().backward()
    assert relclose(x.grad, T(m["conv.gx"]), 2e-5)
    for k, p in conv.named_parameters():
        assert relclose(p.grad, T(m["conv.grad." + k]), 2e-5), k


def _office_data(office_build, office_mp):
    from bridged_gnn_b200.data import Data
    return Data(x=T(office_build["x"]), edge_index=T(office_mp["edge_index_undirected"]), y=T(office_build["y"]),
                central_mask=T(office_build["central_mask"]), train_mask=T(office_build["train_mask"])).to("cuda:0")


def test_ktgnn_model_eval_and_train_match_reference(office_mp, office_build):
    """Config 2: KT-GNN 2-layer hidden=64 on the office A->D bridged graph, to_undirected."""
    from bridged_gnn_b200.models import KTGNN_no_complement
    m = office_mp
    data = _office_data(office_build, office_mp)
    model = KTGNN_no_complement(256, 31, 2, 64, root_weight=False, use_bn=True, dim_share=256, need_complement=False)
    model.load_state_dict(sub_state(m, "ktgnn.sd."))
    model.cuda().eval()
    with torch.no_grad():
        lb, lt, ltt, _ = model(data)
    assert relclose(lb, T(m["ktgnn.eval.logp_base"]))
    assert relclose(lt, T(m["ktgnn.eval.logp_target"]))
    assert relclose(ltt, T(m["ktgnn.eval.logp_trans"]))
    assert model.edge_index1.shape[1] == 25055 and model.edge_index2.shape[1] == 12467
    model.train()
    model.dropout = 0.0
    model.zero_grad()
    model.edge_index = None
    model.prepare_graph(data)             # partition + CSR / transposed CSR / row orders ahead of the forward
    graph = _ops().cached_graph(model.edge_index, data.x.shape[0])
    assert graph._t is not None and graph._csr_to_csc is not None and graph._order is not None
    lb, lt, ltt, _ = model(data)
    assert _ops().cached_graph(model.edge_index, data.x.shape[0]) is graph
    tm = data.train_mask
    nll = torch.nn.functional.nll_loss
    loss = nll(lb[tm], data.y[tm]) + nll(lt[tm], data.y[tm]) + nll(ltt[tm], data.y[tm])
    assert abs(loss.item() - float(m["ktgnn.train.loss"])) < 1e-5 * max(1.0, abs(float(m["ktgnn.train.loss"])))
    loss.backward()
    for k, p in model.named_parameters():
        assert relclose(p.grad, T(m["ktgnn.train.grad." + k]), 5e-5), k


def test_sage_gcn_models_match_reference(office_mp, office_build):
    import types
    from bridged_gnn_b200.models import GCNNet, GraphSAGE
    from bridged_gnn_b200.models.models import GraphEncoder
    m = office_mp
    data = _office_data(office_build, office_mp)
    ds = types.SimpleNamespace(num_features=256, num_classes=31)
    sage = GraphSAGE(ds, layer_num=2, hidden=64)
    sage.load_state_dict(sub_state(m, "sage.sd."))
    gcn = GCNNet(ds, layer_num=2, hidden=64)
    gcn.load_state_dict(sub_state(m, "gcn.sd."))
    enc = GraphEncoder(256, 64, dim_hidden=64, layer_num=2, norm_mode="None")
    enc.load_state_dict(sub_state(m, "enc.sd."))
    with torch.no_grad():
        assert relclose(sage.cuda().eval()(data), T(m["sage.logp"]))
        assert relclose(gcn.cuda().eval()(data), T(m["gcn.logp"]))
        assert relclose(enc.cuda().eval()(data.x, data.edge_index), T(m["enc.z"]))


def test_sage_backward_matches_oracle(office_mp, office_build):
    import types
    from bridged_gnn_b200.models import GraphSAGE
    m = office_mp
    data = _office_data(office_build, office_mp)
    ds = types.SimpleNamespace(num_features=256, num_classes=31)
    sage = GraphSAGE(ds, layer_num=2, hidden=64)
    sd = sub_state(m, "sage.sd.")
    sage.load_state_dict(sd)
    sage.cuda().train()
    for mod in sage.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x, ei = T(office_build["x"]), T(m["edge_index_undirected"])
    # oracle in train mode without dropout == eval graph
    lp_ref = mo.graphsage(x, ei, P)
    tm, y = T(office_build["train_mask"]), T(office_build["y"])
    torch.nn.functional.nll_loss(lp_ref[tm], y[tm]).backward()
    sage.eval()      # dropout off; no BN in GraphSAGE, so eval == train numerically
    lp = sage(data)
    torch.nn.functional.nll_loss(lp[data.train_mask], data.y[data.train_mask]).backward()
    for k, p in sage.named_parameters():
        assert relclose(p.grad, P[k].grad, 2e-5), k


def test_gcn_backward_matches_oracle(office_mp, office_build):
    """GCNNet training gradients (scaled SpMM backward over the transposed CSR, NodeLinear weight gradients) against
    autograd through the oracle's restatement of GCNConv (models/backbones.py:269-277, PyG gcn_norm)."""
    import types
    from bridged_gnn_b200.models import GCNNet
    m = office_mp
    data = _office_data(office_build, office_mp)
    ds = types.SimpleNamespace(num_features=256, num_classes=31)
    gcn = GCNNet(ds, layer_num=2, hidden=64)
    sd = sub_state(m, "gcn.sd.")
    gcn.load_state_dict(sd)
    gcn.cuda().eval()        # dropout off; no BatchNorm in GCNNet, so eval == train numerically
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x, ei = T(office_build["x"]), T(m["edge_index_undirected"])
    lp_ref = mo.gcn_net(x, ei, P)
    tm, y = T(office_build["train_mask"]), T(office_build["y"])
    torch.nn.functional.nll_loss(lp_ref[tm], y[tm]).backward()
    lp = gcn(data)
    assert relclose(lp, lp_ref.detach())
    torch.nn.functional.nll_loss(lp[data.train_mask], data.y[data.train_mask]).backward()
    for k, p in gcn.named_parameters():
        assert relclose(p.grad, P[k].grad, 2e-5), k


# ------------------------------------------------------------------ full-size properties
def test_large_graph_properties():
    """Config-4-sized aggregation (2^20 nodes, ~1.6e7 edges): constant features must aggregate to the
    constant (softmax weights sum to one; mean of ones is one), and SpMM is linear."""
    ops = _ops()
    n, c = 1 << 20, 64
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    ei = torch.randint(0, n, (2, 15 * n), generator=g, device=dev)
    loops = torch.arange(n, device=dev).unsqueeze(0).repeat(2, 1)
    ei = torch.cat((ei, loops), 1)
    graph = ops.CSRGraph(ei, n)
    rp = graph.rowptr.long()
    assert int(rp[-1]) == ei.shape[1] and bool((rp[1:] >= rp[:-1]).all())
    assert torch.equal(torch.bincount(ei[1], minlength=n), rp[1:] - rp[:-1])
    cm = (torch.arange(n, device=dev) % 4 != 0)
    ones = torch.ones(n, c, device=dev)
    a = torch.randn(c, generator=g, device=dev)
    y = ops.gat_aggregate(ones * 0.5, ones * 2.0, a, -a, graph, cm.to(torch.uint8), 0.1)
    assert torch.allclose(y[cm], ones[cm] * 0.5, rtol=2e-6) and torch.allclose(y[~cm], ones[~cm] * 2.0, rtol=2e-6)
    x1, x2 = torch.randn(n, c, generator=g, device=dev), torch.randn(n, c, generator=g, device=dev)
    lhs = ops.spmm(graph, x1 + 2 * x2, "mean")
    rhs = ops.spmm(graph, x1, "mean") + 2 * ops.spmm(graph, x2, "mean")
    assert float((lhs - rhs).abs().max()) < 1e-4
    assert torch.allclose(ops.spmm(graph, ones, "mean"), ones, rtol=1e-6)
    deg = ops.spmm(graph, ones[:, :1].contiguous(), "sum").view(-1)
    assert torch.equal(deg.long(), rp[1:] - rp[:-1])


def test_train_gnn_entry_point_learns(office_mp, office_build):
    """Config 2 through the reference-shaped entry point (main_graph_knowledge_transfer.train_gnn): KT-GNN
    2-layer hidden 64 on the office A->D bridged graph, undirected; 40 epochs must fit the training split."""
    from bridged_gnn_b200.data import Data
    from bridged_gnn_b200.main_graph_knowledge_transfer import test as evaluate
    from bridged_gnn_b200.main_graph_knowledge_transfer import train_gnn
    g = office_build
    torch.manual_seed(0)
    data = Data(x=T(g["x"]), edge_index=T(office_mp["edge_index_undirected"]), y=T(g["y"]), central_mask=T(g["central_mask"]),
                train_mask=T(g["train_mask"]), val_mask=T(g["val_mask"]), test_mask=T(g["test_mask"]))
    model, best = train_gnn(data, gnn="KTGNN", num_layer=2, hidden=64, num_epoch=40, verbose=False)
    tr, va, te = evaluate(data, model, gnn="KTGNN")
    assert tr > 0.8 and te > 0.5, (tr, va, te)
    model2, best2 = train_gnn(data, gnn="GraphSAGE", num_layer=2, hidden=64, num_epoch=30, verbose=False)
    assert evaluate(data, model2, gnn="GraphSAGE")[0] > 0.8
