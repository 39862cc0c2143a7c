"""CPU tests of the host-side logic around the kernels.  The CUDA operators are replaced by the oracle
here (test infrastructure only) so that the algebra of the Python layers -- the restructured node-wise
part of AdaptedConv, the BatchNorm fold of the mlp similarity head, model wiring and state_dict
compatibility -- is checked against the reference-generated golden vectors without a GPU."""
import types

import pytest
import torch

from conftest import sub_state
from oracle import build_oracle as bo
from oracle import mp_oracle as mo

T = torch.from_numpy


def relclose(a, b, tol=1e-5):
    a, b = a.detach(), b.detach()
    return float((a - b).abs().max()) <= tol * float(b.abs().max()) + 1e-7


class _FakeGraph:
    def __init__(self, edge_index, n):
        self.edge_index, self.n = edge_index, n
        self.deg = torch.bincount(edge_index[1], minlength=n).to(torch.float32)


@pytest.fixture()
def oracle_backed_ops(monkeypatch):
    """ops.gat_aggregate / ops.spmm / ops.cached_graph -> oracle implementations on the CPU."""
    from bridged_gnn_b200 import ops

    def cached_graph(edge_index, n):
        return _FakeGraph(edge_index, n)

    def gat_aggregate(Hs, Ht, a1, a2, graph, dst_is_src, slope=0.1):
        cm = dst_is_src.bool()
        ei = graph.edge_index
        m1 = cm[ei[1]]
        return mo.adapted_conv_aggregate(Hs, Ht, ei[:, m1], ei[:, ~m1], cm, a1.view(-1), a2.view(-1), slope)

    def spmm(graph, X, reduce="sum", edge_weight=None, gather_scale=None, out_scale=None):
        ei = graph.edge_index
        w = edge_weight
        if gather_scale is not None or out_scale is not None:
            w = torch.ones(ei.shape[1]) if w is None else w
            if gather_scale is not None:
                w = w * gather_scale[ei[0]]
            if out_scale is not None:
                w = w * out_scale[ei[1]]
        return mo.spmm(ei, X, graph.n, reduce, w)

    monkeypatch.setattr(ops, "adapted_transform", mo.adapted_transform_epilogue)
    monkeypatch.setattr(ops, "cached_graph", cached_graph)
    monkeypatch.setattr(ops, "gat_aggregate", gat_aggregate)
    monkeypatch.setattr(ops, "spmm", spmm)
    monkeypatch.setattr(ops, "CSRGraph", _FakeGraph)
    return ops


def test_adapted_conv_algebra_matches_reference(oracle_backed_ops, office_mp, office_build):
    from bridged_gnn_b200.models import AdaptedConv
    m = office_mp
    conv = AdaptedConv(64, 31, root_weight=False)
    conv.load_state_dict(sub_state(m, "conv.sd."))
    c = T(office_build["central_mask"])
    e1, e2 = T(m["ktgnn.ei1"]), T(m["ktgnn.ei2"])
    x = T(m["conv.x"]).clone().requires_grad_(True)
    y = conv(x, torch.cat((e1, e2), 1), e1, e2, c)
    assert relclose(y, T(m["conv.y"]))
    (y * T(m["conv.gout"])).sum().backward()
    assert relclose(x.grad, T(m["conv.gx"]), 2e-5)
    for k, p in conv.named_parameters():
        assert relclose(p.grad, T(m["conv.grad." + k]), 2e-5), k


def test_ktgnn_wiring_matches_reference(oracle_backed_ops, office_mp, office_build):
    from bridged_gnn_b200.data import Data
    from bridged_gnn_b200.models import KTGNN_no_complement
    m = office_mp
    data = Data(x=T(office_build["x"]), edge_index=T(m["edge_index_undirected"]), y=T(office_build["y"]),
                central_mask=T(office_build["central_mask"]), train_mask=T(office_build["train_mask"]))
    model = KTGNN_no_complement(256, 31, 2, 64, root_weight=False, use_bn=True, dim_share=256, need_complement=False)
    model.load_state_dict(sub_state(m, "ktgnn.sd."))
    model.eval()
    with torch.no_grad():
        lb, lt, ltt, _ = model(data)
    assert relclose(lb, T(m["ktgnn.eval.logp_base"])) and relclose(lt, T(m["ktgnn.eval.logp_target"]))
    assert relclose(ltt, T(m["ktgnn.eval.logp_trans"]))
    assert torch.equal(model.edge_index1, T(m["ktgnn.ei1"])) and torch.equal(model.edge_index2, T(m["ktgnn.ei2"]))
    model.train()
    model.dropout = 0.0
    lb, lt, ltt, _ = model(data)
    tm = data.train_mask
    nll = torch.nn.functional.nll_loss
    loss = nll(lb[tm], data.y[tm]) + nll(lt[tm], data.y[tm]) + nll(ltt[tm], data.y[tm])
    assert abs(loss.item() - float(m["ktgnn.train.loss"])) < 2e-5
    loss.backward()
    for k, p in model.named_parameters():
        assert relclose(p.grad, T(m["ktgnn.train.grad." + k]), 5e-5), k


def test_sage_gcn_wiring_matches_reference(oracle_backed_ops, office_mp, office_build):
    from bridged_gnn_b200.data import Data
    from bridged_gnn_b200.models import GCNNet, GraphSAGE
    m = office_mp
    data = Data(x=T(office_build["x"]), edge_index=T(m["edge_index_undirected"]))
    ds = types.SimpleNamespace(num_features=256, num_classes=31)
    sage = GraphSAGE(ds, layer_num=2, hidden=64)
    sage.load_state_dict(sub_state(m, "sage.sd."))
    gcn = GCNNet(ds, layer_num=2, hidden=64)
    gcn.load_state_dict(sub_state(m, "gcn.sd."))
    with torch.no_grad():
        assert relclose(sage.eval()(data), T(m["sage.logp"]))
        assert relclose(gcn.eval()(data), T(m["gcn.logp"]))


def test_mlp_head_fold_matches_pair_path(office_build):
    """Eval-mode BatchNorm fold of Similar_v2(mode='mlp'): sum_h w2 relu(U_db + U_q) + b2 == the reference's
    pair path logits (to fp32 rounding), on real checkpoint weights and embeddings."""
    from bridged_gnn_b200.models import Similar_v2
    g = office_build
    W = sub_state(g, "ckpt.")
    head = Similar_v2(128, 31, mode="mlp")
    head.load_state_dict({k[len("source_learner.sim_net."):]: v for k, v in W.items() if k.startswith("source_learner.sim_net.")})
    head.eval()
    z_src, z_tar = T(g["z_src"]), T(g["z_tar"])
    U_db, U_q, w2, b2 = head.mlp_operands(z_src, z_tar)
    rows = T(g["sim_rows_idx"])
    logit = (torch.relu(U_q[rows][:, None, :] + U_db[None, :, :]) * w2).sum(-1) + b2
    sim = torch.sigmoid(logit)
    assert float((sim - T(g["sim_rows"])).abs().max()) < 2e-6
    with pytest.raises(RuntimeError):
        head.train().mlp_operands(z_src, z_tar)


def test_checkpoints_load_strict_shapes(office_build, fb_build):
    """Module trees mirror the reference's: every golden checkpoint tensor finds its parameter."""
    from bridged_gnn_b200.data import Data
    from bridged_gnn_b200.models import Adversarial_Learner, Adversarial_Learner_v2
    g = office_build
    d = Data(x=torch.zeros(4, 256), y=torch.tensor([0, 30, 1, 2]))
    m2 = Adversarial_Learner_v2(d, d, dim_hidden=128, num_layer=2, source_clf=True, use_norm=True, norm_mode="None",
                                norm_scale=1.0, backbone="mlp", sim_mode="mlp")
    res = m2.load_state_dict(sub_state(g, "ckpt."), strict=False)
    assert not res.unexpected_keys
    assert all(k.startswith(("target_learner.decoder", "discriminator")) for k in res.missing_keys)
    z = m2.eval().embed_source(Data(x=T(g["x"])[:2817]))
    assert relclose(z, T(g["z_src"]), 1e-5)
    z = m2.embed_target(Data(x=T(g["x"])[2817:]))
    assert relclose(z, T(g["z_tar"]), 1e-5)
    d1 = Data(x=torch.zeros(4, 1685), y=torch.tensor([0, 1, 1, 0]))
    m1 = Adversarial_Learner(d1, d1, dim_hidden=64, num_layer=2, source_clf=True, norm_mode="None", norm_scale=1.0)
    res = m1.load_state_dict(sub_state(fb_build, "ckpt."), strict=False)
    assert not res.unexpected_keys
    u = m1.eval().source_learner.sim_net.cosine_operand(T(fb_build["z_src"]))
    assert u.shape == (1200, 128)


def test_graph_partition_and_self_loops(office_mp, office_build):
    from bridged_gnn_b200.models import graph_partition
    e1, e2, e = graph_partition(T(office_mp["edge_index_undirected"]), T(office_build["central_mask"]))
    assert torch.equal(e1, T(office_mp["ktgnn.ei1"])) and torch.equal(e2, T(office_mp["ktgnn.ei2"]))
    assert e.shape[1] == 37522


def test_reorder_matches_reference_semantics():
    """§8(f) row 2: reorder (main_bridged_graph.py:195-222) is pure index logic -- run on the CPU here against the oracle
    restatement (pinned on the reference's output in test_oracle_golden.py; the device run against the golden file is
    tests/test_gpu_assemble.py).  The edge-validity filters have no CPU path: CPU tensors raise."""
    from bridged_gnn_b200.data import Data
    from bridged_gnn_b200 import main_bridged_graph as mb
    gen = torch.Generator().manual_seed(0)
    perm = torch.randperm(40, generator=gen).tolist()
    msrc = {perm[i]: i for i in range(25)}
    mtar = {perm[25 + i]: i for i in range(15)}
    xm, ym = torch.randn(40, 3, generator=gen), torch.arange(40)
    eim = torch.randint(0, 40, (2, 100), generator=gen)
    masks = {k: torch.rand(40, generator=gen) < 0.5 for k in ("train_mask", "val_mask", "test_mask", "central_mask")}
    d = Data(x=xm.clone(), y=ym.clone(), edge_index=eim.clone(), **{k: v.clone() for k, v in masks.items()})
    d = mb.reorder(d, Data(x=torch.zeros(25, 3)), msrc, mtar)
    xr, yr, mr, er = bo.reorder(xm, ym, masks, eim, 25, msrc, mtar)
    assert torch.equal(d.x, xr) and torch.equal(d.y, yr) and torch.equal(d.edge_index, er)
    assert all(torch.equal(getattr(d, k), v) for k, v in mr.items())
    src = Data(x=torch.randn(25, 3), y=torch.zeros(25, dtype=torch.long), train_mask=torch.ones(25, dtype=torch.bool))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        mb.check_added_edges_within_domain_validity(eim % 25, torch.rand(100), src, torch.rand(25, 2), 0.1, 0.0, verbose=False)


def test_device_f1_matches_sklearn():
    from sklearn.metrics import f1_score
    from bridged_gnn_b200.main_graph_knowledge_transfer import _f1_macro
    g = torch.Generator().manual_seed(0)
    for nc in (2, 5, 31):
        y, p = torch.randint(0, nc, (500,), generator=g), torch.randint(0, nc, (500,), generator=g)
        assert abs(_f1_macro(y, p, nc) - f1_score(y.numpy(), p.numpy(), average="macro")) < 1e-6
    y, p = torch.tensor([0, 0, 3, 3]), torch.tensor([0, 3, 3, 3])      # classes 1, 2 absent from both
    assert abs(_f1_macro(y, p, 5) - f1_score(y.numpy(), p.numpy(), average="macro")) < 1e-6


def test_tf32_planes_reconstruct_the_matrix():
    """Host-side operand split of the 3 x TF32 GEMMs: both planes are exactly representable in tf32 (low 13 mantissa
    bits zero), zero padded to (16, 32) multiples, and hi + lo reproduces w to 2^-22 relative."""
    from bridged_gnn_b200 import ops
    g = torch.Generator().manual_seed(3)
    w = torch.randn(130, 100, generator=g) * torch.exp(3 * torch.randn(130, 1, generator=g))
    hi, lo = ops.tf32_planes(w)
    assert hi.shape == (144, 128) and lo.shape == (144, 128)
    assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0 and int((lo.view(torch.int32) & 0x1FFF).abs().max()) == 0
    assert float(hi[130:].abs().max()) == 0.0 and float(hi[:, 100:].abs().max()) == 0.0
    rec = (hi.double() + lo.double())[:130, :100]
    assert float(((rec - w.double()).abs() / w.double().abs()).max()) <= 2.0 ** -22
    assert float(((hi[:130, :100].double() - w.double()).abs() / w.double().abs()).max()) <= 2.0 ** -11


def test_run_sequential_pairs_batchnorm_with_relu_on_cpu():
    """clf_transformer on CPU tensors goes through the modules themselves (no CUDA library involved) and equals
    nn.Sequential's own forward; NodeLinear keeps nn.Linear's parameters and state_dict keys."""
    from bridged_gnn_b200.models.KTGNN import NodeLinear, run_sequential
    torch.manual_seed(0)
    seq = torch.nn.Sequential(NodeLinear(8, 8), torch.nn.BatchNorm1d(8), torch.nn.ReLU(), NodeLinear(8, 8))
    ref = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.BatchNorm1d(8), torch.nn.ReLU(), torch.nn.Linear(8, 8))
    assert list(seq.state_dict().keys()) == list(ref.state_dict().keys())
    ref.load_state_dict(seq.state_dict())
    x = torch.randn(50, 8)
    assert torch.allclose(run_sequential(seq, x), ref(x), atol=1e-6)
    seq.eval(), ref.eval()
    assert torch.allclose(run_sequential(seq, x), ref(x), atol=1e-6)


def test_graph_save_load_round_trip_and_reference_pickle(tmp_path):
    """step 2 must read what step 1 writes.  save_graph -> load_graph round-trips every tensor; load_graph
    also reads the reference's artefact, a pickled torch_geometric Data (main_bridged_graph.py:317-320), emulated here
    by attribute-bag classes under the PyG module names."""
    import sys
    import types
    from bridged_gnn_b200.data import Data, load_graph, save_graph
    g = torch.Generator().manual_seed(0)
    d = Data(x=torch.randn(9, 4, generator=g), edge_index=torch.randint(0, 9, (2, 20), generator=g), y=torch.arange(9) % 3,
             train_mask=torch.rand(9, generator=g) < 0.5, val_mask=torch.zeros(9, dtype=torch.bool),
             test_mask=torch.ones(9, dtype=torch.bool), central_mask=torch.arange(9) < 6)
    p = str(tmp_path / "toy_bridged_graph.pt")
    save_graph(d, p)
    back = load_graph(p)
    assert sorted(back.keys()) == sorted(d.keys())
    assert all(torch.equal(getattr(back, k), getattr(d, k)) for k in d.keys())
    # the step-2 entry point reads it through the same loader
    from bridged_gnn_b200 import main_graph_knowledge_transfer as step2
    assert step2.load_graph is load_graph
    # a pickled PyG-style object
    mods = {}
    for name, classes in (("torch_geometric", []), ("torch_geometric.data", []), ("torch_geometric.data.data", ["Data"]),
                          ("torch_geometric.data.storage", ["GlobalStorage"])):
        m = types.ModuleType(name)
        for c in classes:
            cls = type(c, (), {})
            cls.__module__ = name
            setattr(m, c, cls)
        mods[name] = m
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        store = mods["torch_geometric.data.storage"].GlobalStorage()
        store.__dict__["_mapping"] = {k: getattr(d, k) for k in d.keys()}
        obj = mods["torch_geometric.data.data"].Data()
        obj.__dict__["_store"] = store
        p2 = str(tmp_path / "ref_bridged_graph.dat")
        torch.save(obj, p2)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    back2 = load_graph(p2)
    assert all(torch.equal(getattr(back2, k), getattr(d, k)) for k in d.keys())


def test_mask_caches_hold_the_mask_not_its_address():
    """AdaptedConv's per-mask caches must not be fooled by an allocator that recycles an address."""
    from bridged_gnn_b200.models import AdaptedConv
    conv = AdaptedConv(4, 2, root_weight=False)
    m1 = torch.tensor([True, True, False, False])
    u1 = conv._dst_is_src(m1)
    assert conv._dst_is_src(m1) is u1                          # same tensor, same version: cached
    m2 = torch.tensor([True, False, False, False])
    u2 = conv._dst_is_src(m2)
    assert u2.tolist() == [1, 0, 0, 0]
    m1[1] = False                                              # in-place edit bumps the version
    assert conv._dst_is_src(m1).tolist() == [1, 0, 0, 0]
    assert conv._domain_counts(m1).tolist() == [1.0, float(torch.tensor(1.0) / 3)]
