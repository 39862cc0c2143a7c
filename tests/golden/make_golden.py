"""Generate golden vectors by running the UNMODIFIED reference Python.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It installs the PyG stand-ins from ``pyg_shim.py`` (PyG itself is not
installable here), imports the reference's own ``main_bridged_graph.py``,
``models/models.py``, ``models/KTGNN.py`` and ``models/backbones.py`` from
/root/reference/Bridged-GNN, drives them on the shipped office fixture
(``data_bridged_graph/office_amazon2dslr_bridged_graph.dat`` +
``ckpt/model_AdvLearner_office_amazon2dslr_best.ckpt``) and on seeded synthetic
inputs with the shipped fb_hamilton2caltech checkpoint, and writes small
``.npz`` fixtures next to this file.  The GPU box has no /root/reference; tests
read only the committed ``.npz`` files.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, HERE)
import pyg_shim  # noqa: E402

pyg_shim.install()
sys.path.insert(0, os.path.join(REF, "Bridged-GNN"))
sys.path.insert(0, os.path.join(REF, "Bridged-GNN", "models"))

# datasets.py loads Facebook100 from disk at import time (datasets.py:134-139); stub it.
_ds = types.ModuleType("datasets")
_ds.prepare_datasets = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("stub"))
sys.modules["datasets"] = _ds

import main_bridged_graph as ref_build  # noqa: E402  (reference, unmodified)
import models as ref_models  # noqa: E402  (reference models/models.py)
import KTGNN as ref_ktgnn  # noqa: E402
import backbones as ref_backbones  # noqa: E402
from pyg_shim import Data, to_undirected  # noqa: E402

torch.set_num_threads(os.cpu_count())


def load_dat(path):
    """Unpickle a PyG Data .dat with attribute-bag stand-ins; returns dict of tensors."""
    class _Bag:
        def __setstate__(self, s):
            self.__dict__.update(s)
    saved = {k: sys.modules.get(k) for k in ("torch_geometric.data.data", "torch_geometric.data.storage")}
    dd = types.ModuleType("torch_geometric.data.data")
    st = types.ModuleType("torch_geometric.data.storage")
    for n in ("Data", "DataEdgeAttr", "DataTensorAttr"):
        setattr(dd, n, type(n, (_Bag,), {}))
    for n in ("GlobalStorage", "BaseStorage", "NodeStorage", "EdgeStorage"):
        setattr(st, n, type(n, (_Bag,), {}))
    sys.modules["torch_geometric.data.data"] = dd
    sys.modules["torch_geometric.data.storage"] = st
    try:
        obj = torch.load(path, map_location="cpu", weights_only=False)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return dict(obj.__dict__["_store"].__dict__["_mapping"])


def split_domains(d):
    """What utils.dataset_conversion produces (utils.py:41-99) for a graph whose source nodes are a prefix."""
    c = d["central_mask"]
    ns = int(c.sum())
    assert bool(c[:ns].all()) and not bool(c[ns:].any())
    ei = d["edge_index"]
    m_s = c[ei[0]] & c[ei[1]]
    m_t = (~c[ei[0]]) & (~c[ei[1]])
    src = Data(x=d["x"][:ns].clone(), edge_index=ei[:, m_s].clone(), y=d["y"][:ns].clone(),
               train_mask=d["train_mask"][:ns].clone(), val_mask=d["val_mask"][:ns].clone(),
               test_mask=d["test_mask"][:ns].clone())
    tar = Data(x=d["x"][ns:].clone(), edge_index=(ei[:, m_t] - ns).clone(), y=d["y"][ns:].clone(),
               train_mask=d["train_mask"][ns:].clone(), val_mask=d["val_mask"][ns:].clone(),
               test_mask=d["test_mask"][ns:].clone())
    return src, tar, ns


def np_(t):
    return t.detach().cpu().numpy().copy()


def office_build():
    d = load_dat(f"{REF}/data_bridged_graph/office_amazon2dslr_bridged_graph.dat")
    src, tar, ns = split_domains(d)
    model = ref_models.Adversarial_Learner_v2(src, tar, dim_hidden=128, num_layer=2, use_norm=True, source_clf=True,
                                              norm_mode="None", norm_scale=1.0, sim_mode="mlp", backbone="mlp")
    sd = torch.load(f"{REF}/ckpt/model_AdvLearner_office_amazon2dslr_best.ckpt", map_location="cpu", weights_only=False)
    model.load_state_dict(sd)
    model.eval()
    with torch.no_grad():
        z_src = model.source_learner.backbone(src.x, src.edge_index)
        z_tar, _ = model.target_learner.encode(tar)
    ei_c, sim_c, idx_c, p_src, p_tar = ref_build.add_topk_sim_cross_domain_edges(src, tar, model, epsilon=0.5, k=20, batch_size=1000)
    ei_s, sim_s, idx_s = ref_build.add_topk_sim_within_domain_edges(src, model, k=3, batch_size=100, domain="source")
    ei_t, sim_t, idx_t = ref_build.add_topk_sim_within_domain_edges(tar, model, k=3, batch_size=100, domain="target")
    # a few full similarity rows (reference pair path) for value-level checks
    rows = torch.tensor([0, 1, 7, 100, 333, 590])
    all_src = torch.arange(ns).unsqueeze(-1)
    pairs = ref_models.pair_enumeration(all_src, rows.unsqueeze(-1)).t()
    with torch.no_grad():
        probs, *_ = model.get_probs_cross_domain(src, tar, pairs[0], pairs[1], return_representation=True)
    sim_rows = probs.squeeze(-1).view(-1, ns)
    golden_st = d["edge_index"][:, d["central_mask"][d["edge_index"][0]] & ~d["central_mask"][d["edge_index"][1]]]
    keep = {k: np_(v) for k, v in sd.items() if k.startswith(("source_learner.", "target_learner.equavilent", "target_learner.encoder"))}
    np.savez_compressed(
        os.path.join(HERE, "office_a2d_build.npz"),
        x=np_(d["x"]), y=np_(d["y"]), central_mask=np_(d["central_mask"]), train_mask=np_(d["train_mask"]),
        val_mask=np_(d["val_mask"]), test_mask=np_(d["test_mask"]),
        edge_index=np_(d["edge_index"]), shipped_cross_edges=np_(golden_st),
        z_src=np_(z_src), z_tar=np_(z_tar),
        cross_edge_index=np_(ei_c), cross_sim=np_(sim_c), cross_idx=np_(idx_c),
        probs_clf_src=np_(p_src), probs_clf_tar=np_(p_tar),
        within_src_edge_index=np_(ei_s), within_src_sim=np_(sim_s), within_src_idx=np_(idx_s),
        within_tar_edge_index=np_(ei_t), within_tar_sim=np_(sim_t), within_tar_idx=np_(idx_t),
        sim_rows_idx=np_(rows), sim_rows=np_(sim_rows),
        **{"ckpt." + k: v for k, v in keep.items()},
    )
    print("office build:", ei_c.shape, ei_s.shape, ei_t.shape)
    return d


def fb_cosine_build():
    """v1 cosine head with the shipped fb_hamilton2caltech weights; data is not shipped, so the graph and
    features are seeded synthetic stand-ins at reduced node counts (Ns=1200, Nt=400, D_in=1685)."""
    g = torch.Generator().manual_seed(0)
    ns, nt, din = 1200, 400, 1685
    def feats(n):
        x = torch.zeros(n, din)
        cols = torch.randint(0, din, (n, 6), generator=g)
        x.scatter_(1, cols, 1.0)
        return x
    def er(n, m):
        e = torch.randint(0, n, (2, m), generator=g)
        return to_undirected(e, n)
    src = Data(x=feats(ns), edge_index=er(ns, 20000), y=torch.randint(0, 2, (ns,), generator=g),
               train_mask=torch.ones(ns, dtype=torch.bool))
    tar = Data(x=feats(nt), edge_index=er(nt, 4000), y=torch.randint(0, 2, (nt,), generator=g),
               train_mask=torch.ones(nt, dtype=torch.bool))
    model = ref_models.Adversarial_Learner(src, tar, dim_hidden=64, num_layer=2, source_clf=True, norm_mode="None", norm_scale=1.0)
    sd = torch.load(f"{REF}/ckpt/model_AdvLearner_fb_hamilton2caltech_best.ckpt", map_location="cpu", weights_only=False)
    model.load_state_dict(sd)
    model.eval()
    with torch.no_grad():
        z_src = model.source_learner.backbone(src.x, src.edge_index)
        z_tar, _ = model.target_learner.encode(tar)
    ei_c, sim_c, idx_c, p_src, p_tar = ref_build.add_topk_sim_cross_domain_edges(src, tar, model, epsilon=0.5, k=50, batch_size=1000)
    ei_s, sim_s, idx_s = ref_build.add_topk_sim_within_domain_edges(tar, model, k=5, batch_size=100, domain="target")
    rows = torch.tensor([0, 3, 399])
    pairs = ref_models.pair_enumeration(torch.arange(ns).unsqueeze(-1), rows.unsqueeze(-1)).t()
    with torch.no_grad():
        probs, *_ = model.get_probs_cross_domain(src, tar, pairs[0], pairs[1], return_representation=True)
    keep = {k: np_(v) for k, v in sd.items() if k.startswith("source_learner.sim_net.")}
    np.savez_compressed(
        os.path.join(HERE, "fb_h2c_cosine_build.npz"),
        z_src=np_(z_src), z_tar=np_(z_tar),
        cross_edge_index=np_(ei_c), cross_sim=np_(sim_c), cross_idx=np_(idx_c),
        probs_clf_src=np_(p_src), probs_clf_tar=np_(p_tar),
        within_tar_edge_index=np_(ei_s), within_tar_sim=np_(sim_s), within_tar_idx=np_(idx_s),
        sim_rows_idx=np_(rows), sim_rows=np_(probs.squeeze(-1).view(-1, ns)),
        **{"ckpt." + k: v for k, v in keep.items()},
    )
    print("fb cosine build:", ei_c.shape, ei_s.shape)


def office_mp(d):
    """KT-GNN / AdaptedConv / SAGE / GCN on the shipped office bridged graph (to_undirected applied)."""
    n = d["x"].shape[0]
    ei = to_undirected(d["edge_index"], n)
    data = Data(x=d["x"].clone(), edge_index=ei, y=d["y"].clone(), central_mask=d["central_mask"].clone(),
                train_mask=d["train_mask"].clone())
    out = {"edge_index_undirected": np_(ei)}

    # --- full model, eval mode (main_graph_knowledge_transfer.py:179) ---
    ref_build.set_random_seed(0)
    model = ref_ktgnn.KTGNN_no_complement(256, 31, 2, 64, root_weight=False, use_bn=True, dim_share=256, need_complement=False)
    # give BN non-trivial running stats so eval-mode BN is exercised
    with torch.no_grad():
        for bn in list(model.bns) + [model.clf_transformer[1]]:
            bn.running_mean.uniform_(-0.2, 0.2)
            bn.running_var.uniform_(0.5, 1.5)
    model.eval()
    with torch.no_grad():
        lb, lt, ltt, _ = model(data)
    for k, v in model.state_dict().items():
        out["ktgnn.sd." + k] = np_(v)
    out["ktgnn.eval.logp_base"], out["ktgnn.eval.logp_target"], out["ktgnn.eval.logp_trans"] = np_(lb), np_(lt), np_(ltt)
    out["ktgnn.ei1"], out["ktgnn.ei2"] = np_(model.edge_index1), np_(model.edge_index2)

    # --- full model, train mode, dropout off: loss and parameter grads ---
    model.train()
    model.dropout = 0.0
    model.zero_grad()
    lb, lt, ltt, _ = model(data)
    m = data.train_mask
    loss = torch.nn.functional.nll_loss(lb[m], data.y[m]) + torch.nn.functional.nll_loss(lt[m], data.y[m]) \
        + torch.nn.functional.nll_loss(ltt[m], data.y[m])
    loss.backward()
    out["ktgnn.train.loss"] = np_(loss)
    for k, p in model.named_parameters():
        out["ktgnn.train.grad." + k] = np_(p.grad)

    # --- single AdaptedConv 64->31 with random input: output + grads wrt x and params ---
    ref_build.set_random_seed(1)
    conv = ref_ktgnn.AdaptedConv(64, 31, root_weight=False)
    x = torch.randn(n, 64, requires_grad=True)
    y = conv(x, model.edge_index, model.edge_index1, model.edge_index2, data.central_mask)
    gout = torch.randn(n, 31)
    (y * gout).sum().backward()
    for k, v in conv.state_dict().items():
        out["conv.sd." + k] = np_(v)
    out["conv.x"], out["conv.y"], out["conv.gout"], out["conv.gx"] = np_(x), np_(y), np_(gout), np_(x.grad)
    for k, p in conv.named_parameters():
        out["conv.grad." + k] = np_(p.grad)

    # --- GraphSAGE (--no_dtc path, main_graph_knowledge_transfer.py:414-417) and GCN ---
    ds = types.SimpleNamespace(num_features=256, num_classes=31)
    ref_build.set_random_seed(2)
    sage = ref_backbones.GraphSAGE(ds, layer_num=2, hidden=64)
    sage.eval()
    with torch.no_grad():
        out["sage.logp"] = np_(sage(data))
    for k, v in sage.state_dict().items():
        out["sage.sd." + k] = np_(v)
    ref_build.set_random_seed(3)
    gcn = ref_backbones.GCNNet(ds, layer_num=2, hidden=64)
    gcn.eval()
    with torch.no_grad():
        out["gcn.logp"] = np_(gcn(data))
    for k, v in gcn.state_dict().items():
        out["gcn.sd." + k] = np_(v)
    # v1 GraphEncoder (models.py:220-263): SAGEConv over a Tensor edge_index
    ref_build.set_random_seed(4)
    enc = ref_models.GraphEncoder(256, 64, dim_hidden=64, layer_num=2, norm_mode="None")
    enc.eval()
    with torch.no_grad():
        out["enc.z"] = np_(enc(data.x, data.edge_index))
    for k, v in enc.state_dict().items():
        out["enc.sd." + k] = np_(v)
    np.savez_compressed(os.path.join(HERE, "office_a2d_mp.npz"), **out)
    print("office mp: E_undirected", ei.shape[1], "E1", model.edge_index1.shape[1], "E2", model.edge_index2.shape[1])


def diagnostics(d):
    """utils.py:101-131 (eval_bridged_Graph, eval_homophily) on the shipped office bridged graph and on a seeded
    random graph with unlabelled nodes and duplicate edges.  eval_homophily only prints: its output is parsed."""
    import contextlib
    import io
    import utils as ref_utils  # noqa: E402  (reference, unmodified)
    out = {}

    def run(tag, data):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ratio = ref_utils.eval_bridged_Graph(data)
            ref_utils.eval_homophily(data)
        vals = [float(line.split(":")[1]) for line in buf.getvalue().splitlines() if line.startswith("homophily ratio")]
        out[tag + ".local_ratio"] = np.float64(float(ratio))
        out[tag + ".h1"], out[tag + ".h2"] = np.float64(vals[0]), np.float64(vals[1])
        print(tag, float(ratio), vals)

    run("office", Data(x=d["x"].clone(), edge_index=d["edge_index"].clone(), y=d["y"].clone(), test_mask=d["test_mask"].clone()))
    g = torch.Generator().manual_seed(123)
    n, e = 400, 3000
    ei = torch.randint(0, n, (2, e), generator=g)
    ei = torch.cat((ei, ei[:, :40]), 1)
    y = torch.randint(0, 4, (n,), generator=g)
    y[torch.rand(n, generator=g) < 0.25] = -1
    tm = torch.rand(n, generator=g) < 0.5
    out["rand.edge_index"], out["rand.y"], out["rand.test_mask"] = np_(ei), np_(y), np_(tm)
    run("rand", Data(x=torch.zeros(n, 2), edge_index=ei, y=y, test_mask=tm))
    np.savez_compressed(os.path.join(HERE, "diagnostics.npz"), **out)


def office_assemble(d):
    """The reference's own check_added_edges_{cross,within}_domain_validity (main_bridged_graph.py:225-264,
    123-161), merge_graphs (:163-193) and reorder (:195-222), run on the office fixture with the edge lists /
    similarities / classifier outputs the reference's build produced (stored in office_a2d_build.npz)."""
    import contextlib
    import copy
    import io
    g = dict(np.load(os.path.join(HERE, "office_a2d_build.npz")))
    T = torch.from_numpy
    src, tar, ns = split_domains(d)
    ei_c, sim_c = T(g["cross_edge_index"]), T(g["cross_sim"])
    ei_s, sim_s = T(g["within_src_edge_index"]), T(g["within_src_sim"])
    ei_t, sim_t = T(g["within_tar_edge_index"]), T(g["within_tar_sim"])
    p_src, p_tar = T(g["probs_clf_src"]), T(g["probs_clf_tar"])
    out = {}
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        # cross filter at the CLI defaults (q = 0.1, thres_feat_sim = 0) and with rule 4 active / another quantile
        for tag, q, thr in (("cross_q10_f0", 0.1, 0.0), ("cross_q25_f30", 0.25, 0.3), ("cross_q0_f0", 0.0, 0.0)):
            out[tag] = np_(ref_build.check_added_edges_cross_domain_validity(ei_c.clone(), sim_c.view(-1), src, tar, p_src, p_tar,
                                                                            thres_conf_quantile=q, thres_feat_sim=thr))
        # within filters with the constants hard-coded at :301-306
        out["within_src_q10_f80"] = np_(ref_build.check_added_edges_within_domain_validity(ei_s.clone(), sim_s.view(-1), src, p_src, 0.1, 0.8))
        out["within_tar_q10_f80"] = np_(ref_build.check_added_edges_within_domain_validity(ei_t.clone(), sim_t.view(-1), tar, p_tar, 0.1, 0.8))
        out["within_tar_q50_f0"] = np_(ref_build.check_added_edges_within_domain_validity(ei_t.clone(), sim_t.view(-1), tar, p_tar, 0.5, 0.0))
        # merge (:163-193) with the filtered cross edges and both within-domain lists, exactly like gen_bridged_graph :312-313
        src_u = copy.deepcopy(src)
        src_u.y[torch.arange(0, ns, 97)] = -1          # a few unlabelled source nodes exercise train_mask[y == -1] = False
        merged = ref_build.merge_graphs(src_u, tar, copy.deepcopy(T(out["cross_q10_f0"])), copy.deepcopy(ei_s), copy.deepcopy(ei_t))
        out["merge.unlabelled_src"] = np_(torch.arange(0, ns, 97))
        for k in ("edge_index", "y", "train_mask", "val_mask", "test_mask", "central_mask"):
            out["merge." + k] = np_(getattr(merged, k))
        out["merge.x_checksum"] = np_(merged.x.double().sum(1))
        merged_x = merged.x.clone()
        # merge without within-domain edges (k_within = 0 recipes)
        merged0 = ref_build.merge_graphs(src, tar, copy.deepcopy(ei_c))
        out["merge0.edge_index"] = np_(merged0.edge_index)
        # reorder (:195-222): original ids = a seeded permutation of 0..N-1 split over the two domains
        gen = torch.Generator().manual_seed(7)
        n = merged.x.shape[0]
        orig = torch.randperm(n, generator=gen)
        m_src = {int(orig[i]): i for i in range(ns)}
        m_tar = {int(orig[ns + i]): i for i in range(n - ns)}
        out["reorder.orig_ids"] = np_(orig)
        re = ref_build.reorder(merged, src_u, m_src, m_tar)
        for k in ("edge_index", "y", "train_mask", "val_mask", "test_mask", "central_mask"):
            out["reorder." + k] = np_(getattr(re, k))
        out["reorder.x_checksum"] = np_(re.x.double().sum(1))
        assert torch.equal(re.x, merged_x[torch.argsort(orig)])
    np.savez_compressed(os.path.join(HERE, "office_a2d_assemble.npz"), **out)
    print("office assemble:", {k: v.shape for k, v in out.items() if v.ndim == 2})


if __name__ == "__main__":
    d = office_build()
    office_assemble(d)
    diagnostics(d)
    fb_cosine_build()
    office_mp(d)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
