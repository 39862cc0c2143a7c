"""Bridged-graph construction with the reference's function surface (Bridged-GNN/main_bridged_graph.py).

``add_topk_sim_cross_domain_edges`` / ``add_topk_sim_within_domain_edges`` keep the reference's names,
argument meaning and return values (main_bridged_graph.py:33-75, 77-120) but never enumerate pairs:
node embeddings and the node-wise operands of the similarity head are computed once, then one fused
kernel scores every (query, db) pair on chip and keeps the top-k per row.  ``merge_graphs`` and
``gen_bridged_graph`` assemble the merged graph on the device.
"""
import argparse
import os

import torch

from . import ops
from .data import Data, coalesce, load_graph, save_graph
from .models.models import Adversarial_Learner, Adversarial_Learner_v2
from .utils import eval_bridged_Graph, eval_homophily

__all__ = ["add_topk_sim_cross_domain_edges", "add_topk_sim_within_domain_edges", "check_added_edges_cross_domain_validity",
           "check_added_edges_within_domain_validity", "merge_graphs", "reorder", "gen_bridged_graph", "split_domains", "main"]


def _homophily(y_from, y_to, edge_index):
    lab = (y_from[edge_index[0]] != -1) & (y_to[edge_index[1]] != -1)
    same = (y_from[edge_index[0]] == y_to[edge_index[1]]) & lab
    return (same.sum() / lab.sum().clamp(min=1)).item()


def _knn(sim_net, z_db, z_q, k, algo, eps=None):
    """Top-k db rows per query row under the head's similarity.  Returns idx [nq,k], val [nq,k], gap [nq].
    ``eps``: threshold applied inside the selection epilogue (neighbours with similarity <= eps get idx -1)."""
    if sim_net.mode == "cosine":
        u_db = sim_net.cosine_operand(z_db)
        u_q = u_db if z_q is z_db else sim_net.cosine_operand(z_q)
        idx, val, gap = ops.knn_cosine(u_q, u_db, k, normalize=True, apply_sigmoid=True, algo=algo, eps=eps)[:3]
    else:
        U_db, U_q, w2, b2 = sim_net.mlp_operands(z_db, z_q)
        idx, val, gap = ops.knn_addrelu(U_q, U_db, w2, b2, k, apply_sigmoid=True, eps=eps)[:3]
    return idx, val, gap


def _edges_from_topk(idx):
    """edge = (neighbour, query) for every kept neighbour (idx >= 0), query-major like the reference's buckets."""
    nq, k = idx.shape
    to = torch.arange(nq, device=idx.device).unsqueeze(1).expand(nq, k)
    edges = torch.stack((idx.reshape(-1), to.reshape(-1)), dim=0)
    return edges[:, edges[0] >= 0]


def add_topk_sim_cross_domain_edges(data_src, data_tar, model, epsilon=0.5, k=3, batch_size=1000, apply_epsilon=False,
                                    algo="auto", return_gap=False, verbose=True):
    """For every target node, the k most similar source nodes (main_bridged_graph.py:33-75).

    Returns ``(edge_index [2,E] int64 CPU coalesced (src, tar); e_sim_mat [Nt,k] fp32 CPU; idx_src_mat
    [Nt,k] int64 CPU; probs_clf_src [Ns,C]; probs_clf_tar [Nt,C])`` like the reference.  Within a row the
    neighbours come best first (the reference's ``topk(sorted=False)`` order is unspecified).
    ``epsilon`` is accepted and, as in the reference (where the argument is never read), ignored unless
    ``apply_epsilon=True``, which drops pairs with similarity <= epsilon inside the kernel's selection epilogue
    (``idx_src_mat`` then holds -1 at the dropped places).  ``batch_size`` is ignored: nothing is materialised per batch.
    """
    with torch.no_grad():
        model.eval()
        z_src = model.embed_source(data_src)
        z_tar = model.embed_target(data_tar)
        probs_src, probs_tar = model.clf_probs(z_src), model.clf_probs(z_tar)
        if probs_src is None:   # source_clf=False: the reference returns zeros.exp() == ones
            probs_src = torch.ones((z_src.shape[0], int(data_src.y.max().item()) + 1), device=z_src.device)
            probs_tar = torch.ones((z_tar.shape[0], int(data_tar.y.max().item()) + 1), device=z_tar.device)
        idx, val, gap = _knn(model.source_learner.sim_net, z_src, z_tar, k, algo, eps=epsilon if apply_epsilon else None)
        edge_index = _edges_from_topk(idx)      # the threshold was applied by the kernel's selection epilogue
        if verbose:
            print("Current homophily ratio:", _homophily(data_src.y, data_tar.y, edge_index))
        edge_index = coalesce(edge_index)
    out = (edge_index.cpu(), val.cpu(), idx.cpu(), probs_src, probs_tar)
    return out + (gap.cpu(),) if return_gap else out


def add_topk_sim_within_domain_edges(data_src, model, k=3, batch_size=1000, domain="source", algo="auto",
                                     return_gap=False, verbose=True):
    """k most similar nodes of the same domain for every node (main_bridged_graph.py:77-120); a node's
    own row is not excluded, exactly as in the reference.  Returns ``(edge_index [2,E] (from, to)
    coalesced, e_sim_mat [N,k], idx_mat [N,k])`` on the CPU."""
    with torch.no_grad():
        model.eval()
        z = model.embed_source(data_src) if domain == "source" else model.embed_target(data_src)
        idx, val, gap = _knn(model.source_learner.sim_net, z, z, k, algo)
        edge_index = coalesce(_edges_from_topk(idx))
        if verbose:
            print("Current homophily ratio of Graph:", _homophily(data_src.y, data_src.y, edge_index))
    out = (edge_index.cpu(), val.cpu(), idx.cpu())
    return out + (gap.cpu(),) if return_gap else out


def _report(counts, n_out, y_from, y_to, out, verbose):
    if verbose:
        c = counts.tolist()
        print("1. remove low SimNet Confidence edges:", c[0])
        print("2. remove edges that include node with wrong pred label compared with training label (ground truth):", c[1])
        print("3. remove edges that the clf predicted label of two nodes (end points) are different:", c[2])
        print("4. remove low raw_feat_sim edges:", c[3])
        print("[Done] Totally remove edges: {} | Current total edge num: {}".format(sum(c[:4]), n_out))
        print("Current homophily ratio:", _homophily(y_from, y_to, out))


def check_added_edges_cross_domain_validity(edge_index_added, e_sim, data_src, data_tar, probs_clf_src, probs_clf_tar,
                                            thres_conf_quantile=0.1, thres_feat_sim=0.0, verbose=True):
    """The 4-rule filter of main_bridged_graph.py:225-264 on the device: the similarity quantile by radix select
    (``ops.quantile``: no sort, no 16 M-element cap) and the four rules -- the per-edge raw-feature cosine included --
    in one kernel (``ops.edge_validity``).  ``e_sim`` is matched positionally with the columns of
    ``edge_index_added``, as in the reference; pass similarities in the edge order you want compared (the reference
    passes the [Nt, k] matrix flattened against the re-sorted coalesced list)."""
    dev = data_src.x.device
    ei = edge_index_added.to(dev)
    e_sim = e_sim.to(dev).reshape(-1).float()
    pred_s, pred_t = probs_clf_src.to(dev).argmax(dim=1), probs_clf_tar.to(dev).argmax(dim=1)
    keep, counts = ops.edge_validity(ei, e_sim, ops.quantile(e_sim, thres_conf_quantile), pred_s, data_src.y, pred_t, data_tar.y,
                                     None, data_tar.train_mask, data_src.x, data_tar.x, thres_feat_sim)
    out = ei[:, keep]
    _report(counts, out.shape[1], data_src.y, data_tar.y, out, verbose)
    return out.to(edge_index_added.device)


def check_added_edges_within_domain_validity(edge_index_added, e_sim, data_in, probs_clf, thres_conf_quantile=0.1,
                                             thres_feat_sim=0.0, verbose=True):
    """main_bridged_graph.py:123-161 (both label rules gated by the train mask of the destination end), same kernels."""
    dev = data_in.x.device
    ei = edge_index_added.to(dev)
    e_sim = e_sim.to(dev).reshape(-1).float()
    pred = probs_clf.to(dev).argmax(dim=1)
    keep, counts = ops.edge_validity(ei, e_sim, ops.quantile(e_sim, thres_conf_quantile), pred, data_in.y, pred, data_in.y,
                                     data_in.train_mask, data_in.train_mask, data_in.x, data_in.x, thres_feat_sim)
    out = ei[:, keep]
    _report(counts, out.shape[1], data_in.y, data_in.y, out, verbose)
    return out.to(edge_index_added.device)


def reorder(data_merge, data_src, mapper_idx_src, mapper_idx_tar):
    """main_bridged_graph.py:195-222 without the per-edge Python dict lookups: the two id maps become one
    permutation tensor, node arrays are gathered with it and edge ids are mapped by one index_select."""
    n_src = data_src.x.shape[0]
    dev = data_merge.x.device
    keys = list(mapper_idx_src.keys()) + list(mapper_idx_tar.keys())
    vals = list(mapper_idx_src.values()) + [v + n_src for v in mapper_idx_tar.values()]
    if len(set(keys)) != len(keys):
        raise AssertionError("source and target id maps overlap")
    k = torch.tensor(keys, dtype=torch.long, device=dev)
    v = torch.tensor(vals, dtype=torch.long, device=dev)
    order = v[torch.argsort(k)]                      # merged id of the node with the i-th smallest original id
    inverse = torch.empty(int(v.max()) + 1, dtype=torch.long, device=dev)
    inverse[v] = k                                   # merged id -> original id
    for name in ("train_mask", "val_mask", "test_mask", "central_mask", "x", "y"):
        setattr(data_merge, name, getattr(data_merge, name)[order])
    data_merge.edge_index = inverse[data_merge.edge_index]
    return data_merge


def merge_graphs(data_src, data_tar, edge_index_cross_added, edge_index_added_src=None, edge_index_added_tar=None):
    """main_bridged_graph.py:163-193 on the device: concatenate the two graphs and the added edges with
    target ids offset by Ns, build the masks, coalesce.  Inputs are not modified."""
    dev = data_src.x.device
    n_src, n_tar = data_src.x.shape[0], data_tar.x.shape[0]
    n = n_src + n_tar
    off = torch.tensor([[0], [n_src]], device=dev)
    parts = [data_src.edge_index.to(dev), data_tar.edge_index.to(dev) + n_src, edge_index_cross_added.to(dev) + off]
    if edge_index_added_src is not None:
        parts.append(edge_index_added_src.to(dev))
    if edge_index_added_tar is not None:
        parts.append(edge_index_added_tar.to(dev) + n_src)
    edge_index = coalesce(torch.cat(parts, dim=1), n)
    central = torch.zeros(n, dtype=torch.bool, device=dev)
    central[:n_src] = True
    train = central.clone()
    train[:n_src] &= data_src.y.to(dev) != -1
    train[n_src:] = data_tar.train_mask.to(dev)
    val, test = torch.zeros_like(central), torch.zeros_like(central)
    val[n_src:] = data_tar.val_mask.to(dev)
    test[n_src:] = data_tar.test_mask.to(dev)
    return Data(x=torch.cat((data_src.x, data_tar.x), 0), edge_index=edge_index, y=torch.cat((data_src.y, data_tar.y), 0),
                train_mask=train, val_mask=val, test_mask=test, central_mask=central)


def gen_bridged_graph(args, data_src, data_tar, device, path_ckpt, mapper_idx_src=None, mapper_idx_tar=None,
                      epsilon=0.5, batch_size=1000):
    """main_bridged_graph.py:267-321: load the similarity learner, add cross- and within-domain edges,
    optionally filter them, merge and re-order, print the homophily diagnostics (:315-316, without the dense
    N x N matrix of the reference: bridged_gnn_b200.utils)."""
    if args.version == "v1":
        sim_model = Adversarial_Learner(data_src, data_tar, dim_hidden=args.hidden_dim, num_layer=args.num_layer,
                                        source_clf=True, norm_mode=args.norm_mode, norm_scale=args.norm_scale)
    else:
        sim_model = Adversarial_Learner_v2(data_src, data_tar, dim_hidden=args.hidden_dim, num_layer=args.num_layer,
                                           use_norm=True, source_clf=True, norm_mode=args.norm_mode,
                                           norm_scale=args.norm_scale, sim_mode=args.sim_mode, backbone=args.backbone)
    sim_model.load_state_dict(torch.load(path_ckpt, map_location="cpu"))
    data_src, data_tar, sim_model = data_src.to(device), data_tar.to(device), sim_model.to(device)
    ei_cross, e_sim, idx_src, p_src, p_tar = add_topk_sim_cross_domain_edges(
        data_src, data_tar, sim_model, epsilon=epsilon, k=args.k_cross, batch_size=batch_size)
    if getattr(args, "check_cross", False):
        ei_cross = check_added_edges_cross_domain_validity(ei_cross, e_sim.view(-1), data_src, data_tar, p_src, p_tar,
                                                           thres_conf_quantile=args.thres_conf_quantile,
                                                           thres_feat_sim=args.thres_feat_sim)
    ei_src = ei_tar = None
    if args.k_within > 0:
        ei_src, s_src, _ = add_topk_sim_within_domain_edges(data_src, sim_model, k=args.k_within, batch_size=100, domain="source")
        ei_tar, s_tar, _ = add_topk_sim_within_domain_edges(data_tar, sim_model, k=args.k_within, batch_size=100, domain="target")
        if getattr(args, "check_within", False):      # constants as hard-coded at main_bridged_graph.py:301-306
            ei_src = check_added_edges_within_domain_validity(ei_src, s_src.view(-1), data_src, p_src, 0.1, 0.8)
            ei_tar = check_added_edges_within_domain_validity(ei_tar, s_tar.view(-1), data_tar, p_tar, 0.1, 0.8)
    data_merge = merge_graphs(data_src, data_tar, ei_cross, ei_src, ei_tar)
    if mapper_idx_src is not None and mapper_idx_tar is not None:
        data_merge = reorder(data_merge, data_src, mapper_idx_src, mapper_idx_tar)
    if getattr(args, "diagnostics", True):
        eval_homophily(data_merge)
        eval_bridged_Graph(data_merge)
    if getattr(args, "save", False):
        out_dir = getattr(args, "out_dir", None) or "../data_bridged_graph"
        os.makedirs(out_dir, exist_ok=True)
        # (the reference pickles a PyG Data here, :317-320; data.load_graph reads both that and this dict form)
        save_graph(data_merge, os.path.join(out_dir, f"{args.dataset_name}_bridged_graph.pt"))
    return data_merge


def split_domains(data):
    """(data_src, data_tar) of a merged graph by ``central_mask``, with local ids and the intra-domain edges only --
    what utils.dataset_conversion (utils.py:41-99) hands to gen_bridged_graph -- plus the id maps
    {original id: local id} that ``reorder`` takes."""
    c = data.central_mask
    ids_s, ids_t = torch.nonzero(c).view(-1), torch.nonzero(~c).view(-1)
    local = torch.empty(c.shape[0], dtype=torch.long, device=c.device)
    local[ids_s] = torch.arange(ids_s.numel(), device=c.device)
    local[ids_t] = torch.arange(ids_t.numel(), device=c.device)
    ei = data.edge_index
    ms, mt = c[ei[0]] & c[ei[1]], (~c[ei[0]]) & (~c[ei[1]])
    parts = []
    for ids, m in ((ids_s, ms), (ids_t, mt)):
        kw = {k: getattr(data, k)[ids] for k in ("x", "y", "train_mask", "val_mask", "test_mask") if hasattr(data, k)}
        parts.append(Data(edge_index=local[ei[:, m]], **kw))
    maps = [{int(o): i for i, o in enumerate(ids.tolist())} for ids in (ids_s, ids_t)]
    return parts[0], parts[1], maps[0], maps[1]


def main(args):
    """Step 1 from a graph file and a trained similarity learner (main_bridged_graph.py:325-357 without the
    adversarial training of the learner, which is outside the accelerated path): ``--path_data`` holds the two
    domains in one graph (``central_mask`` marks the source nodes; only intra-domain edges are used),
    ``--path_ckpt`` the learner's state_dict."""
    if not args.path_data or not args.path_ckpt:
        raise SystemExit("--path_data and --path_ckpt are required")
    if not torch.cuda.is_available():
        raise RuntimeError("bridged_gnn_b200 needs a CUDA device (sm_100a); there is no CPU path")
    device = torch.device("cuda", args.gpu)
    data_src, data_tar, m_src, m_tar = split_domains(load_graph(args.path_data))
    return gen_bridged_graph(args, data_src, data_tar, device, args.path_ckpt, m_src, m_tar, epsilon=args.epsilon,
                             batch_size=args.batch_size)


def parse_args(argv=None):
    """The reference's CLI flags that concern graph generation (main_bridged_graph.py:360-393)."""
    p = argparse.ArgumentParser()
    p.add_argument("--dataset_name", type=str, default="office_amazon2dslr")
    p.add_argument("--version", type=str, default="v2", choices=["v1", "v2"])
    p.add_argument("--backbone", type=str, default="mlp", choices=["mlp", "gnn"])
    p.add_argument("--sim_mode", type=str, default="mlp", choices=["mlp", "cosine"])
    p.add_argument("--hidden_dim", type=int, default=128)
    p.add_argument("--num_layer", type=int, default=2)
    p.add_argument("--norm_mode", type=str, default="None")
    p.add_argument("--norm_scale", type=float, default=1.0)
    p.add_argument("--k_cross", type=int, default=20)
    p.add_argument("--k_within", type=int, default=3)
    p.add_argument("--epsilon", type=float, default=0.5)
    p.add_argument("--batch_size", type=int, default=1000)
    p.add_argument("--check_cross", action="store_true")
    p.add_argument("--check_within", action="store_true")
    p.add_argument("--thres_conf_quantile", type=float, default=0.1)
    p.add_argument("--thres_feat_sim", type=float, default=0.0)
    p.add_argument("--save", action="store_true")
    p.add_argument("--gpu", type=int, default=0)
    p.add_argument("--path_data", type=str, default=None, help="bridged-graph .dat to take features / split from")
    p.add_argument("--path_ckpt", type=str, default=None)
    p.add_argument("--out_dir", type=str, default=None, help="where --save writes (default ../data_bridged_graph)")
    return p.parse_args(argv)


if __name__ == "__main__":
    main(parse_args())
