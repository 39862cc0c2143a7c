"""Oracle for bridged-graph construction (similarity + per-row top-k -> edge list).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  All citations are
``Bridged-GNN/<file>:<line>`` in the reference checkout.

Weights are passed as a flat ``dict[str, Tensor]`` keyed by the reference's own
``state_dict`` names (e.g. ``source_learner.sim_net.lin_self.1.weight``).
"""
import os

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # torch.nn.BatchNorm1d default, as constructed at models/models.py:93-95, 920-922


# ----------------------------------------------------------------------------- small pieces
def pair_enumeration(x1, x2):
    """models/models.py:265-282.  x1 [A,1], x2 [B,1] -> [A*B, 2]; x1 varies fastest."""
    assert x1.ndimension() == 2 and x2.ndimension() == 2
    x1_ = x1.repeat(x2.size(0), 1)
    x2_ = x2.repeat(1, x1.size(0)).view(-1, x1.size(1))
    return torch.cat((x1_, x2_), dim=1)


def pair_norm(x, mode="None", scale=1.0):
    """models/models.py:49-64 (PairNorm.forward)."""
    if mode == "None":
        return x
    col_mean = x.mean(dim=0)
    if mode == "PN":
        x = x - col_mean
        return scale * x / (1e-6 + x.pow(2).sum(dim=1).mean()).sqrt()
    if mode == "PN-SI":
        x = x - col_mean
        return scale * x / (1e-6 + x.pow(2).sum(dim=1, keepdim=True)).sqrt()
    if mode == "PN-SCS":
        return scale * x / (1e-6 + x.pow(2).sum(dim=1, keepdim=True)).sqrt() - col_mean
    raise ValueError(mode)


def _bn_eval(x, W, prefix):
    return F.batch_norm(x, W[prefix + ".running_mean"], W[prefix + ".running_var"],
                        W[prefix + ".weight"], W[prefix + ".bias"], training=False, eps=BN_EPS)


def _lin(x, W, prefix):
    return F.linear(x, W[prefix + ".weight"], W.get(prefix + ".bias"))


def sage_conv_mean(x, edge_index, W, prefix):
    """PyG SAGEConv(aggr='mean', root_weight=True) [upstream], as used at models/models.py:229-236:
    lin_l(mean_{j->i} x_j) + lin_r(x_i); flow source_to_target (row 0 = j, row 1 = i)."""
    n = x.shape[0]
    agg = torch.zeros_like(x).index_add_(0, edge_index[1], x.index_select(0, edge_index[0]))
    cnt = torch.zeros(n, dtype=x.dtype).index_add_(0, edge_index[1], torch.ones(edge_index.shape[1], dtype=x.dtype))
    agg = agg / cnt.clamp(min=1).unsqueeze(-1)
    return _lin(agg, W, prefix + ".lin_l") + F.linear(x, W[prefix + ".lin_r.weight"])


def graph_encoder(x, edge_index, W, prefix, norm_mode="None", norm_scale=1.0, n_layers=2):
    """models/models.py:245-263 (GraphEncoder.forward), eval mode (dropout is identity)."""
    for i in range(n_layers):
        x = sage_conv_mean(x, edge_index, W, f"{prefix}.convs.{i}")
        if i != n_layers - 1:
            x = F.relu(pair_norm(x, norm_mode, norm_scale))
    return x


def mlp_backbone(x, W, prefix, use_norm=True, norm_mode="None", norm_scale=1.0, n_layers=2):
    """models/models.py:880-893 (MLP.forward), eval mode."""
    for i in range(n_layers):
        x = _lin(x, W, f"{prefix}.layers.{i}")
        if i != n_layers - 1:
            if use_norm:
                x = pair_norm(x, norm_mode, norm_scale)
            x = F.relu(x)
    return x


def embed(x_src, ei_src, x_tar, ei_tar, W, cfg):
    """z_src = source_learner.backbone(x, ei); z_tar = target_learner.encode(data)[0]
    (models/models.py:835-836, 1133-1134; encode at :740-744, 1092-1096)."""
    nm, nsc = cfg.get("norm_mode", "None"), cfg.get("norm_scale", 1.0)
    z_src = z_tar = None
    if x_src is not None:
        if cfg["backbone"] == "gnn":
            z_src = graph_encoder(x_src, ei_src, W, "source_learner.backbone", nm, nsc)
        else:
            z_src = mlp_backbone(x_src, W, "source_learner.backbone", True, nm, nsc)
    if x_tar is not None:
        h0 = torch.tanh(pair_norm(_lin(x_tar, W, "target_learner.equavilent_trans_layer.0"), nm, nsc))
        if cfg["backbone"] == "gnn":
            z_tar = graph_encoder(h0, ei_tar, W, "target_learner.encoder", nm, nsc)
        else:
            z_tar = mlp_backbone(h0, W, "target_learner.encoder", True, nm, nsc)
    return z_src, z_tar


def clf_probs(z, W):
    """exp(log_softmax(lin_clf(relu(z)))) -- models/models.py:137-140 + :843 (eval: dropout identity)."""
    return F.log_softmax(_lin(F.relu(z), W, "source_learner.sim_net.lin_clf"), dim=-1).exp()


# ----------------------------------------------------------------------------- pair similarity (faithful)
def _cos_head(z, W):
    """lin_self then u = z' + biasatt(z') -- models/models.py:91-97 (lin_self), :70-74 (biasatt), :125-127."""
    p = "source_learner.sim_net"
    h = _bn_eval(z, W, p + ".lin_self.0")
    h = F.linear(h, W[p + ".lin_self.1.weight"])
    h = torch.tanh(_bn_eval(h, W, p + ".lin_self.2"))
    return F.linear(h, W[p + ".lin_self.4.weight"])


def _biasatt(u, W):
    p = "source_learner.sim_net.biasatt"
    return _lin(torch.tanh(_lin(u, W, p + ".0")), W, p + ".2")


def sim_pairs(z_db, z_q, idx_db, idx_q, W, sim_mode):
    """sigmoid similarity of enumerated pairs, reference op order (materialises every pair).
    cosine: models/models.py:124-130 / 945-948;  mlp: models/models.py:949-954 (lin_self at :918-925)."""
    if sim_mode == "cosine":
        a = _cos_head(z_db, W)
        b = _cos_head(z_q, W)
        ga, gb = a[idx_db], b[idx_q]
        alpha = torch.nn.CosineSimilarity(dim=1)(ga + _biasatt(ga, W), gb + _biasatt(gb, W))
    elif sim_mode == "mlp":
        p = "source_learner.sim_net.lin_self"
        xp = torch.cat((z_db[idx_db], z_q[idx_q]), dim=1)
        h = _bn_eval(xp, W, p + ".0")
        h = F.relu(_bn_eval(_lin(h, W, p + ".1"), W, p + ".2"))
        alpha = _lin(h, W, p + ".4").squeeze(-1)
    else:
        raise ValueError(sim_mode)
    return torch.sigmoid(alpha)


def coalesce(edge_index, num_nodes=None):
    """torch_geometric.utils.coalesce [upstream]: sort by row*N+col, drop duplicates
    (call sites main_bridged_graph.py:75, 113)."""
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    key = edge_index[0] * n + edge_index[1]
    key, perm = torch.sort(key)
    ei = edge_index[:, perm]
    keep = torch.ones_like(key, dtype=torch.bool)
    keep[1:] = key[1:] != key[:-1]
    return ei[:, keep]


def add_topk_sim_cross_domain_edges(x_src, ei_src, x_tar, ei_tar, W, cfg, k=3, batch_size=1000, canonical=False):
    """main_bridged_graph.py:33-75, faithful: target-row chunks, all pairs materialised, backbones
    recomputed per chunk, topk(sorted=False), edges (src, tar), coalesce.  ``canonical=True`` replaces
    torch.topk by the (value desc, index asc) selection used as the parity key."""
    ns, nt = x_src.shape[0], x_tar.shape[0]
    all_src = torch.arange(ns).unsqueeze(-1)
    start, buck, sims, idxs = 0, [], [], []
    with torch.no_grad():
        while start < nt:
            end = min(start + batch_size, nt)
            b_tar = torch.arange(start, end).unsqueeze(-1)
            pairs = pair_enumeration(all_src, b_tar).transpose(0, 1)
            z_src, z_tar = embed(x_src, ei_src, x_tar, ei_tar, W, cfg)          # recomputed every batch (F4)
            p_src, p_tar = clf_probs(z_src, W), clf_probs(z_tar, W)
            sim_mat = sim_pairs(z_src, z_tar, pairs[0], pairs[1], W, cfg["sim_mode"]).view(-1, ns)
            vals, ind = canonical_topk(sim_mat, k) if canonical else sim_mat.topk(k=k, dim=1, largest=True, sorted=False)
            tar_col = torch.cat([b_tar for _ in range(k)], dim=1).view(-1)
            buck.append(torch.stack((ind.reshape(-1), tar_col), dim=0))
            sims.append(vals)
            idxs.append(ind)
            start = end
    ei = torch.cat(buck, dim=1)
    return coalesce(ei), torch.cat(sims, 0), torch.cat(idxs, 0), p_src, p_tar


def add_topk_sim_within_domain_edges(x, ei, W, cfg, k=3, batch_size=1000, domain="source", canonical=False):
    """main_bridged_graph.py:77-120, faithful (self is NOT excluded; edge = (neighbour, query))."""
    n = x.shape[0]
    all_idx = torch.arange(n).unsqueeze(-1)
    start, buck, sims, idxs = 0, [], [], []
    with torch.no_grad():
        while start < n:
            end = min(start + batch_size, n)
            b = torch.arange(start, end).unsqueeze(-1)
            pairs = pair_enumeration(all_idx, b).transpose(0, 1)
            if domain == "source":
                z, _ = embed(x, ei, None, None, W, cfg)
            else:
                _, z = embed(None, None, x, ei, W, cfg)
            sim_mat = sim_pairs(z, z, pairs[0], pairs[1], W, cfg["sim_mode"]).view(-1, n)
            vals, ind = canonical_topk(sim_mat, k) if canonical else sim_mat.topk(k=k, dim=1, largest=True, sorted=False)
            to_col = torch.cat([b for _ in range(k)], dim=1).view(-1)
            buck.append(torch.stack((ind.reshape(-1), to_col), dim=0))
            sims.append(vals)
            idxs.append(ind)
            start = end
    return coalesce(torch.cat(buck, dim=1)), torch.cat(sims, 0), torch.cat(idxs, 0)


# ----------------------------------------------------------------------------- parity key
def canonical_topk(sim_mat, k):
    """Per-row top-k under the parity key (post-sigmoid fp32 value desc, column index asc).
    torch.topk's tie choice is implementation-defined (main_bridged_graph.py:60 uses sorted=False)."""
    v, i = torch.sort(sim_mat, dim=1, descending=True, stable=True)
    return v[:, :k].contiguous(), i[:, :k].contiguous()


def near_tie_rows(sim_mat, k, tol=1e-6):
    """Rows whose k-th and (k+1)-th best values are closer than ``tol`` (|d sim| < 1e-6 per the spec)."""
    if sim_mat.shape[1] <= k:
        return torch.zeros(sim_mat.shape[0], dtype=torch.bool)
    v, _ = torch.sort(sim_mat, dim=1, descending=True, stable=True)
    return (v[:, k - 1] - v[:, k]).abs() < tol


def full_sim_matrix(z_db, z_q, W, sim_mode, chunk=256):
    """[Nq, Ndb] similarity via the faithful pair path, chunked over query rows."""
    ndb = z_db.shape[0]
    all_db = torch.arange(ndb).unsqueeze(-1)
    out = []
    with torch.no_grad():
        for s in range(0, z_q.shape[0], chunk):
            b = torch.arange(s, min(s + chunk, z_q.shape[0])).unsqueeze(-1)
            pairs = pair_enumeration(all_db, b).transpose(0, 1)
            out.append(sim_pairs(z_db, z_q, pairs[0], pairs[1], W, sim_mode).view(-1, ndb))
    return torch.cat(out, 0)


def cosine_knn_rows(u_db, u_q, k, rows=None, chunk=128):
    """Faithful cosine kNN on node vectors ``u`` fed to the cosine directly (synthetic configs 4/5:
    SURVEY 8d): per pair sigmoid(CosineSimilarity(u_db[i], u_q[j])) via gathered pairs, then canonical
    top-k.  Returns (vals, idx, near_tie_mask)."""
    if rows is None:
        rows = torch.arange(u_q.shape[0])
    ndb = u_db.shape[0]
    all_db = torch.arange(ndb).unsqueeze(-1)
    V, I, T = [], [], []
    with torch.no_grad():
        for s in range(0, rows.numel(), chunk):
            b = rows[s:s + chunk].unsqueeze(-1)
            pairs = pair_enumeration(all_db, b).transpose(0, 1)
            sim = torch.sigmoid(torch.nn.CosineSimilarity(dim=1)(u_db[pairs[0]], u_q[pairs[1]])).view(-1, ndb)
            v, i = canonical_topk(sim, k)
            V.append(v); I.append(i); T.append(near_tie_rows(sim, k))
    return torch.cat(V), torch.cat(I), torch.cat(T)


def set_threads():
    torch.set_num_threads(os.cpu_count() or 1)
    return torch.get_num_threads()


# ----------------------------------------------------------------------------- edge-validity filters / reorder
def check_added_edges_cross_domain_validity(edge_index_added, e_sim, x_src, y_src, x_tar, y_tar, train_mask_tar,
                                            probs_clf_src, probs_clf_tar, thres_conf_quantile=0.1, thres_feat_sim=0.0):
    """main_bridged_graph.py:225-264, restated on plain tensors.  ``e_sim`` is indexed positionally against
    the columns of ``edge_index_added`` exactly as the reference does (:237-239)."""
    pred_s, pred_t = probs_clf_src.argmax(dim=1), probs_clf_tar.argmax(dim=1)
    e0, e1 = edge_index_added[0], edge_index_added[1]
    rm = torch.zeros(edge_index_added.shape[1], dtype=torch.bool)
    e_sim = e_sim.view(-1)
    rm[e_sim < e_sim.quantile(q=thres_conf_quantile)] = True
    rm[pred_s[e0] != y_src[e0]] = True
    rm[(pred_t[e1] != y_tar[e1]) * train_mask_tar[e1]] = True
    rm[pred_s[e0] != pred_t[e1]] = True
    rm[F.cosine_similarity(x_src[e0], x_tar[e1]) < thres_feat_sim] = True
    return edge_index_added[:, ~rm]


def check_added_edges_within_domain_validity(edge_index_added, e_sim, x, y, train_mask, probs_clf,
                                             thres_conf_quantile=0.1, thres_feat_sim=0.0):
    """main_bridged_graph.py:123-161 (note :141-142: both label rules are gated by train_mask of the
    DESTINATION end, as in the reference)."""
    pred = probs_clf.argmax(dim=1)
    e0, e1 = edge_index_added[0], edge_index_added[1]
    rm = torch.zeros(edge_index_added.shape[1], dtype=torch.bool)
    e_sim = e_sim.view(-1)
    rm[e_sim < e_sim.quantile(q=thres_conf_quantile)] = True
    rm[(pred[e0] != y[e0]) * train_mask[e1]] = True
    rm[(pred[e1] != y[e1]) * train_mask[e1]] = True
    rm[pred[e0] != pred[e1]] = True
    rm[F.cosine_similarity(x[e0], x[e1]) < thres_feat_sim] = True
    return edge_index_added[:, ~rm]


def reorder(x, y, masks, edge_index, n_src, mapper_idx_src, mapper_idx_tar):
    """main_bridged_graph.py:195-222: put the merged graph back into the original node order.
    ``mapper_*``: dict original id -> local id.  Returns (x, y, masks, edge_index) re-indexed."""
    merge = dict(mapper_idx_src)
    for k, v in mapper_idx_tar.items():
        assert k not in merge
        merge[k] = v + n_src
    inv = {v: k for k, v in merge.items()}
    items = torch.tensor(sorted(merge.items(), key=lambda t: t[0]), dtype=torch.long)
    order = items[:, 1]
    ei = torch.tensor([[inv[int(i)] for i in edge_index[0]], [inv[int(i)] for i in edge_index[1]]], dtype=torch.long)
    return x[order], y[order], {k: m[order] for k, m in masks.items()}, ei


def merge_graphs(x_src, y_src, ei_src, x_tar, y_tar, ei_tar, train_mask_tar, val_mask_tar, test_mask_tar,
                 edge_index_cross_added, edge_index_added_src=None, edge_index_added_tar=None):
    """main_bridged_graph.py:163-193 on plain tensors: concatenate the two graphs and the added edges (target ids
    offset by N_src), build the masks (:181-189), ``Data(...).coalesce()`` (:191-193).  Returns a dict."""
    n_src, n_tar = x_src.shape[0], x_tar.shape[0]
    n = n_src + n_tar
    cross = edge_index_cross_added.clone()
    cross[1, :] += n_src
    parts = [ei_src, ei_tar + n_src, cross]
    if edge_index_added_src is not None:
        parts.append(edge_index_added_src)
    if edge_index_added_tar is not None:
        parts.append(edge_index_added_tar + n_src)
    central = torch.zeros(n, dtype=torch.bool)
    central[:n_src] = True
    train, val, test = torch.zeros(n, dtype=torch.bool), torch.zeros(n, dtype=torch.bool), torch.zeros(n, dtype=torch.bool)
    train[central] = True
    train[torch.where(y_src == -1)] = False
    train[torch.where(train_mask_tar)[0] + n_src] = True
    val[torch.where(val_mask_tar)[0] + n_src] = True
    test[torch.where(test_mask_tar)[0] + n_src] = True
    return dict(x=torch.cat((x_src, x_tar), 0), edge_index=coalesce(torch.cat(parts, dim=1), n), y=torch.cat((y_src, y_tar), 0),
                train_mask=train, val_mask=val, test_mask=test, central_mask=central)
