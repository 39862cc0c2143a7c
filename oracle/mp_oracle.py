"""Oracle for message passing over the bridged graph (KT-GNN AdaptedConv, SAGE, GCN).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Citations are
``Bridged-GNN/<file>:<line>``.  Everything is differentiable torch (CPU), so
gradients for backward parity come from torch autograd over the reference's op
order (gather -> per-edge ops -> scatter-add).
"""
import torch
import torch.nn.functional as F


def _scatter_add(src, index, n):
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add(0, index, src)


def pyg_softmax(src, index, num_nodes):
    """torch_geometric.utils.softmax [upstream] (call: models/KTGNN.py:299):
    exp(s - max_dst) / (sum_dst exp(s - max_dst) + 1e-16), max taken without gradient."""
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    mx = torch.full((num_nodes,) + tuple(src.shape[1:]), float("-inf"), dtype=src.dtype)
    mx = mx.scatter_reduce(0, idx, src.detach(), reduce="amax", include_self=True)
    out = (src - mx.index_select(0, index)).exp()
    s = _scatter_add(out, index, num_nodes) + 1e-16
    return out / s.index_select(0, index)


def to_undirected(edge_index, num_nodes):
    """torch_geometric.transforms.ToUndirected(merge=True) [upstream]
    (main_graph_knowledge_transfer.py:410-411): add reversed edges, coalesce."""
    row, col = edge_index
    ei = torch.stack([torch.cat([row, col]), torch.cat([col, row])], 0)
    key = ei[0] * num_nodes + ei[1]
    key, perm = torch.sort(key)
    ei = ei[:, perm]
    keep = torch.ones_like(key, dtype=torch.bool)
    keep[1:] = key[1:] != key[:-1]
    return ei[:, keep]


def graph_partition(edge_index, central_mask):
    """models/KTGNN.py:385-398: remove_self_loops -> add_self_loops (appended at the end) ->
    split by central_mask[dst].  Returns (ei1 [dst in source], ei2 [dst in target], cat)."""
    n = central_mask.shape[0]
    ei = edge_index[:, edge_index[0] != edge_index[1]]
    loop = torch.arange(n, dtype=ei.dtype).unsqueeze(0).repeat(2, 1)
    ei = torch.cat([ei, loop], dim=1)
    m1 = central_mask[ei[1]]
    ei1, ei2 = ei[:, m1], ei[:, ~m1]
    return ei1, ei2, torch.cat((ei1, ei2), dim=-1)


def adapted_conv(x, ei, ei1, ei2, central_mask, P, negative_slope=0.1, prefix=""):
    """models/KTGNN.py:263-315 (AdaptedConv.forward) + message :317-319, root_weight handled if
    ``lin_r.weight`` is present.  P: dict of parameter tensors keyed like the state_dict."""
    g = lambda k: P[prefix + k]
    n = x.shape[0]
    c = central_mask
    diff = x[c].mean(0, keepdim=True) - x[~c].mean(0, keepdim=True)
    diff = diff.expand(x.shape)
    cat = torch.cat((x, diff), dim=-1)
    shift_s2t = torch.tanh(F.linear(cat, g("a_g_s2t.weight"))) * diff
    shift_t2s = torch.tanh(F.linear(cat, g("a_g_t2s.weight"))) * diff
    x_s2t = x - shift_s2t * c.unsqueeze(-1)
    x_t2s = x + shift_t2s * (~c).unsqueeze(-1)
    x_s2t = F.linear(x_s2t, g("lin_t.weight"), g("lin_t.bias"))
    x_t2s = F.linear(x_t2s, g("lin_s.weight"), g("lin_s.bias"))
    a_t2s = F.leaky_relu(x_t2s[ei1[0]] + x_t2s[ei1[1]], negative_slope)
    a_s2t = F.leaky_relu(x_s2t[ei2[0]] + x_s2t[ei2[1]], negative_slope)
    alpha1 = F.linear(a_t2s, g("a_f_t2s.weight"))
    alpha2 = F.linear(a_s2t, g("a_f_s2t.weight"))
    alpha = pyg_softmax(torch.cat((alpha1, alpha2), 0), ei[1], n)
    e1 = alpha1.shape[0]
    out = _scatter_add(x_t2s[ei1[0]] * alpha[:e1], ei1[1], n) + _scatter_add(x_s2t[ei2[0]] * alpha[e1:], ei2[1], n)
    if (prefix + "lin_r.weight") in P:
        out = out + F.linear(x, g("lin_r.weight"))
    return out


def adapted_conv_aggregate(Hs, Ht, ei1, ei2, central_mask, af_t2s, af_s2t, negative_slope=0.1):
    """Only the edge part of AdaptedConv (models/KTGNN.py:292-305): scores, softmax over destination,
    weighted scatter-add.  Hs = lin_s(x_t2s), Ht = lin_t(x_s2t).  This is what the fused CUDA kernel computes."""
    n = Hs.shape[0]
    a1 = F.linear(F.leaky_relu(Hs[ei1[0]] + Hs[ei1[1]], negative_slope), af_t2s.view(1, -1))
    a2 = F.linear(F.leaky_relu(Ht[ei2[0]] + Ht[ei2[1]], negative_slope), af_s2t.view(1, -1))
    alpha = pyg_softmax(torch.cat((a1, a2), 0), torch.cat((ei1[1], ei2[1])), n)
    e1 = a1.shape[0]
    return _scatter_add(Hs[ei1[0]] * alpha[:e1], ei1[1], n) + _scatter_add(Ht[ei2[0]] * alpha[e1:], ei2[1], n)


def _bn(x, P, prefix, training):
    if training:
        return F.batch_norm(x, None, None, P[prefix + ".weight"], P[prefix + ".bias"], training=True, eps=1e-5)
    return F.batch_norm(x, P[prefix + ".running_mean"], P[prefix + ".running_var"], P[prefix + ".weight"],
                        P[prefix + ".bias"], training=False, eps=1e-5)


def ktgnn_no_complement(x, edge_index, central_mask, P, n_layers=2, training=False):
    """models/KTGNN.py:401-435 (forward), use_bn=True, need_complement=False, dropout treated as
    identity (eval mode, or train mode with p=0)."""
    ei1, ei2, ei = graph_partition(edge_index, central_mask)
    for i in range(n_layers - 1):
        x = adapted_conv(x, ei, ei1, ei2, central_mask, P, prefix=f"convs.{i}.")
        x = F.relu(_bn(x, P, f"bns.{i}", training))
    lb = adapted_conv(x, ei, ei1, ei2, central_mask, P, prefix="clf_base.")
    t = F.linear(x, P["clf_transformer.0.weight"], P["clf_transformer.0.bias"])
    t = F.relu(_bn(t, P, "clf_transformer.1", training))
    t = F.linear(t, P["clf_transformer.3.weight"], P["clf_transformer.3.bias"])
    ltt = adapted_conv(t, ei, ei1, ei2, central_mask, P, prefix="clf_target.")
    lt = adapted_conv(x, ei, ei1, ei2, central_mask, P, prefix="clf_target.")
    return F.log_softmax(lb, 1), F.log_softmax(lt, 1), F.log_softmax(ltt, 1)


def spmm(edge_index, x, n, reduce="sum", edge_weight=None):
    """torch_sparse.matmul(SparseTensor(row=dst, col=src), x, reduce) [upstream]
    (call: models/backbones.py:464-468 via SAGEConv.message_and_aggregate)."""
    msg = x.index_select(0, edge_index[0])
    if edge_weight is not None:
        msg = msg * edge_weight.view(-1, 1)
    out = _scatter_add(msg, edge_index[1], n)
    if reduce == "mean":
        cnt = _scatter_add(torch.ones(edge_index.shape[1], dtype=x.dtype), edge_index[1], n).clamp(min=1)
        out = out / cnt.view(-1, 1)
    return out


def sage_conv(x, edge_index, P, prefix):
    """PyG SAGEConv(mean) [upstream]: lin_l(mean_j x_j) + lin_r(x_i)."""
    agg = spmm(edge_index, x, x.shape[0], "mean")
    return F.linear(agg, P[prefix + "lin_l.weight"], P[prefix + "lin_l.bias"]) + F.linear(x, P[prefix + "lin_r.weight"])


def graphsage(x, edge_index, P, n_layers=2):
    """models/backbones.py:462-473 (GraphSAGE.forward), eval mode."""
    for i in range(n_layers):
        x = sage_conv(x, edge_index, P, f"convs.{i}.")
        if i != n_layers - 1:
            x = F.relu(x)
    return F.log_softmax(x, dim=1)


def gcn_norm(edge_index, n, dtype=torch.float32):
    """PyG gcn_norm [upstream]: add_remaining_self_loops(fill 1), deg by dst, d^-1/2[src] d^-1/2[dst]."""
    m = edge_index[0] != edge_index[1]
    loop = torch.arange(n, dtype=edge_index.dtype).unsqueeze(0).repeat(2, 1)
    ei = torch.cat([edge_index[:, m], loop], 1)
    w = torch.ones(ei.shape[1], dtype=dtype)
    deg = _scatter_add(w, ei[1], n)
    dis = deg.pow(-0.5)
    dis = dis.masked_fill(dis == float("inf"), 0)
    return ei, dis[ei[0]] * w * dis[ei[1]]


def gcn_conv(x, edge_index, P, prefix):
    ei, w = gcn_norm(edge_index, x.shape[0], x.dtype)
    h = F.linear(x, P[prefix + "lin.weight"])
    return spmm(ei, h, x.shape[0], "sum", w) + P[prefix + "bias"]


def gcn_net(x, edge_index, P, n_layers=2):
    """models/backbones.py:269-277 (GCNNet.forward), eval mode."""
    for i in range(n_layers):
        x = gcn_conv(x, edge_index, P, f"convs.{i}.")
        if i != n_layers - 1:
            x = F.relu(x)
    return F.log_softmax(x, dim=1)


def adapted_transform_epilogue(P, wd, kg, is_src, bias=None):
    """Node-wise epilogue of AdaptedConv after the single contraction P = x [W_s; W_t; a_g_s2t[:D]; a_g_t2s[:D]]^T + b
    (algebraically equal to models/KTGNN.py:277-284; equality with the reference's own op order is checked in
    tests/test_host_logic.py::test_adapted_conv_algebra_matches_reference).  Returns (Hs, Ht)."""
    c = (P.shape[1] - 2) // 2
    if bias is not None:            # (b_s, b_t) of lin_s / lin_t when the contraction was done without them
        P = P + torch.cat((bias.reshape(-1), bias.new_zeros(2)))
    g = torch.tanh(P[:, 2 * c:] + kg.view(1, 2))
    cf = is_src.to(P.dtype)
    wd = wd.reshape(-1)
    Hs = P[:, :c] + ((1.0 - cf) * g[:, 1]).unsqueeze(1) * wd[:c]
    Ht = P[:, c:2 * c] - (cf * g[:, 0]).unsqueeze(1) * wd[c:]
    return Hs, Ht


def eval_bridged_graph(edge_index, y, test_mask, n):
    """utils.py:101-113: per node, the label histogram of its in-neighbours (unlabelled neighbours, y = -1, are
    ignored), local homophily = share of the node's own label in it, and the fraction of test nodes whose local
    homophily exceeds 0.5."""
    adj = torch.zeros((n, n))
    adj.index_put_((edge_index[1], edge_index[0]), torch.ones(edge_index.shape[1]), accumulate=True)   # SparseTensor(row=dst, col=src), duplicates add
    y_onehot = F.one_hot(y + 1).float()[:, 1:]
    lbl_dist = adj @ y_onehot
    deg = lbl_dist.sum(1)
    nonzero = (lbl_dist.sum(1) != 0) & (y != -1)
    deg = deg + (~nonzero).float() * 1e-3
    local = (lbl_dist * y_onehot).sum(1) / deg
    return ((local[test_mask] > 0.5).sum() / test_mask.sum()).item(), local


def eval_homophily(edge_index, y, n):
    """utils.py:115-131: share of same-label pairs among the labelled edges, and among the labelled pairs of the
    2-hop pattern nonzero(A A) with A = SparseTensor(row=edge_index[0], col=edge_index[1]) (dense here)."""
    a = torch.zeros((n, n))
    a.index_put_((edge_index[0], edge_index[1]), torch.ones(edge_index.shape[1]), accumulate=True)
    two = torch.nonzero(a @ a, as_tuple=False).t()
    m1 = (y[edge_index[0]] != -1) & (y[edge_index[1]] != -1)
    r1 = ((y[edge_index[0]] == y[edge_index[1]]) & m1).sum() / m1.sum()
    m2 = (y[two[0]] != -1) & (y[two[1]] != -1)
    r2 = ((y[two[0]] == y[two[1]]) & m2).sum() / m2.sum()
    return r1.item(), r2.item()

