"""Graph diagnostics of the reference (Bridged-GNN/utils.py:101-131) without dense N x N matrices.

``eval_bridged_Graph`` is a label-histogram SpMM (the CSR SpMM kernel with F = number of classes);
``eval_homophily``'s 2-hop ratio needs the PATTERN of A A: it is enumerated per block of start nodes
(i -> j -> k through the transposed CSR), de-duplicated with a sort per block, and never materialised as a
matrix -- the reference's ``adj_sp.to_dense()`` (utils.py:121) is O(N^2) memory and stops near 5e4 nodes.
Both run on the device; there is no CPU path.
"""
import torch
import torch.nn.functional as F

from . import ops

__all__ = ["eval_bridged_Graph", "eval_homophily"]


def eval_bridged_Graph(data_merge, verbose=True):
    """utils.py:101-113.  Fraction of test nodes whose labelled in-neighbours mostly share their label; returns the
    ratio (0-dim tensor, like the reference).  Unlabelled neighbours (y = -1) are ignored."""
    x, edge_index, y = data_merge.x, data_merge.edge_index, data_merge.y
    n = x.shape[0]
    y_onehot = F.one_hot(y + 1).to(torch.float32)[:, 1:].contiguous()
    graph = ops.cached_graph(edge_index, n)                      # rows = destinations, like SparseTensor(row=dst, col=src)
    ones = torch.ones(edge_index.shape[1], dtype=torch.float32, device=edge_index.device)
    lbl_dist = ops.spmm(graph, y_onehot, "sum", ones)            # duplicate edges count twice, as in the reference
    deg = lbl_dist.sum(1)
    mask_deg_nonzero = (deg != 0) & (y != -1)
    deg = deg + (~mask_deg_nonzero).to(deg.dtype) * 1e-3
    local_homophily = (lbl_dist * y_onehot).sum(1) / deg
    test = data_merge.test_mask
    ratio = (local_homophily[test] > 0.5).sum() / test.sum()
    if verbose:
        print(ratio)
    return ratio


def _two_hop_counts(graph, y, max_pairs):
    """(#labelled distinct 2-hop pairs, #of those with equal labels) over the pattern nonzero(A A),
    A[i, j] = 1 for an edge i -> j.  Blocks of start nodes i are sized so that a block enumerates at most
    ``max_pairs`` (i, j, k) paths; all paths of a start node fall into one block, so de-duplicating per block
    is exact."""
    n = graph.n
    t_rowptr, t_col, _ = graph.t                                  # rows = sources, entries = destinations
    t_rowptr = t_rowptr.to(torch.int64)
    t_col = t_col.to(torch.int64)
    out_deg = t_rowptr[1:] - t_rowptr[:-1]
    # paths starting at i: sum over its out-neighbours j of out_deg[j]
    per_edge = out_deg[t_col]
    csum = torch.zeros(per_edge.numel() + 1, dtype=torch.int64, device=per_edge.device)
    csum[1:] = torch.cumsum(per_edge, 0)
    paths_from = csum[t_rowptr[1:]] - csum[t_rowptr[:-1]]         # [n]
    cum_nodes = torch.cumsum(paths_from, 0).cpu()
    labelled = same = 0
    i0 = 0
    while i0 < n:
        base = int(cum_nodes[i0 - 1]) if i0 > 0 else 0
        i1 = int(torch.searchsorted(cum_nodes, torch.tensor(base + max_pairs), right=True))
        i1 = min(n, max(i1, i0 + 1))                              # at least one start node per block
        e0, e1 = int(t_rowptr[i0]), int(t_rowptr[i1])
        if e1 > e0:
            j = t_col[e0:e1]
            cnt = out_deg[j]
            start_i = torch.repeat_interleave(torch.arange(i0, i1, device=j.device), out_deg[i0:i1])
            i_rep = torch.repeat_interleave(start_i, cnt)
            first = torch.repeat_interleave(t_rowptr[j], cnt)
            seg = torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt)
            k = t_col[first + (torch.arange(i_rep.numel(), device=j.device) - seg)]
            key = torch.unique(i_rep * n + k)
            yi, yk = y[key // n], y[key % n]
            lab = (yi != -1) & (yk != -1)
            labelled += int(lab.sum())
            same += int(((yi == yk) & lab).sum())
        i0 = i1
    return labelled, same


def eval_homophily(data, verbose=True, max_pairs=1 << 26):
    """utils.py:115-131.  Returns (homophily ratio over the labelled edges, the same over the labelled pairs of the
    2-hop pattern nonzero(A A)) as Python floats and prints them like the reference."""
    x, edge_index, y = data.x, data.edge_index, data.y
    n = x.shape[0]
    lab1 = (y[edge_index[0]] != -1) & (y[edge_index[1]] != -1)
    homo_ratio_1st = (((y[edge_index[0]] == y[edge_index[1]]) & lab1).sum() / lab1.sum()).item()
    graph = ops.cached_graph(edge_index, n)
    labelled, same = _two_hop_counts(graph, y, max_pairs)
    homo_ratio_2rd = same / labelled if labelled else float("nan")
    if verbose:
        print("homophily ratio:", homo_ratio_1st)
        print("homophily ratio 2rd neibors:", homo_ratio_2rd)
    return homo_ratio_1st, homo_ratio_2rd
