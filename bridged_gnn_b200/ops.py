"""Torch-facing operators over the C ABI: kNN build, CSR graphs, SpMM, the fused AdaptedConv aggregation
(single conv and multi-head), the node-wise transforms (CUDA-core for classifier heads, tcgen05 3 x TF32 for wide
outputs), row-panel / weight-gradient GEMMs and BatchNorm1d + ReLU -- all differentiable.  Everything here runs
hand-written sm_100a kernels on the current CUDA stream; nothing falls back to the CPU, and an operator given a
shape it does not cover raises (the ``*_supported`` predicates tell callers beforehand).  The one exception: the
backward contractions of ``linear`` / ``adapted_wide`` use ``torch.matmul`` on the GPU when an operand's row stride
is not a multiple of 16 bytes (TMA cannot address it) -- never the case for the widths KT-GNN uses.
"""
import os

import torch

from . import _lib
from ._lib import KNN_SIMT_F32, KNN_TC_1XTF32, KNN_TC_3XTF32, KNN_TC_F16  # noqa: F401

_ALGOS = {"simt": KNN_SIMT_F32, "tc3": KNN_TC_3XTF32, "tc1": KNN_TC_1XTF32, "f16": KNN_TC_F16}
# below this many pairs the tensor-core sweep cannot fill the machine and the CUDA-core sweep is used
_TC_MIN_PAIRS = 1 << 24


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


# ----------------------------------------------------------------------------------- kNN build
def knn_cosine(q, db, k, normalize=True, apply_sigmoid=True, algo="auto", eps=None):
    """Per-row top-k of sigmoid(cos(q_i, db_j)) without materialising the [nq, ndb] matrix.

    Replaces the pair enumeration + ``Similar.similarity*`` + ``sim_mat.topk`` of the reference
    (main_bridged_graph.py:45-67, 90-111; models/models.py:124-130, 945-948).  Selection key:
    fp32 similarity desc, db index asc.  Returns ``(idx int64 [nq,k], val fp32 [nq,k] best first,
    gap fp32 [nq] = v_k - v_(k+1), stats int32 [4])``; ``stats[0]`` = rows the tensor-core path
    re-did exactly.  ``q is db`` is the within-domain case (self matches are kept).

    ``eps`` (float): the bridge-matching threshold fused into the selection epilogue -- neighbours whose similarity
    is not > eps get ``idx = -1`` (their similarity stays in ``val``) and a fifth output ``count int32 [nq]`` gives
    the number of neighbours kept per row (a prefix: rows are best first).
    """
    lib = _lib.load()
    q, db = _f32c(q), (_f32c(db) if db is not q else None)
    if db is None:
        db = q
    nq, d = q.shape
    ndb = db.shape[0]
    if db.shape[1] != d:
        raise ValueError("feature widths differ")
    if algo == "auto":
        algo = "f16" if nq * ndb >= _TC_MIN_PAIRS else "simt"
    a = _ALGOS[algo]
    dev = q.device
    idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
    val = torch.empty((nq, k), dtype=torch.float32, device=dev)
    gap = torch.empty((nq,), dtype=torch.float32, device=dev)
    stats = torch.zeros((4,), dtype=torch.int32, device=dev)
    nbytes = lib.bgnn_knn_cosine_workspace_bytes(nq, ndb, d, k, a)
    if nbytes == 0:
        raise ValueError("invalid kNN arguments (need 1 <= k <= min(ndb, 255))")
    ws = _lib.workspace(nbytes, dev)
    count = None if eps is None else torch.empty((nq,), dtype=torch.int32, device=dev)
    with _lib.call("bgnn_knn_cosine_f32", "bgnn_knn_cosine_f32[simt]" if a == KNN_SIMT_F32 else None):
        _lib.check(lib.bgnn_knn_cosine_eps_f32(_lib.ptr(q), nq, _lib.ptr(db), ndb, d, k, int(normalize),
                                               int(apply_sigmoid), a, float("nan") if eps is None else float(eps),
                                               _lib.ptr(idx), _lib.ptr(val), _lib.ptr(gap), _lib.ptr(count, allow_none=True),
                                               _lib.ptr(stats), _lib.ptr(ws), ws.numel(), _lib.stream(dev)))
    return (idx, val, gap, stats) if eps is None else (idx, val, gap, stats, count)


def knn_addrelu(Uq, Udb, w2, b2, k, apply_sigmoid=True, eps=None):
    """Per-row top-k of sigmoid(sum_h w2[h] relu(Uq[i,h] + Udb[j,h]) + b2): the eval-mode fold of
    ``Similar_v2(mode='mlp')`` (models/models.py:918-925, 949-954).  Returns (idx, val, gap), plus ``count`` when
    ``eps`` is given (see knn_cosine)."""
    lib = _lib.load()
    Uq, w2 = _f32c(Uq), _f32c(w2).view(-1)
    Udb = Uq if Udb is Uq else _f32c(Udb)
    nq, h = Uq.shape
    ndb = Udb.shape[0]
    if Udb.shape[1] != h or w2.numel() != h:
        raise ValueError("feature widths differ")
    dev = Uq.device
    idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
    val = torch.empty((nq, k), dtype=torch.float32, device=dev)
    gap = torch.empty((nq,), dtype=torch.float32, device=dev)
    nbytes = lib.bgnn_knn_addrelu_workspace_bytes(nq, ndb, h, k)
    if nbytes == 0:
        raise ValueError("invalid kNN arguments (need 1 <= k <= min(ndb, 255))")
    ws = _lib.workspace(nbytes, dev)
    count = None if eps is None else torch.empty((nq,), dtype=torch.int32, device=dev)
    with _lib.call("bgnn_knn_addrelu_f32"):
        _lib.check(lib.bgnn_knn_addrelu_eps_f32(_lib.ptr(Uq), nq, _lib.ptr(Udb), ndb, h, _lib.ptr(w2), float(b2), k,
                                                int(apply_sigmoid), float("nan") if eps is None else float(eps),
                                                _lib.ptr(idx), _lib.ptr(val), _lib.ptr(gap),
                                                _lib.ptr(count, allow_none=True), _lib.ptr(ws), ws.numel(), _lib.stream(dev)))
    return (idx, val, gap) if eps is None else (idx, val, gap, count)


# ----------------------------------------------------------------------------------- graph format
def edges_to_csr(src, dst, n, dedup=False, want_perm=True):
    """(src[e], dst[e]) int64 -> destination-major CSR: rowptr int32 [n+1], col int32 [e'] (sources,
    rows sorted by (dst, src)), perm int64 [e'] (position in the input list), e' (python int)."""
    lib = _lib.load()
    src, dst = src.contiguous(), dst.contiguous()
    e = src.numel()
    dev = src.device
    rowptr = torch.empty((n + 1,), dtype=torch.int32, device=dev)
    col = torch.empty((max(e, 1),), dtype=torch.int32, device=dev)
    perm = torch.empty((max(e, 1),), dtype=torch.int64, device=dev) if want_perm else None
    e_out = torch.zeros((2,), dtype=torch.int64, device=dev)
    ws = _lib.workspace(lib.bgnn_edges_to_csr_workspace_bytes(e), dev)
    with _lib.call("bgnn_edges_to_csr"):
        _lib.check(lib.bgnn_edges_to_csr(_lib.ptr(src, torch.int64) if e else None,
                                         _lib.ptr(dst, torch.int64) if e else None, e, n, int(dedup), _lib.ptr(rowptr),
                                         _lib.ptr(col), _lib.ptr(perm, allow_none=True), _lib.ptr(e_out), _lib.ptr(ws),
                                         ws.numel(), _lib.stream(dev)))
    ne, bad = e_out.tolist()      # one sync per graph build (graphs are cached by the layers)
    if bad:
        raise IndexError("edge_index holds %d edge(s) with a node id outside [0, %d)" % (bad, n))
    return rowptr, col[:ne], (perm[:ne] if want_perm else None), ne


def coalesce(edge_index, num_nodes=None):
    """torch_geometric.utils.coalesce (main_bridged_graph.py:75, 113): sort by (row, col), drop duplicates."""
    if edge_index.numel() == 0:
        return edge_index
    n = int(edge_index.max().item()) + 1 if num_nodes is None else num_nodes
    rowptr, col, _, _ = edges_to_csr(edge_index[1], edge_index[0], n, dedup=True, want_perm=False)
    counts = (rowptr[1:] - rowptr[:-1]).to(torch.int64)
    row = torch.repeat_interleave(torch.arange(n, device=edge_index.device), counts)
    return torch.stack((row, col.to(torch.int64)), 0)


# ----------------------------------------------------------------------------------- edge-validity filters
def quantile(v, q):
    """``v.quantile(q)`` (linear interpolation, torch's float32 rank arithmetic) for a 1-D fp32 CUDA tensor by radix
    select: no sort and none of torch.quantile's 16 M-element cap (main_bridged_graph.py:134, 236).  Returns a
    0-dim device tensor; no host synchronisation."""
    import numpy as np
    lib = _lib.load()
    v = _f32c(v).view(-1)
    n = v.numel()
    if n == 0:
        raise ValueError("quantile of an empty tensor")
    if n - 1 < (1 << 24):      # torch: ranks = q * (n - 1) in the input's dtype
        rank = np.float32(q) * np.float32(n - 1)
        lo = int(np.floor(rank))
        w = float(np.float32(rank - np.float32(lo)))
    else:                      # beyond torch.quantile's own limit: double arithmetic
        rank = float(q) * (n - 1)
        lo = int(rank)
        w = rank - lo
    lo = min(max(lo, 0), n - 1)
    out = torch.empty((3,), dtype=torch.float32, device=v.device)
    ws = _lib.workspace(lib.bgnn_quantile_workspace_bytes(), v.device)
    with _lib.call("bgnn_quantile_f32"):
        _lib.check(lib.bgnn_quantile_f32(_lib.ptr(v), n, lo, w, _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream(v.device)))
    return out[2]


def edge_validity(edge_index, e_sim, thr_conf, pred_a, y_a, pred_b, y_b, gate_a, gate_b, x_a, x_b, thres_feat_sim):
    """The four removal rules of check_added_edges_{cross,within}_domain_validity (main_bridged_graph.py:225-264,
    123-161) in one kernel, warp per edge.  Returns (keep bool [E], counts int64 [5] = newly removed by rule 1..4,
    kept).  ``thr_conf``: 0-dim device tensor (``quantile``); gates: bool [n_b] or None."""
    lib = _lib.load()
    dev = x_a.device
    ei = edge_index.to(torch.int64).contiguous()
    e = ei.shape[1]
    i64, u8 = torch.int64, torch.uint8
    keep = torch.empty((max(e, 1),), dtype=u8, device=dev)
    counts = torch.empty((5,), dtype=i64, device=dev)
    x_a, x_b = _f32c(x_a), (_f32c(x_b) if x_b is not x_a else None)
    if x_b is None:
        x_b = x_a
    g_a = None if gate_a is None else gate_a.to(u8).contiguous()
    g_b = None if gate_b is None else gate_b.to(u8).contiguous()
    thr = None if thr_conf is None else thr_conf.to(torch.float32).reshape(1).contiguous()
    with _lib.call("bgnn_edge_validity_f32"):
        _lib.check(lib.bgnn_edge_validity_f32(_lib.ptr(ei[0].contiguous()) if e else None, _lib.ptr(ei[1].contiguous()) if e else None, e,
                                              _lib.ptr(_f32c(e_sim).view(-1)) if e else None, _lib.ptr(thr, allow_none=True),
                                              _lib.ptr(pred_a.to(i64).contiguous()), _lib.ptr(y_a.to(i64).contiguous()),
                                              _lib.ptr(pred_b.to(i64).contiguous()), _lib.ptr(y_b.to(i64).contiguous()),
                                              _lib.ptr(g_a, u8, True), _lib.ptr(g_b, u8, True), _lib.ptr(x_a), _lib.ptr(x_b),
                                              x_a.shape[1], float(thres_feat_sim), _lib.ptr(keep), _lib.ptr(counts),
                                              _lib.stream(dev)))
    return keep[:e].bool(), counts


WIDE_ROW = 32         # feature widths from here on (rows of >= 128 B) process rows in degree order


def rows_by_degree(rowptr, n, min_degree=0):
    """Row ids sorted by descending CSR row length (stable), int32 [n], on the device; rows shorter than
    ``min_degree`` follow in natural order."""
    lib = _lib.load()
    order = torch.empty((n,), dtype=torch.int32, device=rowptr.device)
    if n == 0:
        return order
    ws = _lib.workspace(lib.bgnn_rows_by_degree_workspace_bytes(n), rowptr.device)
    with _lib.call("bgnn_rows_by_degree"):
        _lib.check(lib.bgnn_rows_by_degree(_lib.ptr(rowptr, torch.int32), n, int(min_degree), _lib.ptr(order), _lib.ptr(ws), ws.numel(),
                                           _lib.stream(rowptr.device)))
    return order


class CSRGraph:
    """Destination-major CSR of an edge list plus, lazily, the CSR of the transposed graph (needed by
    the backward passes).  Built once per graph and cached by the layers, where the reference rebuilds
    a SparseTensor on every forward (models/backbones.py:464).

    Destination-partitioned (multi-GPU) form: ``n_rows`` / ``row_off`` say that the edge list holds only the edges
    whose destination lies in [row_off, row_off + n_rows) (global ids).  Rows are then LOCAL (0 .. n_rows-1), column
    entries stay global (0 .. n-1), and the transposed CSR has n rows (every source) with local destination entries."""

    def __init__(self, edge_index, num_nodes, n_rows=None, row_off=0):
        self.n = int(num_nodes)
        self.n_src = self.n
        self.n_rows = self.n if n_rows is None else int(n_rows)
        self.row_off = int(row_off)
        self.edge_index = edge_index
        self._dst = edge_index[1] if self.row_off == 0 else edge_index[1] - self.row_off
        rowptr, self.col, self.perm, self.e = edges_to_csr(edge_index[0], self._dst, self.n)
        if self.e and self.n_rows < self.n and int(rowptr[self.n_rows]) != self.e:
            raise IndexError("edge_index holds destinations outside [%d, %d)" % (self.row_off, self.row_off + self.n_rows))
        self.rowptr = rowptr[: self.n_rows + 1]
        self._t = None
        self._deg = None
        self._csr_to_csc = None
        self._order = None
        self._t_order = None

    @classmethod
    def prepared(cls, edge_index, num_nodes, rewrite_self_loops=True, training=True):
        """CSR, transposed CSR, CSR -> CSC slot map and both processing orders of a NEW edge list in one library call
        (``bgnn_graph_prepare``).  ``rewrite_self_loops``: graph_partition's ``add_self_loops(remove_self_loops(.))``
        (models/KTGNN.py:385-398) happens inside the key construction, so no filtered / concatenated copy of the edge
        list is made.  The result equals ``CSRGraph(graph_partition(edge_index, mask)[2], n)`` after ``prepare()``,
        except that ``perm`` (the map back to an input edge list) is None: the aggregation has no per-edge inputs."""
        lib = _lib.load()
        if not edge_index.is_cuda or edge_index.dtype != torch.int64:
            raise TypeError("CSRGraph.prepared needs an int64 CUDA edge_index")
        n, e = int(num_nodes), int(edge_index.shape[1])
        dev = edge_index.device
        src, dst = edge_index[0].contiguous(), edge_index[1].contiguous()
        cap = max(e + (n if rewrite_self_loops else 0), 1)
        i32 = torch.int32
        g = cls.__new__(cls)
        g.n = g.n_src = g.n_rows = n
        g.row_off = 0
        g.edge_index, g._dst, g.perm = None, None, None
        rowptr = torch.empty((n + 1,), dtype=i32, device=dev)
        col = torch.empty((cap,), dtype=i32, device=dev)
        order = torch.empty((n,), dtype=i32, device=dev)
        t_rowptr = t_col = c2c = t_order = None
        if training:
            t_rowptr = torch.empty((n + 1,), dtype=i32, device=dev)
            t_col = torch.empty((cap,), dtype=i32, device=dev)
            c2c = torch.empty((cap,), dtype=i32, device=dev)
            t_order = torch.empty((n,), dtype=i32, device=dev)
        e_out = torch.zeros((2,), dtype=torch.int64, device=dev)
        ws = _lib.workspace(lib.bgnn_graph_prepare_workspace_bytes(e, n, int(rewrite_self_loops)), dev)
        with _lib.call("bgnn_graph_prepare"):
            _lib.check(lib.bgnn_graph_prepare(_lib.ptr(src, torch.int64) if e else None, _lib.ptr(dst, torch.int64) if e else None,
                                              e, n, int(rewrite_self_loops), _lib.ptr(rowptr), _lib.ptr(col),
                                              _lib.ptr(t_rowptr, allow_none=True), _lib.ptr(t_col, allow_none=True),
                                              _lib.ptr(c2c, allow_none=True), _lib.ptr(order), _lib.ptr(t_order, allow_none=True),
                                              _lib.ptr(e_out), _lib.ptr(ws), ws.numel(), _lib.stream(dev)))
        ne, bad = e_out.tolist()      # one sync per graph build
        if bad:
            raise IndexError("edge_index holds %d edge(s) with a node id outside [0, %d)" % (bad, n))
        g.e = int(ne)
        g.rowptr, g.col = rowptr, col[: max(g.e, 1)]
        g._t = (t_rowptr, t_col[: max(g.e, 1)], None) if training else None
        g._csr_to_csc = c2c[: max(g.e, 1)] if training else None
        g._order, g._t_order = order, t_order
        g._deg = None
        return g

    @property
    def t(self):
        if self._t is None:
            if self.edge_index is None:
                raise RuntimeError("this graph was prepared without its transposed CSR (training=False)")
            self._t = edges_to_csr(self._dst, self.edge_index[0], self.n)[:3]
        return self._t

    @property
    def csr_to_csc(self):
        """int32 [e]: slot of every CSR edge in the transposed CSR (both index the same input edge list)."""
        if self._csr_to_csc is None:
            t_perm = self.t[2]
            inv_t = torch.empty_like(t_perm)
            inv_t[t_perm] = torch.arange(t_perm.numel(), device=t_perm.device)
            self._csr_to_csc = inv_t[self.perm].to(torch.int32).contiguous()
        return self._csr_to_csc

    def order(self, c=WIDE_ROW):
        """int32 [n] or None: processing order of the gather kernels for feature width c -- rows by descending
        in-degree (hub rows first, rows of similar length share a warp).  Narrow rows keep the natural order:
        measured on the sync-1M graph, the order lookup in front of every (short, latency-bound) row costs
        more than the hub tail it removes."""
        if c < WIDE_ROW:
            return None
        if self._order is None:
            self._order = rows_by_degree(self.rowptr, self.n_rows, 0)
        return self._order

    def t_order(self, c=WIDE_ROW):
        """The same for the transposed CSR (rows = sources, by descending out-degree)."""
        if c < WIDE_ROW:
            return None
        if self._t_order is None:
            self._t_order = rows_by_degree(self.t[0], self.n, 0)
        return self._t_order

    def prepare(self, training=True, wide=True):
        """Builds now what the aggregation kernels would build lazily on first use: the processing orders and, for
        training, the transposed CSR and the CSR -> CSC slot map.  Needs only the edge list, so a caller can run it
        while the node features are still on their way to the device."""
        if wide:
            self.order(WIDE_ROW)
        if training:
            self.csr_to_csc          # builds self.t too
            if wide:
                self.t_order(WIDE_ROW)
        return self

    @property
    def deg(self):
        if self._deg is None:
            self._deg = (self.rowptr[1:] - self.rowptr[:-1]).to(torch.float32)     # [n_rows]
        return self._deg


_graph_cache = {}


def _graph_key(edge_index, num_nodes, n_rows=None, row_off=0):
    return (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, int(num_nodes), str(edge_index.device),
            n_rows, row_off)


def register_graph(edge_index, num_nodes, graph):
    """Makes ``cached_graph(edge_index, num_nodes)`` return ``graph`` (a CSRGraph built by other means for this very
    edge list, e.g. ``CSRGraph.prepared`` of the un-partitioned list)."""
    if len(_graph_cache) >= 4:
        _graph_cache.pop(next(iter(_graph_cache)))
    graph._pinned_edges = edge_index          # the key holds an address: keep the tensor alive with the entry
    _graph_cache[_graph_key(edge_index, num_nodes)] = graph


def cached_graph(edge_index, num_nodes, n_rows=None, row_off=0):
    if isinstance(edge_index, CSRGraph):
        return edge_index
    key = _graph_key(edge_index, num_nodes, n_rows, row_off)
    g = _graph_cache.get(key)
    if g is None:
        if len(_graph_cache) >= 4:      # a handful of live graphs; each entry pins its edge list and two CSRs
            _graph_cache.pop(next(iter(_graph_cache)))
        g = CSRGraph(edge_index, num_nodes, n_rows, row_off)
        _graph_cache[key] = g
    return g


# ----------------------------------------------------------------------------------- SpMM
def _spmm_raw(rowptr, col, X, n_rows, reduce_mean=False, edge_w=None, gather_scale=None, out_scale=None, out=None):
    """X [*, f] and ``out`` [n_rows, f] may be column panels (row-strided views) of wider matrices."""
    lib = _lib.load()
    f32 = torch.float32
    if X.stride(1) != 1:
        X = X.contiguous()
    f = X.shape[1]
    Y = torch.empty((n_rows, f), dtype=f32, device=X.device) if out is None else out
    if X.dtype != f32 or Y.dtype != f32 or not X.is_cuda or Y.stride(1) != 1:
        raise TypeError("spmm needs fp32 CUDA matrices with unit column stride")
    with _lib.call("bgnn_spmm_csr_f32"):
        _lib.check(lib.bgnn_spmm_csr_ld_f32(_lib.ptr(rowptr, torch.int32), _lib.ptr(col, torch.int32),
                                            _lib.ptr(edge_w, f32, True), _lib.ptr(gather_scale, f32, True),
                                            _lib.ptr(out_scale, f32, True), X.data_ptr(), X.stride(0), n_rows, f,
                                            int(reduce_mean), Y.data_ptr(), Y.stride(0), _lib.stream(X.device)))
    return Y


class _SpmmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, graph, reduce, edge_weight, gather_scale, out_scale):
        ctx.graph, ctx.reduce, ctx.edge_weight = graph, reduce, edge_weight
        ctx.gather_scale, ctx.out_scale = gather_scale, out_scale
        ew = None if edge_weight is None else edge_weight[graph.perm].contiguous()
        return _spmm_raw(graph.rowptr, graph.col, X.to(torch.float32), graph.n_rows, reduce == "mean", ew, gather_scale,
                         out_scale)

    @staticmethod
    def backward(ctx, gY):
        # dX[j] = gather_scale[j] * sum_{i : j->i} w_ji * out_scale[i] * (1/deg_i) * dY[i]: the same kernel on
        # the transposed CSR with the per-row factors moved to the gather side.
        g = ctx.graph
        t_rowptr, t_col, t_perm = g.t
        ew = None if ctx.edge_weight is None else ctx.edge_weight[t_perm].contiguous()
        gs = ctx.out_scale
        if ctx.reduce == "mean":
            inv = 1.0 / g.deg.clamp(min=1.0)
            gs = inv if gs is None else gs * inv
        gX = _spmm_raw(t_rowptr, t_col, gY.to(torch.float32).contiguous(), g.n, False, ew,
                       None if gs is None else gs.contiguous(), ctx.gather_scale)
        return gX, None, None, None, None, None


def spmm(graph, X, reduce="sum", edge_weight=None, gather_scale=None, out_scale=None):
    """Y[i] = out_scale[i] * reduce_{j->i} w_ji * gather_scale[j] * X[j] over a CSRGraph.  Replaces
    torch_sparse.matmul(adj_t, x, reduce) (models/backbones.py:464-468) and the gather/scatter of
    SAGEConv / GCNConv; differentiable in X.  ``edge_weight`` is given in the order of the graph's
    original edge list; the scales are per node."""
    return _SpmmFn.apply(X, graph, reduce, edge_weight, gather_scale, out_scale)


# ----------------------------------------------------------------------------------- fused AdaptedConv aggregation
def _pick(Hs, Ht):
    """(Hs, Ht) with a missing one replaced by the other as a never-read stand-in (a destination-partitioned rank that
    owns no destination row of a domain is given no H of that domain)."""
    if Hs is None and Ht is None:
        raise ValueError("at least one of Hs, Ht is needed")
    return (Ht if Hs is None else Hs), (Hs if Ht is None else Ht)


class _GatAggFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Hs, Ht, af_t2s, af_s2t, graph, dst_is_src, slope):
        lib = _lib.load()
        f32 = torch.float32
        has = (Hs is not None, Ht is not None)
        Hs, Ht = _pick(Hs, Ht)
        Hs, Ht = Hs.to(f32).contiguous(), Ht.to(f32).contiguous()
        a1, a2 = af_t2s.to(f32).contiguous().view(-1), af_s2t.to(f32).contiguous().view(-1)
        c = Hs.shape[1]
        if Hs.shape[0] != graph.n_src or Ht.shape[0] != graph.n_src:
            raise ValueError("H must have one row per node of the graph (%d)" % graph.n_src)
        n = graph.n_rows
        dev = Hs.device
        out = torch.empty((n, c), dtype=f32, device=dev)
        row_max = torch.empty((n,), dtype=f32, device=dev)
        row_sum = torch.empty((n,), dtype=f32, device=dev)
        # training: the per-edge scores stay for the backward (4 B per edge instead of a recomputation per edge)
        keep = any(ctx.needs_input_grad[:4])
        score = torch.empty((max(graph.e, 1),), dtype=f32, device=dev) if keep else None
        with _lib.call("bgnn_gatv2_fwd_part_f32", "bgnn_gatv2_fwd_f32[c=%d]" % c):
            _lib.check(lib.bgnn_gatv2_fwd_part_f32(_lib.ptr(graph.rowptr), _lib.ptr(graph.col),
                                                   _lib.ptr(graph.order(c), torch.int32, True),
                                                   _lib.ptr(dst_is_src, torch.uint8), _lib.ptr(Hs), _lib.ptr(Ht),
                                                   _lib.ptr(a1), _lib.ptr(a2), float(slope), n, graph.row_off, c,
                                                   _lib.ptr(out), _lib.ptr(row_max), _lib.ptr(row_sum),
                                                   _lib.ptr(score, allow_none=True), _lib.stream(dev)))
        ctx.save_for_backward(Hs, Ht, a1, a2, out, row_max, row_sum, score)
        ctx.graph, ctx.dst_is_src, ctx.slope, ctx.has = graph, dst_is_src, float(slope), has
        ctx.a_shapes = (af_t2s.shape, af_s2t.shape)
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        Hs, Ht, a1, a2, out, row_max, row_sum, score = ctx.saved_tensors
        g = ctx.graph
        t_rowptr, t_col, _ = g.t
        c = Hs.shape[1]
        dev = Hs.device
        gout = gout.to(torch.float32).contiguous()
        gHs = torch.empty_like(Hs) if ctx.has[0] else None
        gHt = torch.empty_like(Ht) if ctx.has[1] else None
        ga1, ga2 = torch.empty_like(a1), torch.empty_like(a2)
        ws = _lib.workspace(lib.bgnn_gatv2_bwd_workspace_bytes(g.n_src, g.e, c), dev)
        with _lib.call("bgnn_gatv2_bwd_part_f32", "bgnn_gatv2_bwd_f32[c=%d]" % c):
            _lib.check(lib.bgnn_gatv2_bwd_part_f32(_lib.ptr(g.rowptr), _lib.ptr(g.col), _lib.ptr(t_rowptr), _lib.ptr(t_col),
                                                   _lib.ptr(g.csr_to_csc, torch.int32), _lib.ptr(g.order(c), torch.int32, True),
                                                   _lib.ptr(g.t_order(c), torch.int32, True), g.e,
                                                   _lib.ptr(ctx.dst_is_src), _lib.ptr(Hs), _lib.ptr(Ht), _lib.ptr(a1),
                                                   _lib.ptr(a2), ctx.slope, g.n_rows, g.row_off, g.n_src, c, _lib.ptr(out),
                                                   _lib.ptr(row_max), _lib.ptr(row_sum), _lib.ptr(score, allow_none=True),
                                                   _lib.ptr(gout), _lib.ptr(gHs, allow_none=True),
                                                   _lib.ptr(gHt, allow_none=True), _lib.ptr(ga1), _lib.ptr(ga2), _lib.ptr(ws),
                                                   ws.numel(), _lib.stream(dev)))
        return gHs, gHt, ga1.view(ctx.a_shapes[0]), ga2.view(ctx.a_shapes[1]), None, None, None


def gat_aggregate(Hs, Ht, af_t2s, af_s2t, graph, dst_is_src, slope=0.1):
    """Edge part of AdaptedConv (models/KTGNN.py:292-305 + message :317-319) as one fused kernel per
    direction of autograd: scores a.leaky_relu(H[src]+H[dst]), softmax over each destination's incoming
    edges (PyG softmax, +1e-16), weighted sum of H[src]; (H, a) = (Hs, af_t2s) for destinations in the
    source domain, (Ht, af_s2t) otherwise.  dst_is_src: uint8 [n].

    With a destination-partitioned ``graph`` (``CSRGraph(..., n_rows, row_off)``) Hs / Ht / dst_is_src cover ALL nodes,
    the result the rank's own ``n_rows`` destination rows; Hs or Ht may be None on a rank without destination rows of
    that domain."""
    return _GatAggFn.apply(Hs, Ht, af_t2s, af_s2t, graph, dst_is_src, slope)


class _GatHeadsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Hs, Ht, af_t2s, af_s2t, graph, dst_is_src, slope, heads):
        lib = _lib.load()
        f32 = torch.float32
        has = (Hs is not None, Ht is not None)
        Hs, Ht = _pick(Hs, Ht)
        Hs, Ht = Hs.to(f32).contiguous(), Ht.to(f32).contiguous()
        a1, a2 = af_t2s.to(f32).contiguous().view(-1), af_s2t.to(f32).contiguous().view(-1)
        f = Hs.shape[1]
        c = f // heads
        n = graph.n_rows
        dev = Hs.device
        out = torch.empty((n, f), dtype=f32, device=dev)
        row_max = torch.empty((n, heads), dtype=f32, device=dev)
        row_sum = torch.empty((n, heads), dtype=f32, device=dev)
        with _lib.call("bgnn_gatv2_heads_fwd_part_f32", "bgnn_gatv2_heads_fwd_f32[%dx%d]" % (heads, c)):
            _lib.check(lib.bgnn_gatv2_heads_fwd_part_f32(_lib.ptr(graph.rowptr), _lib.ptr(graph.col),
                                                         _lib.ptr(dst_is_src, torch.uint8), _lib.ptr(Hs), _lib.ptr(Ht),
                                                         _lib.ptr(a1), _lib.ptr(a2), float(slope), n, graph.row_off, heads, c,
                                                         _lib.ptr(out), _lib.ptr(row_max), _lib.ptr(row_sum), _lib.stream(dev)))
        ctx.save_for_backward(Hs, Ht, a1, a2, out, row_max, row_sum)
        ctx.graph, ctx.dst_is_src, ctx.slope, ctx.heads, ctx.has = graph, dst_is_src, float(slope), heads, has
        ctx.a_shapes = (af_t2s.shape, af_s2t.shape)
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        Hs, Ht, a1, a2, out, row_max, row_sum = ctx.saved_tensors
        g, heads = ctx.graph, ctx.heads
        t_rowptr, t_col, _ = g.t
        f = Hs.shape[1]
        c = f // heads
        dev = Hs.device
        gout = gout.to(torch.float32).contiguous()
        gHs = torch.empty_like(Hs) if ctx.has[0] else None
        gHt = torch.empty_like(Ht) if ctx.has[1] else None
        ga1, ga2 = torch.empty_like(a1), torch.empty_like(a2)
        ws = _lib.workspace(lib.bgnn_gatv2_heads_bwd_workspace_bytes(g.n_src, g.e, heads, c), dev)
        with _lib.call("bgnn_gatv2_heads_bwd_part_f32", "bgnn_gatv2_heads_bwd_f32[%dx%d]" % (heads, c)):
            _lib.check(lib.bgnn_gatv2_heads_bwd_part_f32(_lib.ptr(g.rowptr), _lib.ptr(g.col), _lib.ptr(t_rowptr), _lib.ptr(t_col),
                                                         _lib.ptr(g.csr_to_csc, torch.int32), g.e, _lib.ptr(ctx.dst_is_src),
                                                         _lib.ptr(Hs), _lib.ptr(Ht), _lib.ptr(a1), _lib.ptr(a2), ctx.slope,
                                                         g.n_rows, g.row_off, g.n_src, heads, c, _lib.ptr(out),
                                                         _lib.ptr(row_max), _lib.ptr(row_sum), _lib.ptr(gout),
                                                         _lib.ptr(gHs, allow_none=True), _lib.ptr(gHt, allow_none=True),
                                                         _lib.ptr(ga1), _lib.ptr(ga2), _lib.ptr(ws), ws.numel(),
                                                         _lib.stream(dev)))
        return gHs, gHt, ga1.view(ctx.a_shapes[0]), ga2.view(ctx.a_shapes[1]), None, None, None, None


def gat_heads_supported(heads, c):
    return bool(_lib.load().bgnn_gatv2_heads_supported(int(heads), int(c)))


def gat_aggregate_heads(Hs, Ht, af_t2s, af_s2t, graph, dst_is_src, slope=0.1, heads=3):
    """``heads`` narrow AdaptedConv aggregations over the same graph in one pass (KT-GNN's classifier convs,
    models/KTGNN.py:432-434).  Hs, Ht [n, heads*c] with head h in columns h*c .. h*c+c-1, af_* [heads*c]; per head
    identical to ``gat_aggregate`` (scores, softmax and sums never mix heads).  heads in {2, 3}, c <= 4."""
    return _GatHeadsFn.apply(Hs, Ht, af_t2s, af_s2t, graph, dst_is_src, slope, heads)


# ----------------------------------------------------------------------------------- AdaptedConv node-wise epilogue
class _AdaptedTransformFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, P, wd, kg, is_src, bias):
        lib = _lib.load()
        f32 = torch.float32
        P = P.to(f32).contiguous()
        wd_c, kg_c = wd.to(f32).contiguous().view(-1), kg.to(f32).contiguous().view(-1)
        b_c = None if bias is None else bias.to(f32).contiguous().view(-1)
        n = P.shape[0]
        c = (P.shape[1] - 2) // 2
        dev = P.device
        Hs = torch.empty((n, c), dtype=f32, device=dev)
        Ht = torch.empty((n, c), dtype=f32, device=dev)
        gates = torch.empty((n, 2), dtype=f32, device=dev)
        with _lib.call("bgnn_adapted_transform_fwd_f32"):
            _lib.check(lib.bgnn_adapted_transform_fwd_f32(_lib.ptr(P), _lib.ptr(is_src, torch.uint8), _lib.ptr(wd_c),
                                                          _lib.ptr(kg_c), _lib.ptr(b_c, f32, True), n, c, _lib.ptr(Hs),
                                                          _lib.ptr(Ht), _lib.ptr(gates), _lib.stream(dev)))
        ctx.save_for_backward(gates, wd_c, is_src)
        ctx.shapes = (wd.shape, kg.shape, c, None if bias is None else bias.shape)
        ctx.mark_non_differentiable(gates)
        return Hs, Ht, gates

    @staticmethod
    def backward(ctx, gHs, gHt, _ggates):
        lib = _lib.load()
        gates, wd_c, is_src = ctx.saved_tensors
        wd_shape, kg_shape, c, bias_shape = ctx.shapes
        f32 = torch.float32
        gHs, gHt = gHs.to(f32).contiguous(), gHt.to(f32).contiguous()
        n = gHs.shape[0]
        dev = gHs.device
        gP = torch.empty((n, 2 * c + 2), dtype=f32, device=dev)
        red = torch.empty((4 * c + 2,), dtype=f32, device=dev)
        ws = _lib.workspace(lib.bgnn_adapted_transform_bwd_workspace_bytes(c), dev)
        with _lib.call("bgnn_adapted_transform_bwd_f32"):
            _lib.check(lib.bgnn_adapted_transform_bwd_f32(_lib.ptr(gHs), _lib.ptr(gHt), _lib.ptr(gates),
                                                          _lib.ptr(is_src, torch.uint8), _lib.ptr(wd_c), n, c, 2 * c + 2,
                                                          _lib.ptr(gP), _lib.ptr(red), _lib.ptr(ws), ws.numel(),
                                                          _lib.stream(dev)))
        g_bias = None if bias_shape is None else red[2 * c + 2:].view(bias_shape)
        return gP, red[: 2 * c].view(wd_shape), red[2 * c: 2 * c + 2].view(kg_shape), None, g_bias


def adapted_transform(P, wd, kg, is_src, bias=None):
    """Fused node-wise epilogue of AdaptedConv (models/KTGNN.py:277-284 after the single contraction
    P = x [W_s; W_t; a_g_s2t[:D]; a_g_t2s[:D]]^T): returns (Hs, Ht) = (lin_s(x_t2s), lin_t(x_s2t)).
    P [n, 2c+2]; wd [2c] or [1, 2c] = (W_s Delta, W_t Delta); kg [2] = Delta part of the two gate logits;
    is_src uint8 [n]; bias [2c] = (b_s, b_t) or None (added here rather than by a separate pass over P).
    Differentiable in P, wd, kg, bias."""
    Hs, Ht, _ = _AdaptedTransformFn.apply(P, wd, kg, is_src, bias)
    return Hs, Ht


# ----------------------------------------------------------------------------------- wide AdaptedConv transform
def tf32_planes(w, rows_to=16, cols_to=32):
    """(hi, lo) planes of a small fp32 matrix for the 3 x TF32 contractions: hi = w rounded to tf32, lo = w - hi
    rounded to tf32 (the tensor core would truncate), both zero padded to multiples of (16, 32).  One kernel launch on
    the GPU (any strides, e.g. ``weight.t()``); plain torch ops for CPU tensors (tests)."""
    w = w.detach().to(torch.float32)
    r, c = w.shape
    rp, cp = -(-r // rows_to) * rows_to, -(-c // cols_to) * cols_to
    if w.is_cuda:
        lib = _lib.load()
        hi = torch.empty((rp, cp), dtype=torch.float32, device=w.device)
        lo = torch.empty((rp, cp), dtype=torch.float32, device=w.device)
        with _lib.call("bgnn_tf32_planes_f32"):
            _lib.check(lib.bgnn_tf32_planes_f32(w.data_ptr(), r, c, w.stride(0), w.stride(1), rp, cp, _lib.ptr(hi), _lib.ptr(lo),
                                                _lib.stream(w.device)))
        return hi, lo
    wp = torch.zeros((rp, cp), dtype=torch.float32, device=w.device)
    wp[:r, :c] = w

    def tf32(t):
        return ((t.view(torch.int32) + 4096) & -8192).view(torch.float32)
    hi = tf32(wp)
    return hi.contiguous(), tf32(wp - hi).contiguous()


def rowpanel_gemm_supported(k, ld_a, no):
    return bool(_lib.load().bgnn_rowpanel_gemm_supported(int(k), int(ld_a), int(no)))


_ACTS = {None: 0, "none": 0, "relu": 1, "tanh": 2}


def rowpanel_gemm(A, B, bias=None, scale=None, act=None, res=None):
    """act(A [n, k] @ B [no, k]^T * scale + bias) + res on the tcgen05 tensor cores with fp32-grade accuracy (3 x TF32; A is
    split on chip and streamed from HBM once), the epilogue applied to the accumulator tile.  Stands in for the fp32 SIMT
    GEMMs of AdaptedConv's node-wise part (models/KTGNN.py:277-284) and, with the epilogue, for the dense layers of the
    embedding producers (models/models.py:852-893, 70-99, 1092-1096).  scale / bias [no]; act None | "relu" | "tanh";
    res [n, no].  A may be a row-strided view (stride % 4 == 0).  Not differentiable."""
    lib = _lib.load()
    f32 = torch.float32
    A = A.detach().to(f32)
    n, k = A.shape
    if A.stride(1) != 1 or A.stride(0) % 4 != 0 or A.stride(0) < k or A.data_ptr() % 16 != 0:
        buf = torch.zeros((n, k + (-k) % 4), dtype=f32, device=A.device)     # TMA needs 16-byte row strides
        buf[:, :k] = A
        A = buf[:, :k]
    no = B.shape[0]
    hi, lo = tf32_planes(B)
    b_c = None if bias is None else bias.detach().to(f32).contiguous()
    s_c = None if scale is None else scale.detach().to(f32).contiguous()
    r_c = None if res is None else res.detach().to(f32).contiguous()
    if r_c is not None and tuple(r_c.shape) != (n, no):
        raise ValueError("res must be [n, no]")
    Y = torch.empty((n, no), dtype=f32, device=A.device)
    with _lib.call("bgnn_rowpanel_gemm_f32"):
        _lib.ptr(A[:1])          # device / dtype checks; A itself may be a row-strided view
        _lib.check(lib.bgnn_rowpanel_gemm_act_f32(A.data_ptr(), n, k, max(A.stride(0), k + (-k) % 4), _lib.ptr(hi), _lib.ptr(lo),
                                                  _lib.ptr(s_c, f32, True), _lib.ptr(b_c, f32, True), _ACTS[act],
                                                  _lib.ptr(r_c, f32, True), no, no, _lib.ptr(Y), no, _lib.stream(A.device)))
    return Y


def dense_supported(x, in_features, out_features):
    """Whether ``rowpanel_gemm`` takes a dense layer [*, in] -> [*, out] on this input (fp32 CUDA matrix, widths <= 256)."""
    return (torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and not os.environ.get("BGNN_NO_WIDE")
            and rowpanel_gemm_supported(in_features, in_features + (-in_features) % 4, out_features))


def _tma_rows(t):
    """Row-strided fp32 matrix a TMA descriptor can address in place (unit column stride, 16-byte rows and base)."""
    return (t.dim() == 2 and t.dtype == torch.float32 and t.is_cuda and t.stride(1) == 1 and t.stride(0) % 4 == 0
            and t.stride(0) >= t.shape[1] and t.data_ptr() % 16 == 0)


def wgrad_gemm_supported(G, X):
    return (_tma_rows(G) and _tma_rows(X) and G.shape[0] == X.shape[0]
            and bool(_lib.load().bgnn_wgrad_gemm_supported(X.shape[1], X.stride(0), G.shape[1], G.stride(0))))


def wgrad_gemm(G, X, colsum=False):
    """G [n, no]^T @ X [n, d] -> [no, d] on the tcgen05 tensor cores with fp32-grade accuracy (3 x TF32, both
    operands split on chip and read from HBM once, deterministic): the weight gradients of AdaptedConv's dense
    contraction (models/KTGNN.py:277-284) and of the Linear layers of clf_transformer (models/KTGNN.py:363).
    Both operands may be row-strided views (see ``_tma_rows``).  ``colsum=True`` (d <= 96) also returns G.sum(0),
    collected by the same pass through an all-ones feature.  Not differentiable."""
    lib = _lib.load()
    G, X = G.detach(), X.detach()
    if not wgrad_gemm_supported(G, X):
        raise ValueError("wgrad_gemm: unsupported operands (need fp32 CUDA, d <= 128, no <= 256, 16-byte rows)")
    n, no = G.shape
    d = X.shape[1]
    if colsum and d > 96:
        raise ValueError("wgrad_gemm: colsum needs d <= 96")
    W = torch.empty((no, d), dtype=torch.float32, device=G.device)
    cs = torch.empty((no,), dtype=torch.float32, device=G.device) if colsum else None
    ws = _lib.workspace(lib.bgnn_wgrad_gemm_workspace_bytes(no), G.device)
    with _lib.call("bgnn_wgrad_gemm_f32"):
        _lib.check(lib.bgnn_wgrad_gemm_f32(G.data_ptr(), G.stride(0), no, X.data_ptr(), X.stride(0), d, n, _lib.ptr(W), d,
                                           _lib.ptr(cs, allow_none=True), _lib.ptr(ws), ws.numel(), _lib.stream(G.device)))
    return (W, cs) if colsum else W


def wgrad_gemm_cat_supported(blocks, X):
    if not (1 <= len(blocks) <= 3 and _tma_rows(X) and all(_tma_rows(g) and g.shape[0] == X.shape[0] for g in blocks)):
        return False
    if any(g.shape[1] % 32 != 0 for g in blocks[:-1]):
        return False
    no = sum(g.shape[1] for g in blocks)
    return bool(_lib.load().bgnn_wgrad_gemm_supported(X.shape[1], X.stride(0), no, 4))


def wgrad_gemm_cat(blocks, X):
    """``torch.cat(blocks, 1).t() @ X`` without the concatenation: up to three column blocks [n, no_i] (all but the
    last a multiple of 32 wide), each read in place through its own TMA descriptor by the ``wgrad_gemm`` kernel."""
    lib = _lib.load()
    blocks = [g.detach() for g in blocks]
    X = X.detach()
    if not wgrad_gemm_cat_supported(blocks, X):
        raise ValueError("wgrad_gemm_cat: unsupported operands")
    n, d = X.shape
    no = sum(g.shape[1] for g in blocks)
    W = torch.empty((no, d), dtype=torch.float32, device=X.device)
    ws = _lib.workspace(lib.bgnn_wgrad_gemm_workspace_bytes(no), X.device)
    args = []
    for i in range(3):
        if i < len(blocks):
            args += [blocks[i].data_ptr(), blocks[i].stride(0), blocks[i].shape[1]]
        else:
            args += [None, 0, 0]
    with _lib.call("bgnn_wgrad_gemm_cat_f32"):
        _lib.check(lib.bgnn_wgrad_gemm_cat_f32(*args, X.data_ptr(), X.stride(0), d, n, _lib.ptr(W), d, None, _lib.ptr(ws),
                                               ws.numel(), _lib.stream(X.device)))
    return W


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return rowpanel_gemm(x, weight, bias)

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        gy = gy.to(torch.float32).contiguous()
        g_x = rowpanel_gemm(gy, weight.t()) if ctx.needs_input_grad[0] else None
        g_w = g_b = None
        want_b = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1]:
            xd = x.detach()
            if wgrad_gemm_supported(gy, xd):
                if want_b and xd.shape[1] <= 96:
                    g_w, g_b = wgrad_gemm(gy, xd, colsum=True)      # bias gradient from the same pass
                else:
                    g_w = wgrad_gemm(gy, xd)
            else:
                g_w = gy.t() @ xd
        if want_b and g_b is None:
            g_b = gy.sum(0)
        return g_x, g_w, g_b


def linear_supported(x, weight):
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and not os.environ.get("BGNN_NO_WIDE")
            and rowpanel_gemm_supported(weight.shape[1], weight.shape[1] + (-weight.shape[1]) % 4, weight.shape[0])
            and rowpanel_gemm_supported(weight.shape[0], weight.shape[0] + (-weight.shape[0]) % 4, weight.shape[1]))


def linear(x, weight, bias=None):
    """torch.nn.functional.linear for a tall node-feature matrix (n ~ 1e6 rows, <= 256 features in and out): forward
    and input gradient on the row-panel 3 x TF32 tensor-core GEMM (bias fused), weight gradient on ``wgrad_gemm``.
    The Linear layers of KT-GNN's clf_transformer (models/KTGNN.py:363).  Differentiable in x, weight, bias."""
    return _LinearFn.apply(x, weight, bias)


class _AdaptedWideFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w_cat, bias, wd, kg, is_src):
        lib = _lib.load()
        f32 = torch.float32
        x, w_cat = x.to(f32).contiguous(), w_cat.to(f32).contiguous()
        wd_c, kg_c = wd.to(f32).contiguous().view(-1), kg.to(f32).contiguous().view(-1)
        b_c = None if bias is None else bias.to(f32).contiguous().view(-1)
        n, d = x.shape
        c = (w_cat.shape[0] - 2) // 2
        dev = x.device
        hi, lo = tf32_planes(w_cat)
        Hs = torch.empty((n, c), dtype=f32, device=dev)
        Ht = torch.empty((n, c), dtype=f32, device=dev)
        gates = torch.empty((n, 2), dtype=f32, device=dev)
        with _lib.call("bgnn_adapted_wide_fwd_f32"):
            _lib.check(lib.bgnn_adapted_wide_fwd_f32(_lib.ptr(x), n, d, _lib.ptr(hi), _lib.ptr(lo), c,
                                                     _lib.ptr(is_src, torch.uint8), _lib.ptr(wd_c), _lib.ptr(kg_c),
                                                     _lib.ptr(b_c, f32, True), _lib.ptr(Hs), _lib.ptr(Ht), _lib.ptr(gates),
                                                     _lib.stream(dev)))
        ctx.save_for_backward(x, w_cat, wd_c, gates, is_src)
        ctx.meta = (wd.shape, kg.shape, c, None if bias is None else bias.shape)
        return Hs, Ht

    @staticmethod
    def backward(ctx, gHs, gHt):
        lib = _lib.load()
        x, w_cat, wd_c, gates, is_src = ctx.saved_tensors
        wd_shape, kg_shape, c, bias_shape = ctx.meta
        f32 = torch.float32
        gHs, gHt = gHs.to(f32).contiguous(), gHt.to(f32).contiguous()
        n = gHs.shape[0]
        dev = gHs.device
        if not ctx.needs_input_grad[0] and ctx.needs_input_grad[1]:
            # first layer (x is data): dP = [dHs | dHt | d gates] is never materialised -- one pass over (dHs, dHt)
            # for the reductions and the two gate columns, then the weight-gradient GEMM reads the three blocks in place
            dg = torch.empty((n, 4), dtype=f32, device=dev)
            red = torch.empty((4 * c + 2,), dtype=f32, device=dev)
            ws = _lib.workspace(lib.bgnn_adapted_transform_bwd_workspace_bytes(c), dev)
            blocks = [gHs, gHt, dg[:, :2]]
            if wgrad_gemm_cat_supported(blocks, x):
                with _lib.call("bgnn_adapted_transform_bwd_gates_f32"):
                    _lib.check(lib.bgnn_adapted_transform_bwd_gates_f32(_lib.ptr(gHs), _lib.ptr(gHt), _lib.ptr(gates),
                                                                        _lib.ptr(is_src, torch.uint8), _lib.ptr(wd_c), n, c,
                                                                        _lib.ptr(dg), _lib.ptr(red), _lib.ptr(ws), ws.numel(),
                                                                        _lib.stream(dev)))
                g_w = wgrad_gemm_cat(blocks, x)
                g_bias = None if bias_shape is None else red[2 * c + 2:].view(bias_shape)
                return None, g_w, g_bias, red[: 2 * c].view(wd_shape), red[2 * c: 2 * c + 2].view(kg_shape), None
        ldp = 2 * c + 4                     # 2c+2 rounded up to 4: the two GEMMs below read gP in place through TMA
        gP_buf = torch.empty((n, ldp), dtype=f32, device=dev)
        gP = gP_buf[:, : 2 * c + 2]
        red = torch.empty((4 * c + 2,), dtype=f32, device=dev)
        ws = _lib.workspace(lib.bgnn_adapted_transform_bwd_workspace_bytes(c), dev)
        with _lib.call("bgnn_adapted_transform_bwd_f32"):
            _lib.check(lib.bgnn_adapted_transform_bwd_f32(_lib.ptr(gHs), _lib.ptr(gHt), _lib.ptr(gates),
                                                          _lib.ptr(is_src, torch.uint8), _lib.ptr(wd_c), n, c, ldp,
                                                          _lib.ptr(gP_buf), _lib.ptr(red), _lib.ptr(ws), ws.numel(),
                                                          _lib.stream(dev)))
        g_x = g_w = None
        if ctx.needs_input_grad[0]:
            g_x = rowpanel_gemm(gP, w_cat.t()) if rowpanel_gemm_supported(2 * c + 2, ldp, x.shape[1]) else gP @ w_cat
        if ctx.needs_input_grad[1]:
            g_w = wgrad_gemm(gP, x) if wgrad_gemm_supported(gP, x) else gP.t() @ x
        g_bias = None if bias_shape is None else red[2 * c + 2:].view(bias_shape)
        return g_x, g_w, g_bias, red[: 2 * c].view(wd_shape), red[2 * c: 2 * c + 2].view(kg_shape), None


def adapted_wide_supported(c, d):
    if os.environ.get("BGNN_NO_WIDE"):       # A/B switch for tools/profile_mp.py
        return False
    return bool(_lib.load().bgnn_adapted_wide_supported(int(c), int(d)))


def adapted_wide(x, w_cat, bias, wd, kg, is_src):
    """AdaptedConv's node-wise transform for wide outputs (c a multiple of 32) straight from x: the contraction with
    w_cat [2c+2, d] on the tensor cores (3 x TF32) with the gates, biases and rank-1 corrections applied to the
    accumulator tile (models/KTGNN.py:275-284); returns (Hs, Ht).  bias [2c] or None.
    Differentiable in x, w_cat, bias, wd, kg."""
    return _AdaptedWideFn.apply(x, w_cat, bias, wd, kg, is_src)


# ----------------------------------------------------------------------------------- BatchNorm1d + ReLU
class _BnReluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps, relu):
        lib = _lib.load()
        f32 = torch.float32
        x = x.to(f32).contiguous()
        n, c = x.shape
        dev = x.device
        y = torch.empty_like(x)
        stats = torch.empty((4 * c,), dtype=f32, device=dev)
        ws = _lib.workspace(lib.bgnn_bn_relu_workspace_bytes(c), dev)
        w_c = None if weight is None else weight.detach().to(f32).contiguous()
        b_c = None if bias is None else bias.detach().to(f32).contiguous()
        with _lib.call("bgnn_bn_relu_fwd_f32"):
            _lib.check(lib.bgnn_bn_relu_fwd_f32(_lib.ptr(x), n, c, _lib.ptr(w_c, f32, True), _lib.ptr(b_c, f32, True), float(eps),
                                                float(momentum), _lib.ptr(running_mean, f32, True),
                                                _lib.ptr(running_var, f32, True), int(relu), _lib.ptr(y), _lib.ptr(stats),
                                                _lib.ptr(ws), ws.numel(), _lib.stream(dev)))
        ctx.save_for_backward(x, stats)
        ctx.meta = (int(relu), weight is not None, bias is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        x, stats = ctx.saved_tensors
        relu, has_w, has_b = ctx.meta
        f32 = torch.float32
        gy = gy.to(f32).contiguous()
        n, c = x.shape
        dev = x.device
        gx = torch.empty_like(x)
        gwb = torch.empty((2 * c,), dtype=f32, device=dev)
        ws = _lib.workspace(lib.bgnn_bn_relu_workspace_bytes(c), dev)
        with _lib.call("bgnn_bn_relu_bwd_f32"):
            _lib.check(lib.bgnn_bn_relu_bwd_f32(_lib.ptr(gy), _lib.ptr(x), n, c, _lib.ptr(stats), relu, _lib.ptr(gx),
                                                _lib.ptr(gwb), _lib.ptr(ws), ws.numel(), _lib.stream(dev)))
        return gx, (gwb[:c] if has_w else None), (gwb[c:] if has_b else None), None, None, None, None, None


def batch_norm_relu_supported(x, bn):
    """Whether ``batch_norm_relu`` covers this call: fp32 CUDA [n, c] input, c % 4 == 0; training mode with batch
    statistics and a fixed momentum, or inference (no autograd) with running statistics."""
    if os.environ.get("BGNN_NO_WIDE") or not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[0] > 1):
        return False
    if not bool(_lib.load().bgnn_bn_relu_supported(x.shape[1])):
        return False
    if bn.training:
        return bn.momentum is not None or not bn.track_running_stats
    return bn.track_running_stats and not (torch.is_grad_enabled() and (x.requires_grad or (bn.affine and bn.weight.requires_grad)))


def batch_norm_relu(x, bn, relu=True):
    """relu(bn(x)) for an ``nn.BatchNorm1d`` over the node-feature matrix (models/KTGNN.py:363-366, 425-429) in two
    passes over x each way (statistics; normalise + affine + ReLU), against ATen's six forward and four backward.
    Training mode updates bn.running_mean / running_var / num_batches_tracked like torch."""
    lib = _lib.load()
    f32 = torch.float32
    if bn.training:
        track = bn.track_running_stats and bn.running_mean is not None
        if track and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        return _BnReluFn.apply(x, bn.weight, bn.bias, bn.running_mean if track else None, bn.running_var if track else None,
                               bn.momentum if bn.momentum is not None else 0.0, bn.eps, relu)
    x = x.detach().to(f32).contiguous()
    n, c = x.shape
    w = bn.weight.detach() if bn.affine else torch.ones(c, dtype=f32, device=x.device)
    b = bn.bias.detach() if bn.affine else torch.zeros(c, dtype=f32, device=x.device)
    scale = w * torch.rsqrt(bn.running_var + bn.eps)
    stats = torch.cat((bn.running_mean, scale, scale, b)).to(f32).contiguous()
    y = torch.empty_like(x)
    with _lib.call("bgnn_bn_relu_apply_f32"):
        _lib.check(lib.bgnn_bn_relu_apply_f32(_lib.ptr(x), n, c, _lib.ptr(stats), int(relu), _lib.ptr(y), _lib.stream(x.device)))
    return y


class _BnReluDistFn(torch.autograd.Function):
    """BatchNorm1d (+ ReLU), training mode, over a node-feature matrix whose ROWS are partitioned over the ranks of
    ``group``: every rank reduces its own rows with the library's kernels, the per-rank (count, mean, variance) are
    combined exactly (Chan et al.'s pairwise update, in double) after one all-gather of 2c+1 numbers per rank, and the
    normalisation uses the global statistics.  Backward: per-rank column sums, one all-reduce of 2c numbers, one apply
    pass.  Replaces torch.nn.SyncBatchNorm + ReLU in the destination-partitioned KT-GNN."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps, relu, group, n_rows):
        import torch.distributed as dist
        lib = _lib.load()
        f32, f64 = torch.float32, torch.float64
        x = x.to(f32).contiguous()
        c = x.shape[1]
        n_loc = int(n_rows)                     # real rows of this rank (padding rows, if any, follow them)
        dev = x.device
        world = dist.get_world_size(group)
        stats_loc = torch.empty((4 * c,), dtype=f32, device=dev)
        ws = _lib.workspace(lib.bgnn_bn_relu_workspace_bytes(c), dev)
        with _lib.call("bgnn_bn_relu_fwd_f32", "bgnn_bn_relu_stats_f32"):
            _lib.check(lib.bgnn_bn_relu_fwd_f32(_lib.ptr(x), n_loc, c, None, None, 0.0, 0.0, None, None, 0, None,
                                                _lib.ptr(stats_loc), _lib.ptr(ws), ws.numel(), _lib.stream(dev)))
        cnt0 = torch.full((1,), float(n_loc), dtype=f64, device=dev)
        if n_loc == 0:
            mine = torch.zeros((2 * c + 1,), dtype=f64, device=dev)
        else:                                    # eps = 0 above: invstd^-2 is the biased variance
            mine = torch.cat((cnt0, stats_loc[:c].to(f64), stats_loc[c:2 * c].to(f64).pow(-2)))
        allr = torch.empty((world, 2 * c + 1), dtype=f64, device=dev)
        dist.all_gather_into_tensor(allr, mine.view(1, -1), group=group)
        cnt, mean_r, var_r = allr[:, :1], allr[:, 1:c + 1], allr[:, c + 1:]
        n_tot = cnt.sum()
        mean = (cnt * mean_r).sum(0) / n_tot
        var = (cnt * (var_r + (mean_r - mean).pow(2))).sum(0) / n_tot
        invstd = (var + eps).rsqrt()
        w = torch.ones(c, dtype=f64, device=dev) if weight is None else weight.detach().to(f64)
        b = torch.zeros(c, dtype=f64, device=dev) if bias is None else bias.detach().to(f64)
        stats = torch.cat((mean, invstd, w * invstd, b)).to(f32).contiguous()
        if running_mean is not None:
            running_mean.mul_(1.0 - momentum).add_((momentum * mean).to(running_mean.dtype))
            unbiased = var * (n_tot / (n_tot - 1).clamp(min=1.0))
            running_var.mul_(1.0 - momentum).add_((momentum * unbiased).to(running_var.dtype))
        y = torch.empty_like(x)
        with _lib.call("bgnn_bn_relu_apply_f32"):
            _lib.check(lib.bgnn_bn_relu_apply_f32(_lib.ptr(x), x.shape[0], c, _lib.ptr(stats), int(relu), _lib.ptr(y),
                                                  _lib.stream(dev)))
        ctx.save_for_backward(x, stats, n_tot)
        ctx.meta = (int(relu), weight is not None, bias is not None, group, n_loc)
        return y

    @staticmethod
    def backward(ctx, gy):
        import torch.distributed as dist
        lib = _lib.load()
        x, stats, n_tot = ctx.saved_tensors
        relu, has_w, has_b, group, n_loc = ctx.meta
        f32 = torch.float32
        gy = gy.to(f32).contiguous()
        c = x.shape[1]
        dev = x.device
        gwb = torch.empty((2 * c,), dtype=f32, device=dev)          # this rank's (d weight | d bias)
        ws = _lib.workspace(lib.bgnn_bn_relu_workspace_bytes(c), dev)
        with _lib.call("bgnn_bn_relu_bwd_reduce_f32"):
            _lib.check(lib.bgnn_bn_relu_bwd_reduce_f32(_lib.ptr(gy), _lib.ptr(x), n_loc, c, _lib.ptr(stats), relu, _lib.ptr(gwb),
                                                       _lib.ptr(ws), ws.numel(), _lib.stream(dev)))
        tot = gwb.to(torch.float64)
        dist.all_reduce(tot, group=group)
        coef = (torch.cat((tot[c:], tot[:c])) / n_tot).to(f32).contiguous()      # (mean g | mean g xhat) over all rows
        gx = torch.zeros_like(x) if n_loc < x.shape[0] else torch.empty_like(x)
        with _lib.call("bgnn_bn_relu_bwd_apply_f32"):
            _lib.check(lib.bgnn_bn_relu_bwd_apply_f32(_lib.ptr(gy), _lib.ptr(x), n_loc, c, _lib.ptr(stats), relu, _lib.ptr(coef),
                                                      _lib.ptr(gx), _lib.stream(dev)))
        # parameter gradients stay per rank: the caller sums them over ranks with every other parameter gradient
        return gx, (gwb[:c] if has_w else None), (gwb[c:] if has_b else None), None, None, None, None, None, None, None


def batch_norm_relu_dist(x, bn, group, n_rows=None, relu=True):
    """relu(bn(x)) in training mode for x row-partitioned over ``group`` (see _BnReluDistFn).  ``n_rows``: real rows of
    this rank when x carries padding rows at the end (they come out as relu(bn(0)) and get zero gradient)."""
    if not bn.training:
        return batch_norm_relu(x, bn, relu)
    track = bn.track_running_stats and bn.running_mean is not None
    if track and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return _BnReluDistFn.apply(x, bn.weight, bn.bias, bn.running_mean if track else None, bn.running_var if track else None,
                               bn.momentum if bn.momentum is not None else 0.0, bn.eps, relu, group,
                               x.shape[0] if n_rows is None else n_rows)


# ----------------------------------------------------------------------------------- narrow AdaptedConv transform
class _AdaptedSkinnyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w_cat, bias_cat, wd, kg, is_src):
        lib = _lib.load()
        f32 = torch.float32
        x, w_cat = x.to(f32).contiguous(), w_cat.to(f32).contiguous()
        wd_c, kg_c = wd.to(f32).contiguous().view(-1), kg.to(f32).contiguous().view(-1)
        b_c = None if bias_cat is None else bias_cat.to(f32).contiguous()
        n, d = x.shape
        c = (w_cat.shape[0] - 2) // 2
        dev = x.device
        Hs = torch.empty((n, c), dtype=f32, device=dev)
        Ht = torch.empty((n, c), dtype=f32, device=dev)
        gates = torch.empty((n, 2), dtype=f32, device=dev)
        with _lib.call("bgnn_adapted_skinny_fwd_f32"):
            _lib.check(lib.bgnn_adapted_skinny_fwd_f32(_lib.ptr(x), _lib.ptr(is_src, torch.uint8), _lib.ptr(w_cat),
                                                       _lib.ptr(b_c, f32, True), _lib.ptr(wd_c), _lib.ptr(kg_c), n, d, c,
                                                       _lib.ptr(Hs), _lib.ptr(Ht), _lib.ptr(gates), _lib.stream(dev)))
        ctx.save_for_backward(x, w_cat, wd_c, gates, is_src)
        ctx.meta = (wd.shape, kg.shape, c, d, bias_cat is not None)
        return Hs, Ht

    @staticmethod
    def backward(ctx, gHs, gHt):
        lib = _lib.load()
        x, w_cat, wd_c, gates, is_src = ctx.saved_tensors
        wd_shape, kg_shape, c, d, has_bias = ctx.meta
        f32 = torch.float32
        gHs, gHt = gHs.to(f32).contiguous(), gHt.to(f32).contiguous()
        n, dev, o = x.shape[0], x.device, 2 * c + 2
        gx = torch.empty_like(x)
        red = torch.empty((o * d + o + 2 * c,), dtype=f32, device=dev)
        ws = _lib.workspace(lib.bgnn_adapted_skinny_bwd_workspace_bytes(c, d), dev)
        with _lib.call("bgnn_adapted_skinny_bwd_f32"):
            _lib.check(lib.bgnn_adapted_skinny_bwd_f32(_lib.ptr(x), _lib.ptr(is_src, torch.uint8), _lib.ptr(w_cat),
                                                       _lib.ptr(wd_c), _lib.ptr(gates), _lib.ptr(gHs), _lib.ptr(gHt), n, d,
                                                       c, _lib.ptr(gx), _lib.ptr(red), _lib.ptr(ws), ws.numel(),
                                                       _lib.stream(dev)))
        g_w = red[: o * d].view(o, d)
        colsum = red[o * d: o * d + o]
        g_bias = colsum.clone() if has_bias else None
        if g_bias is not None:
            g_bias[2 * c:] = 0.0                       # the gate columns carry no bias parameter
        return gx, g_w, g_bias, red[o * d + o:].view(wd_shape), colsum[2 * c:].view(kg_shape), None


def adapted_skinny_supported(c, d):
    return bool(_lib.load().bgnn_adapted_skinny_supported(int(c), int(d)))


def adapted_skinny(x, w_cat, bias_cat, wd, kg, is_src):
    """AdaptedConv's node-wise transform for narrow outputs (c <= 4 classes) straight from x: the contraction with
    w_cat [2c+2, d] = [W_s; W_t; a_g_s2t[:d]; a_g_t2s[:d]], the gates and the rank-1 corrections in one pass over x
    (models/KTGNN.py:275-284); returns (Hs, Ht).  Differentiable in x, w_cat, bias_cat, wd, kg."""
    return _AdaptedSkinnyFn.apply(x, w_cat, bias_cat, wd, kg, is_src)


class _SkinnyGroupFn(torch.autograd.Function):
    """Domain means + narrow node-wise transform of 1 or 2 heads over the same x, as ONE autograd node: x is read
    by one pass each way (plus the column-sum pass of the means) and receives ONE gradient -- the part that flows
    through the means (Delta -> wd, kg) is added inside the backward kernel instead of by a [n, d] ``where`` and an
    accumulation pass per consumer."""

    @staticmethod
    def forward(ctx, x, is_src, inv_counts, heads, group, *params):
        lib = _lib.load()
        f32 = torch.float32
        x = x.to(f32).contiguous()
        n, d = x.shape
        dev = x.device
        ws_ = [params[3 * h].to(f32) for h in range(heads)]
        bs_ = [params[3 * h + 1] for h in range(heads)]
        as_ = [params[3 * h + 2].to(f32) for h in range(heads)]
        o = ws_[0].shape[0]
        c = (o - 2) // 2
        sums = torch.empty((2, d), dtype=f32, device=dev)
        wsp = _lib.workspace(lib.bgnn_domain_colsum_workspace_bytes(d), dev)
        with _lib.call("bgnn_domain_colsum_f32"):
            _lib.check(lib.bgnn_domain_colsum_f32(_lib.ptr(x), _lib.ptr(is_src, torch.uint8), n, d, _lib.ptr(sums),
                                                  _lib.ptr(wsp), wsp.numel(), _lib.stream(dev)))
        if group is not None:      # rows partitioned over ranks: the domain sums are global
            import torch.distributed as dist
            dist.all_reduce(sums, group=group)
        means = sums * inv_counts.view(2, 1)
        delta = means[0] - means[1]                                        # [d]
        wcat = torch.cat(ws_, 0).contiguous()                              # [heads*o, d]
        atail = torch.cat(as_, 0).contiguous()                             # [heads*2, d]
        has_bias = [b is not None for b in bs_]
        bias = None
        if any(has_bias):
            bias = torch.cat([b.to(f32) if b is not None else torch.zeros(o, dtype=f32, device=dev) for b in bs_]).contiguous()
        wd = (wcat.view(heads, o, d)[:, : 2 * c, :] @ delta).reshape(-1).contiguous()      # [heads*2c]
        kg = (atail @ delta).contiguous()                                  # [heads*2]
        Hs = torch.empty((n, heads * c), dtype=f32, device=dev)
        Ht = torch.empty((n, heads * c), dtype=f32, device=dev)
        gates = torch.empty((n, heads * 2), dtype=f32, device=dev)
        if not os.environ.get("BGNN_NO_SKINNY_TC") and lib.bgnn_adapted_skinny_tc_supported(c, d, heads):
            # the heads * (2c+2) dot products per row on the tensor cores (3 x TF32), gates applied to the accumulator
            hi, lo = tf32_planes(wcat)
            with _lib.call("bgnn_adapted_skinny_heads_tc_fwd_f32"):
                _lib.check(lib.bgnn_adapted_skinny_heads_tc_fwd_f32(_lib.ptr(x), n, d, _lib.ptr(hi), _lib.ptr(lo), c, heads,
                                                                    _lib.ptr(is_src, torch.uint8), _lib.ptr(wd), _lib.ptr(kg),
                                                                    _lib.ptr(bias, f32, True), _lib.ptr(Hs), _lib.ptr(Ht),
                                                                    _lib.ptr(gates), _lib.stream(dev)))
        else:
            with _lib.call("bgnn_adapted_skinny_heads_fwd_f32"):
                _lib.check(lib.bgnn_adapted_skinny_heads_fwd_f32(_lib.ptr(x), _lib.ptr(is_src, torch.uint8), _lib.ptr(wcat),
                                                                 _lib.ptr(bias, f32, True), _lib.ptr(wd), _lib.ptr(kg), n, d,
                                                                 c, heads, _lib.ptr(Hs), _lib.ptr(Ht), _lib.ptr(gates),
                                                                 _lib.stream(dev)))
        ctx.save_for_backward(x, wcat, atail, wd, gates, is_src, delta, inv_counts)
        ctx.meta = (heads, c, d, has_bias, group)
        return Hs, Ht

    @staticmethod
    def backward(ctx, gHs, gHt):
        lib = _lib.load()
        x, wcat, atail, wd, gates, is_src, delta, inv_counts = ctx.saved_tensors
        heads, c, d, has_bias, group = ctx.meta
        f32 = torch.float32
        gHs, gHt = gHs.to(f32).contiguous(), gHt.to(f32).contiguous()
        n, dev, o = x.shape[0], x.device, 2 * c + 2
        ws = _lib.workspace(lib.bgnn_adapted_skinny_heads_bwd_workspace_bytes(c, d, heads), dev)
        pre = torch.empty((heads * o,), dtype=f32, device=dev)
        with _lib.call("bgnn_adapted_skinny_heads_pre_f32"):
            _lib.check(lib.bgnn_adapted_skinny_heads_pre_f32(_lib.ptr(is_src, torch.uint8), _lib.ptr(wd), _lib.ptr(gates),
                                                             _lib.ptr(gHs), _lib.ptr(gHt), n, c, heads, _lib.ptr(pre),
                                                             _lib.ptr(ws), ws.numel(), _lib.stream(dev)))
        pre = pre.view(heads, o)
        g_wd, g_kg = pre[:, : 2 * c], pre[:, 2 * c:]                        # [heads, 2c], [heads, 2]
        w3, a3 = wcat.view(heads, o, d), atail.view(heads, 2, d)
        # wd_h = W_h[:2c] delta, kg_h = A_h delta  ->  d delta, and through the two means into every row of x
        g_delta = torch.einsum("hj,hjd->d", g_wd, w3[:, : 2 * c, :]) + torch.einsum("hk,hkd->d", g_kg, a3)
        if group is not None:      # Delta is a function of EVERY rank's rows: its gradient is the sum over ranks
            import torch.distributed as dist
            g_delta = g_delta.contiguous()
            dist.all_reduce(g_delta, group=group)
        gm = torch.stack((g_delta * inv_counts[0], -g_delta * inv_counts[1])).contiguous()
        gx = torch.empty_like(x)
        red = torch.empty((heads * (o * d + o + 2 * c),), dtype=f32, device=dev)
        with _lib.call("bgnn_adapted_skinny_heads_bwd_f32"):
            _lib.check(lib.bgnn_adapted_skinny_heads_bwd_f32(_lib.ptr(x), _lib.ptr(is_src, torch.uint8), _lib.ptr(wcat),
                                                             _lib.ptr(wd), _lib.ptr(gates), _lib.ptr(gHs), _lib.ptr(gHt),
                                                             _lib.ptr(gm), n, d, c, heads, _lib.ptr(gx), _lib.ptr(red),
                                                             _lib.ptr(ws), ws.numel(), _lib.stream(dev)))
        g_w = red[: heads * o * d].view(heads, o, d).clone()
        g_w[:, : 2 * c, :] += g_wd.unsqueeze(2) * delta.view(1, 1, d)       # wd = W[:2c] delta
        colsum = red[heads * o * d: heads * o * d + heads * o].view(heads, o)
        g_a = g_kg.unsqueeze(2) * delta.view(1, 1, d)                       # [heads, 2, d]
        out = [gx, None, None, None, None]
        for h in range(heads):
            g_b = None
            if has_bias[h]:
                g_b = colsum[h].clone()
                g_b[2 * c:] = 0.0                                           # the gate columns carry no bias parameter
            out += [g_w[h], g_b, g_a[h]]
        return tuple(out)


def adapted_skinny_group_supported(x, c, heads):
    return (x.is_cuda and x.dtype == torch.float32 and not os.environ.get("BGNN_NO_WIDE")
            and bool(_lib.load().bgnn_adapted_skinny_heads_supported(int(c), int(x.shape[1]), int(heads)))
            and domain_colsum_supported(x.shape[1]))


def adapted_skinny_group(x, is_src, inv_counts, head_params, group=None):
    """Node-wise part (models/KTGNN.py:275-284) of 1 or 2 narrow AdaptedConvs over the SAME input x, domain means
    included.  ``head_params``: per head (w_cat [2c+2, d] = [W_s; W_t; a_g_s2t[:d]; a_g_t2s[:d]], b_cat [2c+2] or
    None, a_tail [2, d] = [a_g_s2t[d:]; a_g_t2s[d:]]); ``inv_counts`` = (1/Ns, 1/Nt).  Returns (Hs, Ht), each
    [n, heads*c] with the heads side by side.  Differentiable in x and every parameter.  ``group``: x holds this
    rank's rows of a row-partitioned matrix (inv_counts are then the GLOBAL domain sizes); the parameter gradients
    returned are this rank's share."""
    flat = []
    for w, b, a in head_params:
        flat += [w, b, a]
    return _SkinnyGroupFn.apply(x, is_src, inv_counts, len(head_params), group, *flat)


class _DomainMeansFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, is_src, inv_counts):
        lib = _lib.load()
        x = x.to(torch.float32).contiguous()
        n, d = x.shape
        sums = torch.empty((2, d), dtype=torch.float32, device=x.device)
        ws = _lib.workspace(lib.bgnn_domain_colsum_workspace_bytes(d), x.device)
        with _lib.call("bgnn_domain_colsum_f32"):
            _lib.check(lib.bgnn_domain_colsum_f32(_lib.ptr(x), _lib.ptr(is_src, torch.uint8), n, d, _lib.ptr(sums),
                                                  _lib.ptr(ws), ws.numel(), _lib.stream(x.device)))
        ctx.save_for_backward(inv_counts, is_src)
        return sums * inv_counts.view(2, 1)

    @staticmethod
    def backward(ctx, gmeans):
        inv_counts, is_src = ctx.saved_tensors
        g = gmeans * inv_counts.view(2, 1)
        # every source row receives g[0], every target row g[1]: one elementwise pass, no gather
        return torch.where(is_src.view(-1, 1) != 0, g[0:1], g[1:2]), None, None


def domain_colsum_supported(d):
    return d >= 4 and d % 4 == 0 and d <= 1024


def domain_means(x, is_src, inv_counts):
    """[2, d]: mean of the source-domain rows and of the target-domain rows of x (models/KTGNN.py:275-276), one
    pass over x.  is_src uint8 [n]; inv_counts float32 [2] = (1/Ns, 1/Nt)."""
    return _DomainMeansFn.apply(x, is_src, inv_counts)

