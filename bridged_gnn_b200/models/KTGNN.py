"""KT-GNN layers and models with the reference's constructor / forward / state_dict surface
(models/KTGNN.py:218-328 AdaptedConv, :330-465 KTGNN_no_complement, :467-597 KTGNN_noDTC), running the
edge part of every conv as one fused sm_100a kernel (ops.gat_aggregate, ops.gat_aggregate_heads) instead of PyG's
gather -> softmax -> propagate chain, and the node-wise dense part (AdaptedConv's contraction with its gates,
clf_transformer's Linear layers, BatchNorm1d + ReLU) on this library's tensor-core / two-pass kernels when the input
is an fp32 CUDA matrix of a supported width (torch's own modules otherwise: CPU tensors, SyncBatchNorm, odd widths).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..data import add_self_loops, remove_self_loops


def bn_relu(bn, x, part=None):
    """relu(bn(x)) (models/KTGNN.py:425-429); plain nn.BatchNorm1d on the GPU goes through the fused two-pass kernels
    -- with the batch statistics combined over the ranks of ``part`` (a ``dist.DstPartition``) when x is row-partitioned
    -- anything else (SyncBatchNorm, CPU, other dtypes) through the module itself."""
    if type(bn) is nn.BatchNorm1d and ops.batch_norm_relu_supported(x, bn):
        if part is not None and part.world > 1 and bn.training:
            return ops.batch_norm_relu_dist(x, bn, part.group, part.r1 - part.r0)
        return ops.batch_norm_relu(x, bn)
    if part is not None and part.world > 1 and bn.training and type(bn) is nn.BatchNorm1d:
        raise RuntimeError("row-partitioned BatchNorm1d on this input needs torch.nn.SyncBatchNorm "
                           "(convert_sync_batchnorm) or an fp32 CUDA input of a supported width")
    return F.relu(bn(x))


def run_sequential(seq, x, part=None):
    """``seq(x)`` for an nn.Sequential, with every BatchNorm1d -> ReLU pair taken by ``bn_relu``
    (clf_transformer, models/KTGNN.py:363-366)."""
    mods = list(seq)
    i = 0
    while i < len(mods):
        if isinstance(mods[i], nn.modules.batchnorm._BatchNorm) and i + 1 < len(mods) and type(mods[i + 1]) is nn.ReLU:
            x = bn_relu(mods[i], x, part)
            i += 2
        else:
            x = mods[i](x)
            i += 1
    return x


class NodeLinear(nn.Linear):
    """nn.Linear over the node-feature matrix (same parameters and state_dict keys); on the GPU in fp32 the forward,
    input-gradient and weight-gradient GEMMs run on this library's 3 x TF32 tensor-core kernels instead of cuBLAS'
    fp32 SIMT GEMMs (the Linear layers of clf_transformer, models/KTGNN.py:363)."""

    def forward(self, x):
        if ops.linear_supported(x, self.weight):
            return ops.linear(x, self.weight, self.bias)
        return F.linear(x, self.weight, self.bias)


class AdaptedConv(nn.Module):
    """Domain-adaptive attention conv.  Parameters and their names match the reference
    (lin_s, lin_t, a_g_s2t, a_g_t2s, a_f_s2t, a_f_t2s[, lin_r]) so its state_dicts load unchanged."""

    def __init__(self, in_channels, out_channels, normalize=False, root_weight=True, activation_g=None,
                 negative_slope=0.1, bias=True, **kwargs):
        super().__init__()
        if kwargs.get("aggr", "add") != "add":
            raise NotImplementedError("AdaptedConv aggregates with 'add' (models/KTGNN.py:223)")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.normalize, self.root_weight = normalize, root_weight
        self.activation_g, self.negative_slope = activation_g, negative_slope
        cin = in_channels if isinstance(in_channels, int) else in_channels[0]
        cin_r = in_channels if isinstance(in_channels, int) else in_channels[1]
        if root_weight:
            self.lin_r = nn.Linear(cin_r, out_channels, bias=False)
        self.lin_s = nn.Linear(cin, out_channels, bias=bias)
        self.lin_t = nn.Linear(cin, out_channels, bias=bias)
        self.a_g_s2t = nn.Linear(cin * 2, 1, bias=False)
        self.a_g_t2s = nn.Linear(cin * 2, 1, bias=False)
        self.a_f_s2t = nn.Linear(out_channels, 1, bias=False)
        self.a_f_t2s = nn.Linear(out_channels, 1, bias=False)
        self._mask_key, self._mask_u8 = None, None
        self._rows_key, self._rows = None, None

    def reset_parameters(self):
        for m in self.children():
            m.reset_parameters()

    # The caches below are keyed on the mask TENSOR (held in the key, compared by identity) and its version: an
    # address alone could be recycled by the allocator for a new mask of the same length.
    @staticmethod
    def _same(key, mask, *extra):
        return key is not None and key[0] is mask and key[1] == mask._version and key[2:] == extra

    def _domain_rows(self, central_mask, dtype):
        key = (central_mask, central_mask._version, dtype)
        if not self._same(self._rows_key, central_mask, dtype):
            cf = central_mask.to(dtype)
            ns = cf.sum().clamp(min=1.0)
            nt = (central_mask.shape[0] - cf.sum()).clamp(min=1.0)
            self._rows_key, self._rows = key, torch.stack((cf / ns, (1.0 - cf) / nt), 0).contiguous()
        return self._rows

    def _domain_counts(self, central_mask):
        key = (central_mask, central_mask._version)
        if not self._same(getattr(self, "_counts_key", None), central_mask):
            ns = central_mask.sum().clamp(min=1).to(torch.float32)
            nt = (central_mask.shape[0] - central_mask.sum()).clamp(min=1).to(torch.float32)
            self._counts_key = key
            self._counts = torch.stack((1.0 / ns, 1.0 / nt))
        return self._counts

    def _padded_params(self):
        """lin_s / lin_t / a_f with the output width padded to a multiple of 4 by zero rows when that lets the
        gather kernels use 128-bit loads (e.g. 31 classes -> 32).  The zero columns of H contribute nothing to
        the attention scores (their a_f entries are 0) and are sliced off the output."""
        co = self.out_channels
        cp = (co + 3) // 4 * 4 if (co % 4 != 0 and co > 8) else co
        w_s, w_t, b_s, b_t = self.lin_s.weight, self.lin_t.weight, self.lin_s.bias, self.lin_t.bias
        a1, a2 = self.a_f_t2s.weight, self.a_f_s2t.weight
        if cp != co:
            zw = w_s.new_zeros((cp - co, w_s.shape[1]))
            w_s, w_t = torch.cat((w_s, zw), 0), torch.cat((w_t, zw), 0)
            if b_s is not None:
                zb = b_s.new_zeros(cp - co)
                b_s, b_t = torch.cat((b_s, zb)), torch.cat((b_t, zb))
            za = a1.new_zeros((1, cp - co))
            a1, a2 = torch.cat((a1, za), 1), torch.cat((a2, za), 1)
        return w_s, w_t, b_s, b_t, a1, a2, cp

    def _dst_is_src(self, central_mask):
        key = (central_mask, central_mask._version)
        if not self._same(self._mask_key, central_mask):
            self._mask_key, self._mask_u8 = key, central_mask.to(torch.uint8).contiguous()
        return self._mask_u8

    def forward(self, x, edge_index, edge_index1=None, edge_index2=None, central_mask=None, size=None, part=None):
        """edge_index must be cat(edge_index1, edge_index2) where edge_index1 / edge_index2 hold the edges
        whose destination is a source- / target-domain node (KTGNN.graph_partition).  The fused kernel
        picks the branch per destination row from ``central_mask``, so only ``edge_index`` is read.

        ``part`` (a ``dist.DstPartition``) switches to the multi-GPU layout: ``x`` holds this rank's rows only,
        ``edge_index`` this rank's incoming edges (global ids), ``central_mask`` the padded global mask."""
        if part is not None:
            return self._forward_partitioned(x, edge_index, central_mask, part)
        x_src, x_r = (x, x) if torch.is_tensor(x) else x
        c = central_mask
        h_s, h_t, a_t2s, a_s2t, cp = self.node_part(x_src, c)
        # attention scores, softmax over destinations, weighted aggregation (:292-305) -- one kernel
        graph = ops.cached_graph(edge_index, x_src.shape[0])
        out = ops.gat_aggregate(h_s, h_t, a_t2s, a_s2t, graph, self._dst_is_src(c), self.negative_slope)
        return self._finish(out, cp, x_r)

    def _finish(self, out, cp, x_r):
        co = self.out_channels
        if cp != co:
            out = out[:, :co]
        if self.root_weight and x_r is not None:
            out = out + self.lin_r(x_r)
        if self.normalize:
            out = F.normalize(out, p=2.0, dim=-1)
        return out

    def domain_means(self, x_src, c):
        """[2, D]: mean of the source-domain rows and of the target-domain rows of x (models/KTGNN.py:275-276)."""
        if x_src.is_cuda and x_src.dtype == torch.float32 and ops.domain_colsum_supported(x_src.shape[1]):
            return ops.domain_means(x_src, self._dst_is_src(c), self._domain_counts(c))    # one pass over x
        return self._domain_rows(c, x_src.dtype) @ x_src             # [2, N]: 1/Ns on source rows, 1/Nt on target rows

    def node_part(self, x_src, c, means=None):
        """The node-wise half of the conv (models/KTGNN.py:275-284): returns (lin_s(x_t2s), lin_t(x_s2t), a_f_t2s,
        a_f_s2t, padded width) -- everything the edge part needs.  ``means`` ([2, D] domain means of x_src) may be
        passed in when another conv over the same input has computed them already."""
        n, d = x_src.shape
        # g + f (models/KTGNN.py:275-284) restructured so that x is read by ONE dense contraction and no
        # [N, 2D] / [N, D] intermediate is materialised.  With Delta = mean_src(x) - mean_tar(x):
        #   gate_s2t = tanh(a_g_s2t . [x, Delta]),  gate_t2s = tanh(a_g_t2s . [x, Delta])
        #   h_t = lin_t(x - gate_s2t * Delta * c)      = x W_t^T + b_t - (gate_s2t * c)     (x) (W_t Delta)
        #   h_s = lin_s(x + gate_t2s * Delta * (1-c))  = x W_s^T + b_s + (gate_t2s * (1-c)) (x) (W_s Delta)
        on_gpu = x_src.is_cuda and x_src.dtype == torch.float32
        if means is None:
            means = self.domain_means(x_src, c)
        delta = means[0:1] - means[1:2]                              # [1, D]
        w_s, w_t, b_s, b_t, a_t2s, a_s2t, cp = self._padded_params()
        w_cat = torch.cat((w_s, w_t, self.a_g_s2t.weight[:, :d], self.a_g_t2s.weight[:, :d]), 0)
        b_cat = None if b_s is None else torch.cat((b_s, b_t, b_s.new_zeros(2)))
        k_g = torch.stack(((self.a_g_s2t.weight[:, d:] * delta).sum(), (self.a_g_t2s.weight[:, d:] * delta).sum()))
        wd = delta @ torch.cat((w_s, w_t), 0).t()                    # [1, 2*cp]: W_s Delta, W_t Delta
        if on_gpu and ops.adapted_skinny_supported(cp, d):
            # classifier heads (a few classes): contraction, gates and corrections in one pass over x
            h_s, h_t = ops.adapted_skinny(x_src, w_cat, b_cat, wd, k_g, self._dst_is_src(c))
        elif on_gpu and ops.adapted_wide_supported(cp, d):
            # hidden layers: the contraction on the tensor cores (3 x TF32) with the node-wise epilogue applied to
            # the accumulator tile -- x is read once, P is never written
            b2 = None if b_s is None else torch.cat((b_s, b_t))
            h_s, h_t = ops.adapted_wide(x_src, w_cat, b2, wd, k_g, self._dst_is_src(c))
        else:
            p = x_src @ w_cat.t()                                    # [N, 2*cp + 2]
            b2 = None if b_s is None else torch.cat((b_s, b_t))
            h_s, h_t = ops.adapted_transform(p, wd, k_g, self._dst_is_src(c), b2)   # biases, gates, rank-1 corrections
        return h_s, h_t, a_t2s, a_s2t, cp

    def _local_masks(self, central_mask, part):
        """(is_src of this rank's rows uint8 [n_loc], global is_src uint8 [n_pad], (1/Ns, 1/Nt)), cached per mask."""
        def build(cm):
            glob = part.pad_rows(self._dst_is_src(cm)[: part.n])
            return part.local_rows(glob[: part.n]).contiguous(), glob.contiguous(), self._domain_counts(cm[: part.n])
        return part.local_cached(("masks", id(self)), central_mask, build)

    def node_part_partitioned(self, x, central_mask, part):
        """Node-wise half for this rank's rows of a row-partitioned x: the two domain means are all-reduced, the rest
        is the single-GPU code on the local rows (padding rows count as target-domain rows; nothing reads them)."""
        from .. import dist as bdist
        d = x.shape[1]
        is_src_loc, _, inv_counts = self._local_masks(central_mask, part)
        if x.is_cuda and x.dtype == torch.float32 and ops.domain_colsum_supported(d):
            local = ops.domain_means(x, is_src_loc, inv_counts)       # padding rows of x are zero
        else:
            cf = is_src_loc.to(x.dtype)
            valid = part.local_rows(torch.ones(part.n, dtype=x.dtype, device=x.device))
            local = torch.stack((cf * inv_counts[0], (valid - cf) * inv_counts[1]), 0) @ x
        means = bdist.all_reduce_sum_autograd(local, part.group)
        delta = means[0:1] - means[1:2]
        w_s, w_t, b_s, b_t, a_t2s, a_s2t, cp = self._padded_params()
        w_cat = torch.cat((w_s, w_t, self.a_g_s2t.weight[:, :d], self.a_g_t2s.weight[:, :d]), 0)
        k_g = torch.stack(((self.a_g_s2t.weight[:, d:] * delta).sum(), (self.a_g_t2s.weight[:, d:] * delta).sum()))
        wd = delta @ torch.cat((w_s, w_t), 0).t()
        b2 = None if b_s is None else torch.cat((b_s, b_t))
        on_gpu = x.is_cuda and x.dtype == torch.float32
        if on_gpu and ops.adapted_skinny_supported(cp, d):
            b_cat = None if b_s is None else torch.cat((b_s, b_t, b_s.new_zeros(2)))
            h_s, h_t = ops.adapted_skinny(x, w_cat, b_cat, wd, k_g, is_src_loc)
        elif on_gpu and ops.adapted_wide_supported(cp, d):
            h_s, h_t = ops.adapted_wide(x, w_cat, b2, wd, k_g, is_src_loc)
        else:
            h_s, h_t = ops.adapted_transform(x @ w_cat.t(), wd, k_g, is_src_loc, b2)
        return h_s, h_t, a_t2s, a_s2t, cp

    def _forward_partitioned(self, x, edge_index, central_mask, part):
        """Destination-partitioned forward: node-wise transforms on the local rows, domain-aware halo
        exchange of H (a rank receives lin_s(.) of all nodes only if it owns source-domain destinations, lin_t(.) only
        if it owns target-domain ones), fused aggregation over the LOCAL destination rows; autograd turns the exchange
        into the transposed one for dH."""
        from .. import dist as bdist
        if self.root_weight or self.normalize:
            raise NotImplementedError("partitioned AdaptedConv supports root_weight=False, normalize=False (all recipes)")
        h_s, h_t, a_t2s, a_s2t, cp = self.node_part_partitioned(x, central_mask, part)
        H_s, H_t = bdist.halo_exchange(h_s, h_t, part, central_mask)            # [n_pad, cp] each, or None
        _, is_src_glob, _ = self._local_masks(central_mask, part)
        out = ops.gat_aggregate(H_s, H_t, a_t2s, a_s2t, part.graph(edge_index), is_src_glob, self.negative_slope)
        return out[:, : self.out_channels] if cp != self.out_channels else out

    def __repr__(self):
        return "{}({}, {})".format(self.__class__.__name__, self.in_channels, self.out_channels)


def adapted_convs_shared_graph(convs, xs, edge_index, edge_index1, edge_index2, central_mask, part=None):
    """[conv(x, ...) for conv, x in zip(convs, xs)] for AdaptedConvs over the SAME graph.  Narrow convs of equal
    width (the classifier heads, models/KTGNN.py:432-434) share one aggregation pass: their node-wise parts run
    one by one, the (lin_s, lin_t) outputs are laid side by side and a single multi-head kernel walks the edges."""
    c = central_mask
    widths = {conv.out_channels for conv in convs}
    ok = (len(widths) == 1 and len(convs) in (2, 3) and all(torch.is_tensor(x) and x.is_cuda for x in xs)
          and not any(conv.root_weight or conv.normalize for conv in convs)
          and len({conv.negative_slope for conv in convs}) == 1)
    order = list(range(len(convs)))
    if ok:
        cps = {conv._padded_params()[6] for conv in convs}
        ok = len(cps) == 1 and ops.gat_heads_supported(len(convs), next(iter(cps)))
    if ok:
        cp = next(iter(cps))
        # convs that read the same tensor form a group: one pass over x for the whole group (means included)
        groups = {}
        for i, x in enumerate(xs):
            groups.setdefault(id(x), []).append(i)
        order, hs, ht, af1, af2 = [], [], [], [], []
        for idxs in groups.values():
            x = xs[idxs[0]]
            for j in range(0, len(idxs), 2):
                sub = idxs[j: j + 2]
                if ops.adapted_skinny_group_supported(x, cp, len(sub)):
                    d = x.shape[1]
                    if part is not None:
                        is_src_rows, _, inv_counts = convs[sub[0]]._local_masks(c, part)
                    else:
                        is_src_rows, inv_counts = convs[sub[0]]._dst_is_src(c), convs[sub[0]]._domain_counts(c)
                    hp = []
                    for i in sub:
                        w_s, w_t, b_s, b_t, _, _, _ = convs[i]._padded_params()
                        a1, a2 = convs[i].a_g_s2t.weight, convs[i].a_g_t2s.weight
                        hp.append((torch.cat((w_s, w_t, a1[:, :d], a2[:, :d]), 0),
                                   None if b_s is None else torch.cat((b_s, b_t, b_s.new_zeros(2))),
                                   torch.cat((a1[:, d:], a2[:, d:]), 0)))
                    h_s, h_t = ops.adapted_skinny_group(x, is_src_rows, inv_counts, hp,
                                                        group=None if part is None or part.world == 1 else part.group)
                    hs.append(h_s)
                    ht.append(h_t)
                else:
                    means = None if part is not None else convs[sub[0]].domain_means(x, c)
                    for i in sub:
                        p = convs[i].node_part_partitioned(x, c, part) if part is not None else convs[i].node_part(x, c, means)
                        hs.append(p[0])
                        ht.append(p[1])
                for i in sub:
                    _, _, _, _, a_t2s, a_s2t, _ = convs[i]._padded_params()
                    af1.append(a_t2s.reshape(-1))
                    af2.append(a_s2t.reshape(-1))
                    order.append(i)
    if not ok:
        return [conv(x, edge_index, edge_index1, edge_index2, c, part=part) for conv, x in zip(convs, xs)]
    hs_all, ht_all = torch.cat(hs, 1), torch.cat(ht, 1)
    if part is not None:
        from .. import dist as bdist
        hs_all, ht_all = bdist.halo_exchange(hs_all, ht_all, part, c)      # every head's operands in ONE exchange
        graph, is_src = part.graph(edge_index), convs[0]._local_masks(c, part)[1]
    else:
        graph, is_src = ops.cached_graph(edge_index, xs[0].shape[0]), convs[0]._dst_is_src(c)
    out = ops.gat_aggregate_heads(hs_all, ht_all, torch.cat(af1), torch.cat(af2), graph, is_src, convs[0].negative_slope,
                                  len(convs))
    co = convs[0].out_channels
    res = [None] * len(convs)
    for pos, i in enumerate(order):                 # heads were laid out group by group
        res[i] = out[:, pos * cp: pos * cp + co]
    return res


def graph_partition(edge_index, central_mask, add_self_loop=True):
    """models/KTGNN.py:385-398: rewrite self loops, then split edges by the domain of their destination."""
    if add_self_loop:
        edge_index = add_self_loops(remove_self_loops(edge_index), central_mask.shape[0])
    m1 = central_mask[edge_index[1]]
    e1, e2 = edge_index[:, m1], edge_index[:, ~m1]
    return e1, e2, torch.cat((e1, e2), dim=-1)


class _KTGNNBase(nn.Module):
    def __init__(self, cached_edges, dropout, use_bn, need_complement):
        super().__init__()
        if need_complement:
            # Adapted_complementor (models/KTGNN.py:138) is disabled in every shipped recipe
            # (main_graph_knowledge_transfer.py:179, 332-333) and is outside the accelerated path.
            raise NotImplementedError("need_complement=True is not part of the accelerated hot path")
        self.cached_edges, self.dropout, self.use_bn, self.need_complement = cached_edges, dropout, use_bn, False
        self.convs, self.bns = nn.ModuleList(), nn.ModuleList()
        self._ei = None          # (edge_index1, edge_index2, edge_index) of graph_partition, cached (cached_edges)
        self._fast = None        # (CSRGraph.prepared, data.edge_index, data.central_mask): the graph the kernels use

    graph_partition = staticmethod(graph_partition)

    # The reference caches graph_partition's three edge lists as attributes (models/KTGNN.py:385-398).  On the GPU the
    # aggregation never reads them -- the CSR comes straight from data.edge_index (``_graph``) -- so they are
    # materialised on first access only; assigning None to ``edge_index`` announces a new graph, as in the reference.
    def _materialize(self):
        if self._ei is None and self._fast is not None:
            graph, ei_in, cm = self._fast
            self._ei = graph_partition(ei_in, cm)
            ops.register_graph(self._ei[2], cm.shape[0], graph)     # cached_graph(self.edge_index, n) is this graph
        return self._ei

    @property
    def edge_index(self):
        t = self._materialize()
        return None if t is None else t[2]

    @edge_index.setter
    def edge_index(self, value):
        if value is not None:
            raise AttributeError("edge_index is derived from data.edge_index; assign None to drop the cached graph")
        self._ei = self._fast = None

    @property
    def edge_index1(self):
        t = self._materialize()
        return None if t is None else t[0]

    @property
    def edge_index2(self):
        t = self._materialize()
        return None if t is None else t[1]

    def _edges(self, data):
        if not self.cached_edges:
            return graph_partition(data.edge_index, data.central_mask)
        if self._ei is None and self._fast is None:
            self._ei = graph_partition(data.edge_index, data.central_mask)
        return self._materialize()

    def _graph(self, data):
        """(edge_index1, edge_index2, graph) for the convs.  CUDA + cached_edges: ``graph`` is a ``CSRGraph.prepared`` built
        by ONE library call from ``data.edge_index`` (self-loop rewrite inside the key construction, transposed CSR by
        a stable sort on the source bits, slot map, row orders) and the two partitioned lists stay None -- the kernels
        pick the branch per destination row from ``central_mask``.  Otherwise: graph_partition's tensors."""
        ei_in = data.edge_index
        if not (self.cached_edges and torch.is_tensor(ei_in) and ei_in.is_cuda and ei_in.dtype == torch.int64):
            return self._edges(data)
        training = self.training and torch.is_grad_enabled()
        if self._fast is None or (training and self._fast[0]._t is None):
            if self._ei is not None:        # the partition was materialised before (CPU-style use): keep using it
                return self._ei
            g = ops.CSRGraph.prepared(ei_in, data.central_mask.shape[0], rewrite_self_loops=True, training=training)
            self._fast = (g, ei_in, data.central_mask)
        return None, None, self._fast[0]

    def prepare_graph(self, data):
        """Everything of a forward / backward pass that depends on the graph alone (models/KTGNN.py:385-398
        graph_partition's self-loop rewrite, then the CSR, transposed CSR and row orders of the aggregation kernels),
        cached for the calls that follow.  ``data`` needs ``edge_index`` and ``central_mask`` only: a serving loop can
        call this while ``data.x`` is still being copied to the device on another stream."""
        _, _, ei = self._graph(data)
        ops.cached_graph(ei, data.central_mask.shape[0]).prepare(training=self.training and torch.is_grad_enabled())
        for conv in list(self.convs) + [m for m in self.children() if isinstance(m, AdaptedConv)]:
            conv._dst_is_src(data.central_mask)
            conv._domain_counts(data.central_mask)
        return self

    def _hidden(self, x, ei, ei1, ei2, c, n_convs, part=None):
        if part is not None and self.use_bn and self.training and part.n_pad != part.n:
            raise NotImplementedError("partitioned training with BatchNorm needs num_nodes % world_size == 0 "
                                      "(padding rows would enter the batch statistics)")
        for ind in range(n_convs):
            x = self.convs[ind](x, ei, ei1, ei2, c, part=part)
            x = bn_relu(self.bns[ind], x, part) if self.use_bn else F.relu(x)
            x = F.dropout(x, p=self.dropout, training=self.training)
        return x


class KTGNN_no_complement(_KTGNNBase):
    def __init__(self, num_features, num_classes=2, layer_num=2, hidden=64, root_weight=False, dim_share=300, step=1,
                 hidden_o=128, hidden_u=128, use_dist_loss=False, cached_edges=True, dropout=0.5, use_bn=False,
                 need_complement=False):
        super().__init__(cached_edges, dropout, use_bn, need_complement)
        dim_in = dim_share
        if layer_num == 1:
            self.convs.append(AdaptedConv(dim_in, num_classes, root_weight=root_weight))
        else:
            for num in range(layer_num - 1):
                self.convs.append(AdaptedConv(dim_in if num == 0 else hidden, hidden, root_weight=root_weight))
                if use_bn:
                    self.bns.append(nn.BatchNorm1d(hidden))
        self.clf_base = AdaptedConv(hidden, num_classes, root_weight=root_weight)
        self.clf_target = AdaptedConv(hidden, num_classes, root_weight=root_weight)
        self.clf_transformer = nn.Sequential(NodeLinear(hidden, hidden), nn.BatchNorm1d(hidden), nn.ReLU(),
                                             NodeLinear(hidden, hidden))

    def get_emb(self, data):
        ei1, ei2, ei = self._graph(data)
        return self._hidden(data.x, ei, ei1, ei2, data.central_mask, len(self.convs))

    def forward(self, data):
        """``data.part`` (a ``dist.DstPartition``), if present, selects the destination-partitioned multi-GPU
        forward: ``data.x`` = this rank's rows [n_loc, F], ``data.edge_index`` = this rank's incoming edges
        AFTER graph_partition's self-loop rewrite (global ids), ``data.central_mask`` = padded global mask.
        BatchNorm1d layers combine their batch statistics over the ranks (``ops.batch_norm_relu_dist``); call
        ``part.sync_grads(model)`` after ``backward()``."""
        part = getattr(data, "part", None)
        if part is not None:
            ei1 = ei2 = None
            ei = data.edge_index
        else:
            ei1, ei2, ei = self._graph(data)
        c = data.central_mask
        x = self._hidden(data.x, ei, ei1, ei2, c, len(self.convs), part)
        logits_base, logits_trans, logits_target = adapted_convs_shared_graph(
            (self.clf_base, self.clf_target, self.clf_target), (x, run_sequential(self.clf_transformer, x, part), x), ei, ei1, ei2, c, part)
        return F.log_softmax(logits_base, 1), F.log_softmax(logits_target, 1), F.log_softmax(logits_trans, 1), None


class KTGNN_noDTC(_KTGNNBase):
    def __init__(self, num_features, num_classes=2, layer_num=2, hidden=64, root_weight=False, dim_share=300, step=1,
                 hidden_o=128, hidden_u=128, use_dist_loss=False, cached_edges=True, dropout=0.5, use_bn=False,
                 need_complement=False):
        super().__init__(cached_edges, dropout, use_bn, need_complement)
        dim_in = dim_share
        if layer_num == 1:
            self.convs.append(AdaptedConv(dim_in, num_classes, root_weight=root_weight))
        else:
            # the reference's loop runs range(layer_num-1), so its `num == layer_num-1` branch is never taken
            for num in range(layer_num - 1):
                self.convs.append(AdaptedConv(dim_in if num == 0 else hidden, hidden, root_weight=root_weight))
                if use_bn:
                    self.bns.append(nn.BatchNorm1d(hidden))

    def get_emb(self, data):
        ei1, ei2, ei = self._graph(data)
        return self._hidden(data.x, ei, ei1, ei2, data.central_mask, len(self.convs) - 1)

    def forward(self, data):
        ei1, ei2, ei = self._graph(data)
        c = data.central_mask
        x = self._hidden(data.x, ei, ei1, ei2, c, len(self.convs) - 1)
        x = self.convs[-1](x, ei, ei1, ei2, c)
        return F.log_softmax(x, dim=1), None
