"""Multi-GPU plumbing for the bridged-graph build: one process per GPU (torch.distributed, NCCL over
NVLink on B200; gloo in the CPU tests).  The reference is single-process; this layer is new.

Sharding: query (target) rows are split contiguously over ranks, the database (source) rows are
replicated.  Rows are independent, so every rank runs the fused sweep on its shard and the fixed-size
``[rows, k]`` lists are combined with one all-gather; the result is bit-identical to a 1-rank build.
"""
import torch
import torch.distributed as dist


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def row_shard(n, rank, world):
    """Contiguous, balanced [start, end) of ``n`` rows for ``rank`` (the first n % world ranks get one extra)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_gather_rows(local, n_total, group=None):
    """Concatenate per-rank row blocks (shapes differ by at most one row) into ``[n_total, ...]`` on every rank."""
    rank, world = _world(group)
    if world == 1:
        return local
    max_rows = (n_total + world - 1) // world
    pad = local.new_zeros((max_rows,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    out = local.new_empty((world * max_rows,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    parts = []
    for r in range(world):
        s, e = row_shard(n_total, r, world)
        parts.append(out[r * max_rows: r * max_rows + (e - s)])
    return torch.cat(parts, 0)


def sharded_topk(q, db, k, local_fn, group=None, q_is_sharded=False, n_total=None):
    """Per-row top-k of ``q`` against the replicated ``db``, rows sharded over the ranks of ``group``.

    ``local_fn(q_rows, db, k) -> (idx, val, gap)`` is the single-GPU build (``ops.knn_cosine`` /
    ``ops.knn_addrelu`` wrappers).  ``q`` is either the full query matrix (each rank slices its shard) or,
    with ``q_is_sharded``, this rank's block of an ``n_total``-row matrix.  Returns full ``[n_total, k]``
    lists on every rank."""
    rank, world = _world(group)
    if q_is_sharded:
        n = int(n_total)
        q_loc = q
    else:
        n = q.shape[0]
        s, e = row_shard(n, rank, world)
        q_loc = q[s:e]
    idx, val, gap = local_fn(q_loc, db, k)
    return all_gather_rows(idx, n, group), all_gather_rows(val, n, group), all_gather_rows(gap, n, group)


def edges_from_topk(idx, row_offset=0):
    """(neighbour, query) edge list of a ``[rows, k]`` index block whose first row has global id ``row_offset``."""
    nq, k = idx.shape
    to = torch.arange(row_offset, row_offset + nq, device=idx.device).unsqueeze(1).expand(nq, k)
    return torch.stack((idx.reshape(-1), to.reshape(-1)), 0)


# ----------------------------------------------------------------------------------- destination-partitioned aggregation
class _AllGatherRows(torch.autograd.Function):
    """[n_loc, C] per rank -> [world * n_loc, C] on every rank (dense feature halo).  Backward = reduce-scatter of the
    gradients w.r.t. all rows, i.e. the transpose exchange of SURVEY 8e."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        world = dist.get_world_size(group)
        out = x.new_empty((world * x.shape[0],) + tuple(x.shape[1:]))
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        world = dist.get_world_size(ctx.group)
        rank = dist.get_rank(ctx.group)
        n_loc = g.shape[0] // world
        g = g.contiguous()
        if dist.get_backend(ctx.group) == "nccl":
            gl = g.new_empty((n_loc,) + tuple(g.shape[1:]))
            dist.reduce_scatter_tensor(gl, g, group=ctx.group)
        else:                                   # gloo has no reduce-scatter: all-reduce, keep the own block
            dist.all_reduce(g, group=ctx.group)
            gl = g[rank * n_loc:(rank + 1) * n_loc].clone()
        return gl, None


class _AllReduceSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        y = x.clone()
        dist.all_reduce(y, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        g = g.clone()
        dist.all_reduce(g, group=ctx.group)
        return g, None


def all_gather_rows_autograd(x, group=None):
    return _AllGatherRows.apply(x, group)


def all_reduce_sum_autograd(x, group=None):
    return _AllReduceSum.apply(x, group)


def _exchange(ops_):
    if ops_:
        for req in dist.batch_isend_irecv(ops_):
            req.wait()


class _HaloExchange(torch.autograd.Function):
    """Domain-aware halo exchange of the AdaptedConv operands.  A destination row of the source domain gathers
    ``h_s`` of its in-neighbours, one of the target domain ``h_t`` (models/KTGNN.py:292-293), so a rank needs ALL rows
    of ``h_s`` only if it owns a source-domain destination row, and of ``h_t`` only if it owns a target-domain one.
    With source nodes first in the id order (merge_graphs, main_bridged_graph.py:163-193) most ranks of a contiguous
    row partition own rows of ONE domain and receive half of what an all-gather of both operands would deliver.

    forward : (h_s [n_loc, C], h_t [n_loc, C]) -> (H_s [n_pad, C] or None, H_t [n_pad, C] or None), point-to-point
              sends of the own block to every rank that needs it (one grouped NCCL send/recv over NVLink)
    backward: the transposed exchange -- every rank that needed H_x sends the gradient block of rank i's rows to i,
              which adds the blocks up in rank order (deterministic)."""

    @staticmethod
    def forward(ctx, h_s, h_t, part, needs):
        rank, world, n_loc = part.rank, part.world, part.n_loc
        ctx.part, ctx.needs, ctx.shape = part, needs, tuple(h_s.shape[1:])
        h = (h_s.contiguous(), h_t.contiguous())
        outs, p2p = [], []
        for d in (0, 1):
            if needs[rank][d]:
                H = h[d].new_empty((part.n_pad,) + ctx.shape)
                H[rank * n_loc:(rank + 1) * n_loc] = h[d]
                for j in range(world):
                    if j != rank:
                        p2p.append(dist.P2POp(dist.irecv, H[j * n_loc:(j + 1) * n_loc], part.global_rank(j), group=part.group))
                outs.append(H)
            else:
                outs.append(None)
            for j in range(world):
                if j != rank and needs[j][d]:
                    p2p.append(dist.P2POp(dist.isend, h[d], part.global_rank(j), group=part.group))
        _exchange(p2p)
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_s, g_t):
        part, needs = ctx.part, ctx.needs
        rank, world, n_loc = part.rank, part.world, part.n_loc
        ref = g_s if g_s is not None else g_t
        g, p2p, stage = [], [], [None, None]
        for d, gd in enumerate((g_s, g_t)):
            if needs[rank][d]:
                # an operand this rank received but whose gradient autograd did not produce counts as zero
                gd = ref.new_zeros((part.n_pad,) + ctx.shape) if gd is None else gd.contiguous()
                for i in range(world):
                    if i != rank:
                        p2p.append(dist.P2POp(dist.isend, gd[i * n_loc:(i + 1) * n_loc], part.global_rank(i), group=part.group))
            else:
                gd = None
            g.append(gd)
            senders = [j for j in range(world) if j != rank and needs[j][d]]
            if senders:
                stage[d] = ref.new_empty((len(senders), n_loc) + ctx.shape)
                for k, j in enumerate(senders):
                    p2p.append(dist.P2POp(dist.irecv, stage[d][k], part.global_rank(j), group=part.group))
        _exchange(p2p)
        res = []
        for d in (0, 1):
            acc = None if g[d] is None else g[d][rank * n_loc:(rank + 1) * n_loc]
            if stage[d] is not None:
                recv = stage[d].sum(0) if stage[d].shape[0] > 1 else stage[d][0]
                acc = recv if acc is None else acc + recv
            elif acc is not None:
                acc = acc.clone()
            if acc is None:           # no rank aggregated with this operand: zero gradient
                acc = ref.new_zeros((n_loc,) + ctx.shape)
            res.append(acc)
        return res[0], res[1], None, None


def halo_exchange(h_s, h_t, part, central_mask):
    """(H_s, H_t) over all nodes for this rank's aggregation; an operand this rank's destination rows never read is
    None.  ``central_mask``: the global (padded) domain mask, identical on every rank."""
    return _HaloExchange.apply(h_s, h_t, part, part.needs(central_mask))


class DstPartition:
    """1-D partition of the node set by destination rows for message passing: rank r owns the
    contiguous rows [r*n_loc, (r+1)*n_loc) of a node set padded to a multiple of the world size, keeps the
    edges whose destination it owns (global source ids), and needs H of the other nodes for the gather."""

    def __init__(self, num_nodes, group=None):
        self.group = group
        self.rank, self.world = _world(group)
        self.n = int(num_nodes)
        self.n_loc = (self.n + self.world - 1) // self.world
        self.n_pad = self.n_loc * self.world
        self.r0 = self.rank * self.n_loc
        self.r1 = min(self.n, self.r0 + self.n_loc)
        self._needs = None
        self._loc = {}

    def global_rank(self, r):
        return r if self.group is None else dist.get_global_rank(self.group, r)

    def needs(self, central_mask):
        """needs[r] = (rank r owns a source-domain row, rank r owns a target-domain row), for every rank: computed
        once per mask (one host read of 2 * world flags)."""
        key = (central_mask, central_mask._version)
        if self._needs is None or self._needs[0][0] is not central_mask or self._needs[0][1] != central_mask._version:
            c = self.pad_rows(central_mask[: self.n].to(torch.int64)).view(self.world, self.n_loc)
            valid = self.pad_rows(torch.ones(self.n, dtype=torch.int64, device=central_mask.device)).view(self.world, self.n_loc)
            flags = torch.stack((c.sum(1) > 0, (valid - c).sum(1) > 0), 1).tolist()
            self._needs = (key, tuple((bool(a), bool(b)) for a, b in flags))
        return self._needs[1]

    def local_rows(self, t):
        """Rows of a global per-node tensor owned by this rank, zero padded to n_loc."""
        out = t.new_zeros((self.n_loc,) + tuple(t.shape[1:]))
        if self.r1 > self.r0:
            out[: self.r1 - self.r0] = t[self.r0:self.r1]
        return out

    def local_cached(self, name, t, fn):
        """``fn(t)`` cached per (name, tensor identity, version): per-rank slices of the global masks are built once,
        not on every forward."""
        ent = self._loc.get(name)
        if ent is None or ent[0] is not t or ent[1] != t._version:
            ent = (t, t._version, fn(t))
            self._loc[name] = ent
        return ent[2]

    def pad_rows(self, t, fill=0):
        if t.shape[0] == self.n_pad:
            return t
        out = t.new_full((self.n_pad,) + tuple(t.shape[1:]), fill)
        out[: t.shape[0]] = t
        return out

    def local_edges(self, edge_index):
        """Edges whose destination this rank owns (ids stay global)."""
        m = (edge_index[1] >= self.r0) & (edge_index[1] < self.r1)
        return edge_index[:, m].contiguous()

    def graph(self, edge_index):
        """Cached CSR of this rank's edges: local destination rows, global source columns."""
        from . import ops
        return ops.cached_graph(edge_index, self.n_pad, self.n_loc, self.r0)

    def sync_grads(self, module):
        """Sum parameter gradients over ranks (every rank back-propagated its own destination rows): ONE all-reduce
        of the flattened gradients instead of one per parameter."""
        grads = [p.grad for p in module.parameters() if p.grad is not None]
        if not grads or self.world == 1:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, group=self.group)
        off = 0
        for g in grads:
            k = g.numel()
            g.copy_(flat[off:off + k].view_as(g))
            off += k


def partitioned_spmm(graph, x_loc, part, reduce="mean", panels=4):
    """Y_loc = reduce_{j -> i} X[j] for this rank's destination rows when X is row-partitioned too (config 5: the
    SAGE / GCN aggregation of a graph too large for one GPU).  The dense halo -- every rank needs (nearly) all rows of
    X, kNN edges have no locality -- is gathered COLUMN PANEL by column panel: all panel all-gathers are queued on the
    NCCL stream up front and the SpMM of panel p starts as soon as panel p has landed, so the transfer of panels
    p+1.. overlaps the gather kernel of panel p.  Forward only (inference / diagnostics)."""
    from . import ops
    f = x_loc.shape[1]
    panels = max(1, min(panels, f // 4 if f >= 4 else 1))
    bounds = [(f * p // panels) // 4 * 4 for p in range(panels)] + [f]
    y = torch.empty((graph.n_rows, f), dtype=torch.float32, device=x_loc.device)
    works, bufs = [], []
    for p in range(panels):
        lo, hi = bounds[p], bounds[p + 1]
        src = x_loc[:, lo:hi].contiguous()
        buf = src.new_empty((part.n_pad, hi - lo))
        works.append(dist.all_gather_into_tensor(buf, src, group=part.group, async_op=True))
        bufs.append(buf)
    for p in range(panels):
        lo, hi = bounds[p], bounds[p + 1]
        works[p].wait()
        ops._spmm_raw(graph.rowptr, graph.col, bufs[p], graph.n_rows, reduce == "mean", out=y[:, lo:hi])
    return y
