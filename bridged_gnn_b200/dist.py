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
        rank, world = part.rank, part.world
        ctx.part, ctx.needs, ctx.shape = part, needs, tuple(h_s.shape[1:])
        h = (h_s.contiguous(), h_t.contiguous())
        outs, p2p = [], []
        for d in (0, 1):
            if needs[rank][d]:
                H = h[d].new_empty((part.n_pad,) + ctx.shape)
                b0, b1 = part.block(rank)
                H[b0:b1] = h[d]
                for j in range(world):
                    if j != rank and part.block_rows(j):
                        j0, j1 = part.block(j)
                        p2p.append(dist.P2POp(dist.irecv, H[j0:j1], part.global_rank(j), group=part.group))
                outs.append(H)
            else:
                outs.append(None)
            for j in range(world):
                if j != rank and needs[j][d] and part.block_rows(rank):
                    p2p.append(dist.P2POp(dist.isend, h[d], part.global_rank(j), group=part.group))
        _exchange(p2p)
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_s, g_t):
        part, needs = ctx.part, ctx.needs
        rank, world = part.rank, part.world
        my0, my1 = part.block(rank)
        n_mine = my1 - my0
        ref = g_s if g_s is not None else g_t
        g, p2p, stage = [], [], [None, None]
        for d, gd in enumerate((g_s, g_t)):
            if needs[rank][d]:
                # an operand this rank received but whose gradient autograd did not produce counts as zero
                gd = ref.new_zeros((part.n_pad,) + ctx.shape) if gd is None else gd.contiguous()
                for i in range(world):
                    if i != rank and part.block_rows(i):
                        i0, i1 = part.block(i)
                        p2p.append(dist.P2POp(dist.isend, gd[i0:i1], part.global_rank(i), group=part.group))
            else:
                gd = None
            g.append(gd)
            senders = [j for j in range(world) if j != rank and needs[j][d]]
            if senders and n_mine:
                stage[d] = ref.new_empty((len(senders), n_mine) + ctx.shape)
                for k, j in enumerate(senders):
                    p2p.append(dist.P2POp(dist.irecv, stage[d][k], part.global_rank(j), group=part.group))
        _exchange(p2p)
        res = []
        for d in (0, 1):
            acc = None if g[d] is None else g[d][my0:my1]
            if stage[d] is not None:
                recv = stage[d].sum(0) if stage[d].shape[0] > 1 else stage[d][0]
                acc = recv if acc is None else acc + recv
            elif acc is not None:
                acc = acc.clone()
            if acc is None:           # no rank aggregated with this operand: zero gradient
                acc = ref.new_zeros((n_mine,) + ctx.shape)
            res.append(acc)
        return res[0], res[1], None, None


def halo_exchange(h_s, h_t, part, central_mask):
    """(H_s, H_t) over all nodes for this rank's aggregation; an operand this rank's destination rows never read is
    None.  ``central_mask``: the global (padded) domain mask, identical on every rank."""
    return _HaloExchange.apply(h_s, h_t, part, part.needs(central_mask))


class DstPartition:
    """1-D partition of the node set by destination rows for message passing: rank r owns a contiguous
    block of rows, keeps the edges whose destination it owns (global source ids), and needs H of the other nodes for
    the gather.

    Default: equal blocks of ``n_loc`` rows over a node set padded to ``n_pad = world * n_loc`` (what the all-gather
    based paths need).  With ``bounds`` (world + 1 increasing row offsets, 0 .. num_nodes; ``balanced_bounds`` cuts at
    equal work) the blocks differ in size, nothing is padded, and the exchanges are point-to-point."""

    def __init__(self, num_nodes, group=None, bounds=None):
        if group is None and dist.is_available() and dist.is_initialized():
            group = dist.group.WORLD       # a concrete group: `None` means "not partitioned" to the operators
        self.group = group
        self.rank, self.world = _world(group)
        self.n = int(num_nodes)
        if bounds is None:
            self.uniform = True
            self.n_loc = (self.n + self.world - 1) // self.world
            self.n_pad = self.n_loc * self.world
            self.bounds = [r * self.n_loc for r in range(self.world + 1)]
            self.r0 = self.rank * self.n_loc
            self.r1 = min(self.n, self.r0 + self.n_loc)
        else:
            bounds = [int(b) for b in bounds]
            if len(bounds) != self.world + 1 or bounds[0] != 0 or bounds[-1] != self.n or any(b1 < b0 for b0, b1 in zip(bounds, bounds[1:])):
                raise ValueError("bounds must be world + 1 non-decreasing offsets from 0 to num_nodes")
            self.uniform = False
            self.bounds = bounds
            self.n_pad = self.n
            self.r0, self.r1 = bounds[self.rank], bounds[self.rank + 1]
            self.n_loc = self.r1 - self.r0
        self._needs = None
        self._loc = {}

    def block(self, r):
        """[start, end) of rank r's block in the (padded) global row space."""
        return self.bounds[r], self.bounds[r + 1]

    def block_rows(self, r):
        return self.bounds[r + 1] - self.bounds[r]

    @staticmethod
    def balanced_bounds(dst, num_nodes, world, row_weight=12.0, n_prefix=None, pure_slack=1.12):
        """Row offsets that give every rank the same WORK (+- one row): cut points of the cumulative cost
        ``in-degree + row_weight`` per row.  The aggregation kernels cost per edge, the node-wise dense layers per row;
        measured on the sync-1M KT-GNN step one row costs about as much as 12 edges (6.6 ms of aggregation for 2.2e7
        edges, 3.7 ms of dense kernels for 1e6 rows).  ``dst``: destination ids of the edge list (device tensor).

        ``n_prefix``: number of source-domain nodes when they form a prefix of the id order (merge_graphs puts them
        first).  A rank whose block straddles the domain boundary needs BOTH halo operands -- twice the exchange of every
        other rank, on the critical path.  So the ranks are also split between the two domains in proportion to the
        domains' work, each domain cut evenly among its ranks, and that domain-pure partition is taken when its heaviest
        rank is within ``pure_slack`` of the plainly balanced one.  One host read of a few numbers."""
        cost = torch.bincount(dst, minlength=num_nodes).to(torch.float64) + float(row_weight)
        cum = torch.cumsum(cost, 0)

        def cuts_of(lo, hi, parts):         # parts - 1 interior cut points of rows [lo, hi) at equal cost
            if parts <= 1:
                return []
            base = cum[lo - 1] if lo > 0 else cum.new_zeros(())
            total = cum[hi - 1] - base
            targets = base + torch.arange(1, parts, device=dst.device, dtype=torch.float64) * (total / parts)
            return (torch.searchsorted(cum, targets) + 1).clamp(min=lo, max=hi).tolist()

        def finish(cuts):
            b = [0] + [int(c) for c in cuts] + [int(num_nodes)]
            for i in range(1, len(b)):
                b[i] = max(b[i], b[i - 1])
            return b

        plain = finish(cuts_of(0, num_nodes, world))
        if n_prefix is None or world < 2 or not (0 < n_prefix < num_nodes):
            return plain

        def heaviest(b):
            ends = torch.tensor(b, device=dst.device)
            c0 = torch.cat((cum.new_zeros(1), cum))[ends]
            return float((c0[1:] - c0[:-1]).max())
        w_src = float(cum[n_prefix - 1])
        w_all = float(cum[-1])
        r_src = min(world - 1, max(1, int(round(world * w_src / w_all))))
        pure = finish(cuts_of(0, n_prefix, r_src) + [n_prefix] + cuts_of(n_prefix, num_nodes, world - r_src))
        return pure if heaviest(pure) <= pure_slack * heaviest(plain) else plain

    def global_rank(self, r):
        return r if self.group is None or self.group is dist.group.WORLD else dist.get_global_rank(self.group, r)

    def needs(self, central_mask):
        """needs[r] = (rank r owns a source-domain row, rank r owns a target-domain row), for every rank: computed
        once per mask (one host read of 2 * world flags)."""
        key = (central_mask, central_mask._version)
        if self._needs is None or self._needs[0][0] is not central_mask or self._needs[0][1] != central_mask._version:
            cs = torch.zeros(self.n + 1, dtype=torch.int64, device=central_mask.device)
            cs[1:] = torch.cumsum(central_mask[: self.n].to(torch.int64), 0)
            ends = torch.tensor([min(b, self.n) for b in self.bounds], device=central_mask.device)
            src_cnt = (cs[ends[1:]] - cs[ends[:-1]]).tolist()              # source-domain rows per rank
            rows = (ends[1:] - ends[:-1]).tolist()
            self._needs = (key, tuple((s_ > 0, r_ - s_ > 0) for s_, r_ in zip(src_cnt, rows)))
        return self._needs[1]

    def local_rows(self, t):
        """Rows of a global per-node tensor owned by this rank (zero padded to n_loc in the equal-block layout)."""
        if not self.uniform:
            return t[self.r0:self.r1].contiguous()
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


class _PeerPanels:
    """Persistent symmetric-memory buffer [panels, n_pad, w] of one partitioned SpMM shape: every rank can write into
    every other rank's copy over NVLink with plain device-to-device copies, which run on the COPY ENGINES -- no SM is
    taken from the gather kernel that runs at the same time, and a pair of B200s moves 760 GB/s this way against
    465 GB/s for NCCL's all-gather (tools/dist_bw_check.py)."""

    def __init__(self, part, panels, w, device):
        import torch.distributed._symmetric_memory as symm
        self.buf = symm.empty((panels, part.n_pad, w), dtype=torch.float32, device=device)
        self.hdl = symm.rendezvous(self.buf, part.group)
        self.peers = [self.hdl.get_buffer(part.global_rank(j), (panels, part.n_pad, w), torch.float32) for j in range(part.world)]
        self.stream = torch.cuda.Stream(device=device, priority=-1)
        self.events = [torch.cuda.Event() for _ in range(panels)]


def _peer_panels(part, panels, w, device):
    key = ("spmm_panels", panels, w, str(device))
    if key not in part._loc:
        try:
            part._loc[key] = _PeerPanels(part, panels, w, device)
        except Exception as ex:             # no symmetric memory on this system / backend: NCCL all-gather instead
            part._loc[key] = ex
    v = part._loc[key]
    return None if isinstance(v, Exception) else v


def partitioned_spmm(graph, x_loc, part, reduce="mean", panels=4, transport="auto"):
    """Y_loc = reduce_{j -> i} X[j] for this rank's destination rows when X is row-partitioned too (config 5: the
    SAGE / GCN aggregation of a graph too large for one GPU).  The dense halo -- every rank needs (nearly) all rows of
    X, kNN edges have no locality -- moves COLUMN PANEL by column panel, and the gather kernel of panel p runs while
    panels p+1.. are still on the wire.

    transport "peer" (default on NVLink systems): every rank pushes its rows of a panel straight into the other ranks'
    symmetric-memory buffers (copy engines, no SMs), a device-side barrier per panel tells the consumers it has landed.
    transport "nccl": one NCCL all-gather per panel, all queued up front on the communicator's stream (give the group a
    high-priority stream, otherwise the all-gather kernels wait behind the gather kernel's CTAs).
    Forward only (inference / diagnostics)."""
    from . import ops
    if not part.uniform:
        raise ValueError("partitioned_spmm works on equal row blocks: use the default (equal-block) DstPartition")
    f = x_loc.shape[1]
    panels = max(1, min(panels, f // 4 if f >= 4 else 1))
    while f % panels or (f // panels) % 4:
        panels -= 1                               # equal panel widths, 16-byte aligned rows
    w = f // panels
    y = torch.empty((graph.n_rows, f), dtype=torch.float32, device=x_loc.device)
    pp = _peer_panels(part, panels, w, x_loc.device) if (transport in ("auto", "peer") and x_loc.is_cuda and part.world > 1) else None
    if transport == "peer" and pp is None:
        raise RuntimeError("symmetric memory is not available for this process group")
    if pp is not None:
        cur = torch.cuda.current_stream(x_loc.device)
        srcs = [x_loc[:, p * w:(p + 1) * w].contiguous() for p in range(panels)]
        pp.stream.wait_stream(cur)
        with torch.cuda.stream(pp.stream):
            pp.hdl.barrier()                      # every rank is done with the previous contents of the buffers
            for p in range(panels):
                for j in range(part.world):       # own copy last: the remote ones are the long ones
                    dst = pp.peers[(part.rank + 1 + j) % part.world]
                    dst[p, part.r0:part.r0 + part.n_loc].copy_(srcs[p], non_blocking=True)
                pp.hdl.barrier()                  # panel p has landed everywhere
                pp.events[p].record(pp.stream)
        for p in range(panels):
            cur.wait_event(pp.events[p])
            ops._spmm_raw(graph.rowptr, graph.col, pp.buf[p], graph.n_rows, reduce == "mean", out=y[:, p * w:(p + 1) * w])
        for t in srcs:
            t.record_stream(pp.stream)
        pp.stream.wait_stream(cur)                # the next call's first barrier comes after this call's last kernel
        return y
    works, bufs = [], []
    for p in range(panels):
        src = x_loc[:, p * w:(p + 1) * w].contiguous()
        buf = src.new_empty((part.n_pad, w))
        works.append(dist.all_gather_into_tensor(buf, src, group=part.group, async_op=True))
        bufs.append(buf)
    for p in range(panels):
        works[p].wait()
        ops._spmm_raw(graph.rowptr, graph.col, bufs[p], graph.n_rows, reduce == "mean", out=y[:, p * w:(p + 1) * w])
    return y
