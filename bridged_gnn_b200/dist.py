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
    """[n_loc, C] per rank -> [world * n_loc, C] on every rank (the dense feature halo: kNN edges have no
    locality, so a destination shard needs H of essentially every source).  Backward = reduce-scatter of the
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


class DstPartition:
    """1-D partition of the node set by destination rows for message passing: rank r owns the
    contiguous rows [r*n_loc, (r+1)*n_loc) of a node set padded to a multiple of the world size, keeps the
    edges whose destination it owns (global source ids), and needs H of all nodes for the gather."""

    def __init__(self, num_nodes, group=None):
        self.group = group
        self.rank, self.world = _world(group)
        self.n = int(num_nodes)
        self.n_loc = (self.n + self.world - 1) // self.world
        self.n_pad = self.n_loc * self.world
        self.r0 = self.rank * self.n_loc
        self.r1 = min(self.n, self.r0 + self.n_loc)

    def local_rows(self, t):
        """Rows of a global per-node tensor owned by this rank, zero padded to n_loc."""
        out = t.new_zeros((self.n_loc,) + tuple(t.shape[1:]))
        if self.r1 > self.r0:
            out[: self.r1 - self.r0] = t[self.r0:self.r1]
        return out

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

    def sync_grads(self, module):
        """Sum parameter gradients over ranks (every rank back-propagated its own destination rows)."""
        for p in module.parameters():
            if p.grad is not None:
                dist.all_reduce(p.grad, group=self.group)
