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
