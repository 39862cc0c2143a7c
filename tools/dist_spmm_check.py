"""Panel-pipelined partitioned SpMM (dist.partitioned_spmm): local kernel only, all-gather only, one all-gather then the
kernel, and the pipeline at 2 / 4 / 8 column panels -- on the default NCCL group and on one with a high-priority stream.
    torchrun --nproc-per-node N tools/dist_spmm_check.py [log2_nodes=22] [F=256]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bridged_gnn_b200 import dist as bd, ops  # noqa: E402
from bridged_gnn_b200.data import to_undirected  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    f = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    n = 1 << lg
    g = torch.Generator(device=dev).manual_seed(0)
    ei = to_undirected(torch.randint(0, n, (2, 13 * n), generator=g, device=dev), n)
    opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
    hp = dist.new_group(pg_options=opts)
    for name, group in (("default", None), ("high-priority", hp)):
        part = bd.DstPartition(n, group)
        graph = part.graph(part.local_edges(ei))
        x_loc = torch.randn(part.n_loc, f, generator=g, device=dev)
        full = torch.empty((part.n_pad, f), device=dev)

        def timed(fn, steps=5, warm=2):
            for _ in range(warm):
                fn()
            dist.barrier(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                fn()
            b.record()
            dist.barrier(); torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / steps], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        k = timed(lambda: ops._spmm_raw(graph.rowptr, graph.col, full, graph.n_rows, True))
        ag = timed(lambda: dist.all_gather_into_tensor(full, x_loc, group=part.group))
        res = {p: timed(lambda: bd.partitioned_spmm(graph, x_loc, part, "mean", panels=p, transport="nccl")) for p in (1, 2, 4, 8)}
        resp = {p: timed(lambda: bd.partitioned_spmm(graph, x_loc, part, "mean", panels=p, transport="peer")) for p in (1, 2, 4, 8)}
        if rank == 0:
            print("%s group: n=%d e_loc=%d F=%d | kernel only %.2f ms, all-gather only %.2f ms (%.0f GB/s in), serial sum %.2f | NCCL panels %s | peer-copy panels %s"
                  % (name, n, graph.e, f, k, ag, (part.n_pad - part.n_loc) * f * 4 / ag / 1e6, k + ag,
                     {p: round(v, 2) for p, v in res.items()}, {p: round(v, 2) for p, v in resp.items()}), flush=True)
        del full, x_loc
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
