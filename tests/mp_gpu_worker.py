"""Multi-rank GPU checks, launched by tests/test_gpu_multi.py through torch.distributed.run (one process per GPU,
NCCL).  Every rank also computes the single-GPU result itself and compares:
  (a) row-sharded kNN build + all-gather  == 1-GPU lists, bit for bit
  (b) destination-partitioned KT-GNN      == 1-GPU model: log-probs and every parameter gradient (train mode, BatchNorm
      with batch statistics, the multi-head classifier pass, the domain-aware halo exchange)
  (c) panel-pipelined partitioned SpMM    == 1-GPU SpMM, bit for bit
Prints one line 'MP_GPU_OK world=<n>' from rank 0 on success; any failure raises."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from bridged_gnn_b200 import dist as bd
    from bridged_gnn_b200 import ops
    from bridged_gnn_b200.data import Data, to_undirected
    from bridged_gnn_b200.models import KTGNN_no_complement, graph_partition

    g = torch.Generator(device=dev).manual_seed(0)          # same seed on every rank: identical inputs
    # ---- (a) sharded build ----------------------------------------------------------------------------------
    for nq, ndb, d, k in ((1000, 30000, 128, 20), (257, 5000, 64, 7)):
        q, db = torch.randn(nq, d, generator=g, device=dev), torch.randn(ndb, d, generator=g, device=dev)
        i0, v0, g0, _ = ops.knn_cosine(q, db, k, algo="f16")
        idx, val, gap = bd.sharded_topk(q, db, k, lambda a, b, kk: ops.knn_cosine(a, b, kk, algo="f16")[:3])
        assert torch.equal(idx, i0) and torch.equal(val, v0) and torch.equal(gap, g0), "sharded kNN differs from 1-GPU"

    # ---- (b) partitioned KT-GNN -------------------------------------------------------------------------------
    for n, ns, f_in, n_class, balanced in ((4096, 3072, 128, 2, False), (2048, 1024, 64, 3, False), (4096, 3072, 128, 2, True)):
        x = torch.randn(n, f_in, generator=g, device=dev)
        y = torch.randint(0, n_class, (n,), generator=g, device=dev)
        ei = to_undirected(torch.randint(0, n, (2, 12 * n), generator=g, device=dev), n)
        cm = torch.zeros(n, dtype=torch.bool, device=dev)
        cm[:ns] = True
        tm = torch.rand(n, generator=g, device=dev) < 0.7
        torch.manual_seed(0)
        ref = KTGNN_no_complement(f_in, n_class, 2, 64, root_weight=False, use_bn=True, dim_share=f_in, dropout=0.0).to(dev)
        mod = KTGNN_no_complement(f_in, n_class, 2, 64, root_weight=False, use_bn=True, dim_share=f_in, dropout=0.0).to(dev)
        mod.load_state_dict(ref.state_dict())
        ref.train(), mod.train()
        nll = torch.nn.functional.nll_loss
        cnt = int(tm.sum())
        out_r = ref(Data(x=x, edge_index=ei, central_mask=cm))
        (sum(nll(o[tm], y[tm], reduction="sum") for o in out_r[:3]) / cnt).backward()
        _, _, ei_all = graph_partition(ei, cm)
        part = bd.DstPartition(n, bounds=bd.DstPartition.balanced_bounds(ei_all[1], n, world) if balanced else None)
        data_loc = Data(x=part.local_rows(x), edge_index=part.local_edges(ei_all), central_mask=part.pad_rows(cm), part=part)
        out = mod(data_loc)
        tm_loc, y_loc = part.local_rows(tm), part.local_rows(y)
        (sum(nll(o[tm_loc], y_loc[tm_loc], reduction="sum") for o in out[:3]) / cnt).backward()
        part.sync_grads(mod)
        for o, o_r in zip(out[:3], out_r[:3]):
            want = o_r[part.r0:part.r1]
            err = float((o[: part.r1 - part.r0] - want).abs().max())
            assert err <= 1e-5 * float(want.abs().max()) + 1e-6, "partitioned log-probs differ: %.3e" % err
        for (name, p), q in zip(mod.named_parameters(), ref.parameters()):
            err = float((p.grad - q.grad).abs().max())
            assert err <= 5e-5 * float(q.grad.abs().max()) + 1e-7, "partitioned grad of %s differs: %.3e (max %.3e)" % (
                name, err, float(q.grad.abs().max()))
        for (name, b), b_r in zip(mod.named_buffers(), ref.buffers()):
            if b.dtype.is_floating_point:
                assert float((b - b_r).abs().max()) <= 1e-5 * float(b_r.abs().max()) + 1e-6, "running statistics differ: " + name

    # ---- (c) partitioned SpMM -----------------------------------------------------------------------------------
    n, f = 8192, 256
    X = torch.randn(n, f, generator=g, device=dev)
    ei = to_undirected(torch.randint(0, n, (2, 10 * n), generator=g, device=dev), n)
    part = bd.DstPartition(n)
    full = ops.spmm(ops.CSRGraph(ei, n), X, reduce="mean")
    for transport in ("nccl", "peer", "peer"):       # (twice over peer memory: the buffers are reused across calls)
        y_loc = bd.partitioned_spmm(part.graph(part.local_edges(ei)), part.local_rows(X), part, reduce="mean", panels=4,
                                    transport=transport)
        assert torch.equal(y_loc, full[part.r0:part.r1]), "partitioned SpMM (%s) differs from 1-GPU" % transport
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        print("MP_GPU_OK world=%d" % world, flush=True)
    sys.stdout.flush()
    os._exit(0)      # (symmetric-memory handles and NCCL teardown: leave without the slow destructors)


if __name__ == "__main__":
    main()
