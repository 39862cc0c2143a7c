"""torchrun --nproc-per-node R tools/dist_mp_check.py : destination-partitioned KT-GNN (NCCL all-gather halo,
reduce-scatter of dH, SyncBatchNorm) vs the single-GPU model on the same graph -- parity, then timing."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bridged_gnn_b200 import dist as bd  # noqa: E402
from bridged_gnn_b200.data import Data, to_undirected  # noqa: E402
from bridged_gnn_b200.models import KTGNN_no_complement, graph_partition  # noqa: E402


def run(n_log2, f_in, n_class, hidden, check, steps=5):
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    n = 1 << n_log2
    ns, nt = n * 3 // 4, n // 4
    u_s, u_t, y_s, y_t = bench.make_sync_embeddings(ns, nt, f_in, dev, n_class=n_class)
    x, y = torch.cat((u_s, u_t)).contiguous(), torch.cat((y_s, y_t))
    rnd = bench.make_random_edges(y, 5, 0.7, dev)
    tar = torch.arange(ns, n, device=dev).repeat_interleave(20)
    src = torch.randint(0, ns, (tar.numel(),), device=dev, generator=torch.Generator(device=dev).manual_seed(3))
    ei = to_undirected(torch.cat((rnd, torch.stack((src, tar))), 1), n)
    cm = torch.zeros(n, dtype=torch.bool, device=dev)
    cm[:ns] = True
    tm = torch.rand(n, device=dev, generator=torch.Generator(device=dev).manual_seed(4)) < 0.5
    torch.manual_seed(0)
    ref = KTGNN_no_complement(f_in, n_class, 2, hidden, root_weight=False, use_bn=True, dim_share=f_in, dropout=0.0).to(dev)
    model = KTGNN_no_complement(f_in, n_class, 2, hidden, root_weight=False, use_bn=True, dim_share=f_in, dropout=0.0).to(dev)
    model.load_state_dict(ref.state_dict())
    model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    part = bd.DstPartition(n)
    _, _, ei_all = graph_partition(ei, cm)
    d_loc = Data(x=part.local_rows(x), edge_index=part.local_edges(ei_all), central_mask=part.pad_rows(cm), part=part)
    tm_loc, y_loc = part.local_rows(tm), part.local_rows(y)
    cnt = int(tm.sum())
    nll = torch.nn.functional.nll_loss

    def step_part():
        model.zero_grad(set_to_none=True)
        out = model(d_loc)
        loss = sum(nll(o[tm_loc], y_loc[tm_loc], reduction="sum") for o in out[:3]) / cnt
        loss.backward()
        part.sync_grads(model)
        return out, loss

    model.train()
    out, loss = step_part()
    if check:
        ref.train()
        d_full = Data(x=x, edge_index=ei, central_mask=cm)
        out_r = ref(d_full)
        loss_r = sum(nll(o[tm], y[tm], reduction="sum") for o in out_r[:3]) / cnt
        loss_r.backward()
        ltot = loss.detach().clone()
        dist.all_reduce(ltot)
        ferr = max(float((o[: part.r1 - part.r0] - r[part.r0:part.r1]).abs().max() / r.abs().max()) for o, r in zip(out[:3], out_r[:3]))
        gmax = max(float(q.grad.abs().max()) for q in ref.parameters())
        # per-parameter relative error; gradients that are zero in exact arithmetic (the bias in front of a
        # BatchNorm) are measured against the largest gradient of the model instead
        def denom(q):
            m = float(q.grad.abs().max())
            return gmax if m < 1e-4 * gmax else m + 1e-3 * gmax      # pure rounding noise: judged on the model's scale
        errs = {k: (float((p.grad - q.grad).abs().max()) / denom(q), float(q.grad.abs().max()))
                for (k, p), q in zip(model.named_parameters(), ref.parameters())}
        gerr = max(v[0] for v in errs.values())
        if rank == 0:
            print("gmax %.2e" % gmax, {k: ("%.1e" % v[0], "%.1e" % v[1]) for k, v in errs.items() if v[0] > 1e-5}, flush=True)
        print("[rank %d] n=2^%d: loss %.6f vs %.6f | logits rel err %.2e | grads rel err %.2e" %
              (rank, n_log2, float(ltot), float(loss_r), ferr, gerr), flush=True)
        assert ferr < 1e-5 and gerr < 1e-4, (ferr, gerr)
    for _ in range(2):
        step_part()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_part()
    torch.cuda.synchronize(); dist.barrier()
    ms = (time.perf_counter() - t0) / steps * 1e3
    e_mp = ei_all.shape[1]
    if rank == 0:
        print("partitioned KT-GNN train step, n=2^%d, E_mp=%d, %d ranks: %.2f ms/step -> %.2f GEdges/s (strong scaling of ONE graph)"
              % (n_log2, e_mp, world, ms, e_mp * 8 / ms / 1e6), flush=True)


def main():
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    run(14, 64, 5, 32, check=True)
    run(20, 128, 2, 64, check=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
