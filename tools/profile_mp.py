"""Kernel-level breakdown of one KT-GNN training step on the bench workload (torch.profiler, CUDA time).
Usage: python tools/profile_mp.py [n_log2=20]   -> prints the top kernels by device time."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bridged_gnn_b200.data import Data, to_undirected  # noqa: E402
from bridged_gnn_b200.models import KTGNN_no_complement  # noqa: E402


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    n = 1 << lg
    ns, nt = n * 3 // 4, n // 4
    dev = torch.device("cuda:0")
    u_s, u_t, y_s, y_t = bench.make_sync_embeddings(ns, nt, bench.DIM, dev)
    y = torch.cat((y_s, y_t))
    rnd = bench.make_random_edges(y, bench.RAND_EDGES_PER_NODE, bench.HOMOPHILY, dev)
    tar = torch.arange(ns, n, device=dev).repeat_interleave(bench.K_CROSS)
    src = torch.randint(0, ns, (tar.numel(),), device=dev)
    ei = to_undirected(torch.cat((rnd, torch.stack((src, tar))), 1), n)
    cm = torch.zeros(n, dtype=torch.bool, device=dev)
    cm[:ns] = True
    data = Data(x=torch.cat((u_s, u_t)).contiguous(), edge_index=ei, y=y, central_mask=cm)
    model = KTGNN_no_complement(bench.DIM, bench.N_CLASS, 2, bench.HIDDEN, root_weight=False, use_bn=True,
                                dim_share=bench.DIM, dropout=0.0).to(dev).train()

    w_train = cm.to(torch.float32) / cm.sum()
    y_col = y.unsqueeze(1)

    def nll(lp):
        return -(lp.gather(1, y_col).squeeze(1) * w_train).sum()

    def step():
        model.zero_grad(set_to_none=True)
        lb, lt, ltt, _ = model(data)
        loss = nll(lb) + nll(lt) + nll(ltt)
        loss.backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=70))


if __name__ == "__main__":
    main()
