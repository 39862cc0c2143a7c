"""Kernel-level breakdown of the graph preparation a NEW graph costs (graph_partition + CSR / transposed CSR / orders),
i.e. the part of bench.py's e2e step that is neither copy nor training step.  Usage: python tools/profile_e2e.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bridged_gnn_b200.data import Data, to_undirected  # noqa: E402
from bridged_gnn_b200.models import KTGNN_no_complement  # noqa: E402


def main():
    n = 1 << 20
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
    model = KTGNN_no_complement(bench.DIM, bench.N_CLASS, 2, bench.HIDDEN, root_weight=False, use_bn=True,
                                dim_share=bench.DIM, dropout=0.0).to(dev).train()

    def prep():
        d = Data(x=None, edge_index=ei.clone(), central_mask=cm)
        model.edge_index = None
        model.prepare_graph(d)

    for _ in range(3):
        prep()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        prep()
    torch.cuda.synchronize()
    print("prepare_graph: %.2f ms per call (E = %d)" % ((time.perf_counter() - t0) / 5 * 1e3, ei.shape[1]))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            prep()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))


if __name__ == "__main__":
    main()
