"""Kernel-level timing of the fused AdaptedConv aggregation (fwd / bwd) and SpMM on the bench graph.
Usage: python tools/bench_gat.py [n_log2=20] [reps=10] [widths=2,31,64,128,256]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bridged_gnn_b200 import _lib, ops  # noqa: E402
from bridged_gnn_b200.data import to_undirected  # noqa: E402
from bridged_gnn_b200.models import graph_partition  # noqa: E402


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    n = 1 << lg
    ns, nt = n * 3 // 4, n // 4
    dev = torch.device("cuda:0")
    _, _, y_s, y_t = bench.make_sync_embeddings(ns, nt, 8, dev)
    y = torch.cat((y_s, y_t))
    rnd = bench.make_random_edges(y, bench.RAND_EDGES_PER_NODE, bench.HOMOPHILY, dev)
    tar = torch.arange(ns, n, device=dev).repeat_interleave(bench.K_CROSS)
    if os.environ.get("BGNN_RANDOM_KNN"):
        src = torch.randint(0, ns, (tar.numel(),), device=dev)
    else:      # the bench graph itself: real kNN edges (hub rows, in-degree up to ~1800)
        u_s, u_t, _, _ = bench.make_sync_embeddings(ns, nt, bench.DIM, dev)
        src = ops.knn_cosine(u_t, u_s, bench.K_CROSS)[0].reshape(-1)
        del u_s, u_t
    ei = to_undirected(torch.cat((rnd, torch.stack((src, tar))), 1), n)
    cm = torch.zeros(n, dtype=torch.bool, device=dev)
    cm[:ns] = True
    _, _, eall = graph_partition(ei, cm)
    graph = ops.CSRGraph(eall, n)
    _ = graph.t, graph.csr_to_csc
    if os.environ.get("BGNN_HACK_SEQ_RECORDS"):
        # timing experiment only (results are garbage): pass A writes its per-edge records in CSR order instead of
        # scattering them to their transposed-CSR slots -- what does the scatter cost?  (profiles/r03k)
        graph._csr_to_csc = torch.arange(graph.e, dtype=torch.int32, device=dev)
    e = graph.e
    cm8 = cm.to(torch.uint8)
    peak = bench.load_peaks()["hbm_gbs"]
    print("n=%d e=%d" % (n, e))
    widths = [int(t) for t in sys.argv[3].split(",")] if len(sys.argv) > 3 else [2, 31, 64, 128, 256]
    for c in widths:
        Hs = torch.randn(n, c, device=dev, requires_grad=True)
        Ht = torch.randn(n, c, device=dev, requires_grad=True)
        a1 = torch.randn(c, device=dev, requires_grad=True)
        a2 = torch.randn(c, device=dev, requires_grad=True)
        gout = torch.randn(n, c, device=dev)
        for _ in range(2):
            out = ops.gat_aggregate(Hs, Ht, a1, a2, graph, cm8, 0.1)
            out.backward(gout)
        _lib.start_timing()
        for _ in range(reps):
            out = ops.gat_aggregate(Hs, Ht, a1, a2, graph, cm8, 0.1)
            out.backward(gout)
        t = _lib.stop_timing()
        fwd = t["bgnn_gatv2_fwd_f32[c=%d]" % c][1] / reps
        bwd = t["bgnn_gatv2_bwd_f32[c=%d]" % c][1] / reps
        fb = e * (4 + 4 * c) + n * (13 + 8 * c)
        bb = 2 * e * (4 + 8 * c) + n * 8 * c
        x = torch.randn(n, c, device=dev)
        for _ in range(2):
            ops.spmm(graph, x, "mean")
        _lib.start_timing()
        for _ in range(reps):
            ops.spmm(graph, x, "mean")
        sp = _lib.stop_timing()["bgnn_spmm_csr_f32"][1] / reps
        sb = e * (4 + 4 * c) + n * (4 + 4 * c)
        print("c=%3d  gat fwd %.3f ms (%.0f GB/s, %.2f of HBM peak)  bwd %.3f ms (%.0f GB/s, %.2f)  spmm-mean %.3f ms (%.0f GB/s, %.2f)  fwd %.2f GEdges/s"
              % (c, fwd, fb / fwd / 1e6, fb / fwd / 1e6 / peak, bwd, bb / bwd / 1e6, bb / bwd / 1e6 / peak, sp, sb / sp / 1e6,
                 sb / sp / 1e6 / peak, e / fwd / 1e6))
        del Hs, Ht, gout, out, x


if __name__ == "__main__":
    main()
