"""Runs only the cosine kNN build on the bench workload shape, for ncu / timing.
Usage: python tools/profile_knn.py [algo=f16] [nq=37888] [ndb=786432] [d=128] [k=20] [reps=3]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bridged_gnn_b200 import ops  # noqa: E402


def main():
    a = sys.argv[1:]
    algo = a[0] if len(a) > 0 else "f16"
    nq = int(a[1]) if len(a) > 1 else 37888
    ndb = int(a[2]) if len(a) > 2 else 786432
    d = int(a[3]) if len(a) > 3 else 128
    k = int(a[4]) if len(a) > 4 else 20
    reps = int(a[5]) if len(a) > 5 else 3
    dev = torch.device("cuda:0")
    u_s, u_t, _, _ = bench.make_sync_embeddings(ndb, nq, d, dev)
    for _ in range(2):
        out = ops.knn_cosine(u_t, u_s, k, algo=algo)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = ops.knn_cosine(u_t, u_s, k, algo=algo)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("algo=%s nq=%d ndb=%d d=%d k=%d: %.3f ms/build, %.1f TFLOP/s algorithmic, fallback rows %d"
          % (algo, nq, ndb, d, k, ms, 2.0 * nq * ndb * d / ms / 1e9, int(out[3][0])))


if __name__ == "__main__":
    main()
