"""Accuracy of the weight-gradient contraction at training size against float64: this library's 3 x TF32 tensor-core
kernel next to torch's fp32 GEMM (cuBLAS SIMT), error relative to max |ref| and to sum |g||x|."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bridged_gnn_b200 import ops  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator(device="cuda").manual_seed(0)
for n, no, d, shift in ((1 << 20, 64, 64, 0.0), (1 << 20, 130, 128, 0.0), (1 << 20, 64, 64, 0.5), (1 << 17, 64, 64, 0.0)):
    G = (torch.randn(n, no + (-no) % 4, device="cuda", generator=g) * 1e-3)[:, :no]
    X = torch.randn(n, d, device="cuda", generator=g) + shift
    ref = G.double().t() @ X.double()
    scale = G.double().abs().t() @ X.double().abs()
    for name, W in (("tc 3xTF32", ops.wgrad_gemm(G, X)), ("torch fp32", G.t() @ X)):
        e = (W.double() - ref).abs()
        print("n=%d no=%d d=%d shift=%.1f  %-10s  max err / max|ref| = %.2e   max err / sum|g||x| = %.2e" % (
            n, no, d, shift, name, float(e.max() / ref.abs().max()), float((e / scale).max())))
