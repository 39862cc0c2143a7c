"""Stress loop over every numeric check of __graft_entry__.smoke() (round-1 open item: one bare AssertionError
in 17 smoke() runs on a fresh box).  The CPU references are computed ONCE; the GPU side of every check is then
repeated `reps` times with allocator churn in between (workspaces land on recycled, NaN-filled memory), each
result compared (a) with the CPU reference under smoke()'s own bound and (b) bit for bit with the first GPU
result of the same check.  Any deviation is printed with the failing tensor's worst entries.

    python tools/stress_smoke.py [reps]      # prints one summary line; exit code 1 on any failure
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bridged_gnn_b200 import ops  # noqa: E402
from oracle import build_oracle as bo  # noqa: E402
from oracle import mp_oracle as mo  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
fails, first, worst = [], {}, {}


def report(name, it, what, got, ref, bound):
    err = (got - ref).abs()
    flat = err.flatten()
    top = torch.topk(flat, min(5, flat.numel()))
    fails.append((name, it, what))
    print("FAIL it=%d check=%s (%s): max err %.3e bound %.3e; worst entries %s got %s ref %s; nan=%d" % (
        it, name, what, float(flat.max()), bound, top.indices.tolist(), got.flatten()[top.indices].tolist(),
        ref.flatten()[top.indices].tolist(), int(torch.isnan(got).sum())), flush=True)


def check(name, it, got, ref, tol, absolute=False):
    got = got.detach().float().cpu()
    bound = tol if absolute else tol * float(ref.abs().max()) + 1e-7
    err = float((got - ref).abs().max()) if got.numel() else 0.0
    if not err <= bound:                      # also catches NaN
        report(name, it, "vs CPU reference", got, ref, bound)
    worst[name] = max(worst.get(name, 0.0), err / bound if bound > 0 else 0.0)
    if name not in first:
        first[name] = got.clone()
    elif not torch.equal(first[name], got):
        report(name, it, "differs bitwise from iteration 0", got, first[name], 0.0)


# ---- references, once ------------------------------------------------------------------------------------
u_s, u_t = torch.randn(5000, 128, generator=g), torch.randn(300, 128, generator=g)
v_ref, i_ref, tie = bo.cosine_knn_rows(u_s, u_t, 20)
n, c = 2000, 64
ei = torch.randint(0, n, (2, 20000), generator=g)
cm = torch.zeros(n, dtype=torch.bool)
cm[:1500] = True
e1, e2, eall = mo.graph_partition(ei, cm)
Hs, Ht = torch.randn(n, c, generator=g), torch.randn(n, c, generator=g)
a1, a2 = torch.randn(c, generator=g) * 0.3, torch.randn(c, generator=g) * 0.3
gout = torch.randn(n, c, generator=g)
refp = [t.clone().requires_grad_(True) for t in (Hs, Ht, a1, a2)]
y_ref = mo.adapted_conv_aggregate(refp[0], refp[1], e1, e2, cm, refp[2], refp[3])
(y_ref * gout).sum().backward()
d = 64
x = torch.randn(n, d, generator=g)
w_cat6, b_cat6 = torch.randn(6, d, generator=g) * 0.3, torch.cat((torch.randn(4, generator=g), torch.zeros(2)))
wd6, kg6 = torch.randn(1, 4, generator=g), torch.randn(2, generator=g)
cm8 = cm.to(torch.uint8)
hs6_ref, ht6_ref = mo.adapted_transform_epilogue(x @ w_cat6.t() + b_cat6, wd6, kg6, cm8)
means_ref = torch.stack((x[:1500].mean(0), x[1500:].mean(0)))
cw = 32
w_catw = torch.randn(2 * cw + 2, d, generator=g) * 0.3
wdw, kgw, b2w = torch.randn(1, 2 * cw, generator=g), torch.randn(2, generator=g), torch.randn(2 * cw, generator=g)
hsw_ref, htw_ref = mo.adapted_transform_epilogue(x.double() @ w_catw.double().t(), wdw.double(), kgw.double(), cm8, b2w.double())
w_l, b_l, go = torch.randn(48, d, generator=g) * 0.2, torch.randn(48, generator=g), torch.randn(n, 48, generator=g)
yl_ref = (x.double() @ w_l.double().t() + b_l.double()).float()
gwl_ref, gxl_ref, gbl_ref = (go.double().t() @ x.double()).float(), (go.double() @ w_l.double()).float(), go.double().sum(0).float()
bn_ref = torch.nn.BatchNorm1d(d).double()
xb_ref = x.double().requires_grad_(True)
yb_ref = torch.relu(bn_ref(xb_ref))
(yb_ref * gout).sum().backward()

# ---- device copies, once ---------------------------------------------------------------------------------
D = lambda t: t.to(dev)
u_s_d, u_t_d = D(u_s), D(u_t)
graph = ops.CSRGraph(D(eall), n)
cm8_d, gout_d, x_d, go_d = D(cm8), D(gout), D(x), D(go)
inv_d = D(torch.tensor([1.0 / 1500, 1.0 / 500]))
t0 = time.time()
for it in range(reps):
    junk = [torch.full((np.random.randint(1, 1 << 22),), float("nan"), device=dev) for _ in range(3)]
    del junk
    for algo in ("f16", "tc3", "simt"):
        idx, val, gap, stats = ops.knn_cosine(u_t_d, u_s_d, 20, algo=algo)
        idx_c = idx.cpu()
        for r in range(300):
            if set(idx_c[r].tolist()) != set(i_ref[r].tolist()) and not bool(tie[r]):
                fails.append(("knn_sets", it, algo))
                print("FAIL it=%d kNN set mismatch row %d (%s): got %s ref %s" % (it, r, algo, sorted(idx_c[r].tolist()),
                      sorted(i_ref[r].tolist())), flush=True)
        check("knn_val", it, val, v_ref, 5e-6, absolute=True)          # all algorithms are bit-identical: one key
        check("knn_idx", it, idx.float(), first.get("knn_idx", idx.float().cpu()), 0.5, absolute=True)
    got = [D(t.clone()).requires_grad_(True) for t in (Hs, Ht, a1, a2)]
    y = ops.gat_aggregate(got[0], got[1], got[2], got[3], graph, cm8_d, 0.1)
    (y * gout_d).sum().backward()
    check("gat_y", it, y, y_ref.detach(), 1e-5)
    for nm, a, b in zip(("gHs", "gHt", "ga1", "ga2"), got, refp):
        check("gat_" + nm, it, a.grad, b.grad, 2e-5)
    hs, ht = ops.adapted_skinny(x_d, D(w_cat6), D(b_cat6), D(wd6), D(kg6), cm8_d)
    check("skinny_hs", it, hs, hs6_ref, 1e-5)
    check("skinny_ht", it, ht, ht6_ref, 1e-5)
    check("means", it, ops.domain_means(x_d, cm8_d, inv_d), means_ref, 1e-5)
    hs, ht = ops.adapted_wide(x_d, D(w_catw), D(b2w), D(wdw), D(kgw), cm8_d)
    check("wide_hs", it, hs, hsw_ref.float(), 1e-5)
    check("wide_ht", it, ht, htw_ref.float(), 1e-5)
    xl, wl, bl = (D(t.clone()).requires_grad_(True) for t in (x, w_l, b_l))
    yl = ops.linear(xl, wl, bl)
    (yl * go_d).sum().backward()
    check("lin_y", it, yl, yl_ref, 1e-5)
    check("lin_gw", it, wl.grad, gwl_ref, 2e-5)
    check("lin_gx", it, xl.grad, gxl_ref, 2e-5)
    check("lin_gb", it, bl.grad, gbl_ref, 2e-5)
    bn = torch.nn.BatchNorm1d(d).to(dev)
    xb = D(x.clone()).requires_grad_(True)
    yb = ops.batch_norm_relu(xb, bn)
    (yb * gout_d).sum().backward()
    check("bn_y", it, yb, yb_ref.detach().float(), 1e-5)
    check("bn_gx", it, xb.grad, xb_ref.grad.float(), 2e-5)
    check("bn_gw", it, bn.weight.grad, bn_ref.weight.grad.float(), 2e-5)
torch.cuda.synchronize()
print("stress_smoke: %d iterations, %d failures, %.1f s; worst err/bound per check: %s" % (
    reps, len(fails), time.time() - t0, {k: round(v, 3) for k, v in sorted(worst.items())}), flush=True)
sys.exit(1 if fails else 0)
