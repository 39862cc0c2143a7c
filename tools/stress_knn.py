"""Determinism / race stress: repeats the same kNN builds many times, interleaving algorithms and sizes,
and reports any run whose output differs bitwise from the first run of that configuration."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bridged_gnn_b200 import ops
from bridged_gnn_b200.models import Similar

g = dict(np.load("tests/golden/fb_h2c_cosine_build.npz"))
W = {k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("ckpt.")}
head = Similar(64, 2)
head.load_state_dict({k[len("source_learner.sim_net."):]: v for k, v in W.items()})
head.eval().cuda()
z_src, z_tar = torch.from_numpy(g["z_src"]).cuda(), torch.from_numpy(g["z_tar"]).cuda()
gen = torch.Generator().manual_seed(1)
cases = {}
with torch.no_grad():
    cases["fb"] = (head.cosine_operand(z_tar), head.cosine_operand(z_src), 50)
cases["rnd"] = (torch.randn(700, 128, generator=gen).cuda(), torch.randn(9000, 128, generator=gen).cuda(), 20)
cases["small"] = (torch.randn(33, 17, generator=gen).cuda(), torch.randn(300, 17, generator=gen).cuda(), 60)
# large enough for the threshold-seeding sweep over a db sample (knn_seed_rows > 0)
cases["seeded"] = (torch.randn(400, 128, generator=gen).cuda(), torch.randn(120000, 128, generator=gen).cuda(), 20)
ref = {}
bad = 0
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for it in range(reps):
    # churn the allocator so workspaces land on recycled, dirty memory
    junk = [torch.full((np.random.randint(1, 1 << 22),), float("nan"), device="cuda") for _ in range(3)]
    del junk
    for name, (q, db, k) in cases.items():
        for algo in ("simt", "f16", "tc3", "tc1"):
            idx, val, gap, st = ops.knn_cosine(q, db, k, algo=algo)
            key = (name, "all")          # every algorithm must agree bit for bit
            cur = (idx.clone(), val.clone(), gap.clone())
            if key not in ref:
                ref[key] = cur
            else:
                for a, b, what in zip(ref[key], cur, ("idx", "val", "gap")):
                    if not torch.equal(a, b):
                        bad += 1
                        d = (a != b).nonzero()
                        print("MISMATCH it=%d case=%s algo=%s %s at %s: %s vs %s" % (it, name, algo, what, d[0].tolist(),
                              a[tuple(d[0])].item(), b[tuple(d[0])].item()), "count", d.shape[0])
print("stress done: %d mismatches over %d reps" % (bad, reps))
