import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from bridged_gnn_b200 import ops
from bridged_gnn_b200.models import Similar
from oracle import build_oracle as bo
g = dict(np.load("tests/golden/fb_h2c_cosine_build.npz"))
W = {k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("ckpt.")}
head = Similar(64, 2)
head.load_state_dict({k[len("source_learner.sim_net."):]: v for k, v in W.items()})
head.eval()
z_src, z_tar = torch.from_numpy(g["z_src"]), torch.from_numpy(g["z_tar"])
with torch.no_grad():
    u_src_c, u_tar_c = head.cosine_operand(z_src), head.cosine_operand(z_tar)
    head.cuda()
    u_src_g, u_tar_g = head.cosine_operand(z_src.cuda()), head.cosine_operand(z_tar.cuda())
print("u diff gpu-vs-cpu:", float((u_src_g.cpu() - u_src_c).abs().max()), float(u_src_c.abs().max()))
print("tf32 flags:", torch.backends.cuda.matmul.allow_tf32, torch.get_float32_matmul_precision())
full = bo.full_sim_matrix(z_src, z_tar, W, "cosine")
def sim_from_u(us, ut):
    pairs = bo.pair_enumeration(torch.arange(us.shape[0]).unsqueeze(-1), torch.arange(ut.shape[0]).unsqueeze(-1)).t()
    return torch.sigmoid(torch.nn.CosineSimilarity(dim=1)(us[pairs[0]], ut[pairs[1]])).view(-1, us.shape[0])
sim_g = sim_from_u(u_src_g.cpu(), u_tar_g.cpu())
print("oracle(z) vs oracle(gpu u):", float((full - sim_g).abs().max()))
for algo in ("simt", "f16", "tc3", "simt"):
    for inp, nme in (((u_tar_g, u_src_g), "gpu-u"), ((u_tar_c.cuda(), u_src_c.cuda()), "cpu-u")):
        idx, val, gap, st = ops.knn_cosine(inp[0], inp[1], 50, algo=algo)
        ref = sim_g if nme == "gpu-u" else full
        got = torch.gather(ref, 1, idx.cpu())
        d = (got - val.cpu()).abs()
        r = int(d.max(dim=1).values.argmax())
        print(algo, nme, "max |val - oracle[idx]| = %.3e" % float(d.max()), "row", r, "col", int(idx[r][d[r].argmax()]),
              "sorted-val diff %.3e" % float((val.cpu() - bo.canonical_topk(ref, 50)[0]).abs().max()))
