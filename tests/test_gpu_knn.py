"""GPU parity of the bridged-graph build (fused similarity + top-k) against the oracle and the golden
vectors produced by the reference's own Python.  Bar: edge sets exact under the parity key
(fp32 similarity desc, index asc); rows whose k-th / (k+1)-th similarities are closer than 1e-6 are
near-ties and may differ by summation order -- they are listed, not failed."""
import numpy as np
import pytest
import torch

from conftest import sub_state
from oracle import build_oracle as bo

pytestmark = pytest.mark.gpu
T = torch.from_numpy
NEAR_TIE = 1e-6


def _ops():
    from bridged_gnn_b200 import ops
    return ops


def _check_against_full(sim, idx, val, k, tol_val=2e-6, label="", near_tie=NEAR_TIE):
    """sim: oracle [nq, ndb] similarity (CPU).  idx/val: ours.  Sets must match the canonical top-k except
    where the swapped members are within NEAR_TIE of the oracle's k-th value."""
    cv, ci = bo.canonical_topk(sim, k)
    idx, val = idx.cpu(), val.cpu()
    assert idx.shape == ci.shape
    bad, near = [], []
    for r in range(sim.shape[0]):
        a, b = set(ci[r].tolist()), set(idx[r].tolist())
        assert len(b) == k, "duplicate neighbour in row %d" % r
        if a != b:
            diff = list(a ^ b)
            if float((sim[r, diff] - cv[r, -1]).abs().max()) < near_tie:
                near.append(r)
            else:
                bad.append(r)
    assert not bad, "%s: rows with wrong neighbours: %s" % (label, bad[:10])
    # values: ours best-first vs oracle's sorted values
    assert float((val - cv).abs().max()) < tol_val
    # within a row ours is sorted by (value desc, index asc)
    assert bool((val[:, :-1] >= val[:, 1:]).all())
    return near


# ------------------------------------------------------------------ golden fixtures (reference outputs)
def _office_model(g):
    from bridged_gnn_b200.data import Data
    from bridged_gnn_b200.models import Adversarial_Learner_v2
    ns = 2817
    x, y = T(g["x"]), T(g["y"])
    src = Data(x=x[:ns], y=y[:ns], edge_index=torch.zeros((2, 0), dtype=torch.long))
    tar = Data(x=x[ns:], y=y[ns:], edge_index=torch.zeros((2, 0), dtype=torch.long))
    model = Adversarial_Learner_v2(src, tar, dim_hidden=128, num_layer=2, source_clf=True, use_norm=True,
                                   norm_mode="None", norm_scale=1.0, backbone="mlp", sim_mode="mlp")
    res = model.load_state_dict(sub_state(g, "ckpt."), strict=False)
    # the golden file carries only the tensors the build path reads
    assert all(k.startswith(("target_learner.decoder", "discriminator")) for k in res.missing_keys)
    assert not res.unexpected_keys
    return src, tar, model.eval()


def test_office_cross_mlp_head_kernel_parity(office_build):
    """Config 1 (office A->D, v2 mlp head, k_cross=20) at the kernel boundary: same embeddings in, edge
    sets out must equal the reference's except at near-ties."""
    ops = _ops()
    g = office_build
    _, _, model = _office_model(g)
    head = model.source_learner.sim_net.cuda()
    U_db, U_q, w2, b2 = head.mlp_operands(T(g["z_src"]).cuda(), T(g["z_tar"]).cuda())
    idx, val, gap = ops.knn_addrelu(U_q, U_db, w2, b2, 20)
    W = sub_state(g, "ckpt.")
    full = bo.full_sim_matrix(T(g["z_src"]), T(g["z_tar"]), W, "mlp")
    near = _check_against_full(full, idx, val, 20, label="office cross")
    # the reference's own torch.topk differs from the canonical order only at exact ties (13 rows)
    ref_sets = [set(r.tolist()) for r in g["cross_idx"]]
    differ = [r for r in range(591) if ref_sets[r] != set(idx[r].tolist())]
    tie_rows = set(torch.nonzero(bo.near_tie_rows(full, 20, NEAR_TIE)).view(-1).tolist())
    assert set(differ) <= tie_rows, (differ, near)
    assert len(differ) <= 20
    print("near-tie rows (listed, not failed):", sorted(tie_rows))
    # gap output flags the near-tie rows
    flagged = set(torch.nonzero(gap.cpu() < NEAR_TIE).view(-1).tolist())
    strict = set(torch.nonzero(bo.near_tie_rows(full, 20, 0.5 * NEAR_TIE)).view(-1).tolist())
    assert strict <= flagged | set(near)


def test_office_cross_entry_point(office_build):
    """Same config through add_topk_sim_cross_domain_edges (embeddings recomputed on the GPU, so
    similarities carry cuBLAS-vs-MKL rounding of the backbone: near-tie window 1e-5)."""
    from bridged_gnn_b200.main_bridged_graph import add_topk_sim_cross_domain_edges
    g = office_build
    ns = 2817
    src, tar, model = _office_model(g)
    dev = torch.device("cuda:0")
    model.to(dev)
    ei, sim, idx, p_src, p_tar, gap = add_topk_sim_cross_domain_edges(src.to(dev), tar.to(dev), model, epsilon=0.5, k=20,
                                                                      batch_size=1000, return_gap=True, verbose=False)
    assert ei.device.type == "cpu" and ei.dtype == torch.int64 and sim.shape == (591, 20) and idx.shape == (591, 20)
    W = sub_state(g, "ckpt.")
    full = bo.full_sim_matrix(T(g["z_src"]), T(g["z_tar"]), W, "mlp")
    _check_against_full(full, idx, sim, 20, tol_val=2e-5, label="office cross entry", near_tie=1e-5)
    # coalesced (src, tar) edge list, sorted by (src, tar)
    key = ei[0] * 591 + ei[1]
    assert bool((key[1:] > key[:-1]).all()) and ei.shape[1] == 591 * 20
    # known-answer: shipped s->t edges of the bridged graph lie in our top-20 except at ties (10016/10028)
    shipped = set(map(tuple, g["shipped_cross_edges"].T.tolist()))
    mine = set((a, b + ns) for a, b in ei.t().tolist())
    assert len(shipped & mine) >= 10000
    assert float((p_src.cpu() - T(g["probs_clf_src"])).abs().max()) < 1e-5
    assert float((p_tar.cpu() - T(g["probs_clf_tar"])).abs().max()) < 1e-5


def test_office_within_target_matches_reference(office_build):
    g = office_build
    W = sub_state(g, "ckpt.")
    ops = _ops()
    from bridged_gnn_b200.models import Similar_v2
    head = Similar_v2(128, 31, mode="mlp")
    head.load_state_dict({k[len("source_learner.sim_net."):]: v for k, v in W.items() if k.startswith("source_learner.sim_net.")})
    head.eval().cuda()
    z = T(g["z_tar"]).cuda()
    U_db, U_q, w2, b2 = head.mlp_operands(z, z)
    idx, val, gap = ops.knn_addrelu(U_q, U_db, w2, b2, 3)
    full = bo.full_sim_matrix(T(g["z_tar"]), T(g["z_tar"]), W, "mlp")
    _check_against_full(full, idx, val, 3, label="office within-target")
    # edge list identical to the reference's coalesced list wherever there is no near tie
    ei = ops.coalesce(torch.stack((idx.reshape(-1), torch.arange(591, device="cuda").repeat_interleave(3))))
    ref = T(g["within_tar_edge_index"])
    a, b = set(map(tuple, ei.t().tolist())), set(map(tuple, ref.t().tolist()))
    tie_rows = set(torch.nonzero(bo.near_tie_rows(full, 3, NEAR_TIE)).view(-1).tolist())
    assert {e[1] for e in a ^ b} <= tie_rows


@pytest.mark.parametrize("algo", ["simt", "tc3", "tc1", "f16"])
def test_fb_cosine_head_matches_reference(fb_build, algo):
    """Config 3 stand-in: shipped fb_hamilton2caltech v1 cosine head on seeded embeddings, k=50 cross, k=5 within."""
    g = fb_build
    W = sub_state(g, "ckpt.")
    ops = _ops()
    from bridged_gnn_b200.models import Similar
    head = Similar(64, 2)
    head.load_state_dict({k[len("source_learner.sim_net."):]: v for k, v in W.items()})
    head.eval().cuda()
    z_src, z_tar = T(g["z_src"]).cuda(), T(g["z_tar"]).cuda()
    with torch.no_grad():
        u_src, u_tar = head.cosine_operand(z_src), head.cosine_operand(z_tar)
    idx, val, gap, stats = ops.knn_cosine(u_tar, u_src, 50, algo=algo)
    # The node-wise operands u come out of torch's cuBLAS here and out of MKL in the oracle: the two GEMM chains differ
    # in the last bits run to run (first call of a shape in a process, heuristics), and u_src / u_tar of this
    # checkpoint are nearly parallel, which amplifies that.  So (1) the operands must agree to 1e-4 relative, (2) the
    # KERNEL is held to the tight bar against the oracle evaluated on the very operands it was given, and (3) against
    # the reference's own lists the window is the near-tie window of the operand noise.
    with torch.no_grad():
        u_src_c, u_tar_c = head.cpu().cosine_operand(T(g["z_src"])), head.cosine_operand(T(g["z_tar"]))
        head.cuda()
    for a, b in ((u_src, u_src_c), (u_tar, u_tar_c)):
        assert float((a.cpu() - b).abs().max()) <= 1e-4 * float(b.abs().max())

    def sim_from_u(udb, uq):
        pairs = bo.pair_enumeration(torch.arange(udb.shape[0]).unsqueeze(-1), torch.arange(uq.shape[0]).unsqueeze(-1)).t()
        return torch.sigmoid(torch.nn.CosineSimilarity(dim=1)(udb[pairs[0]], uq[pairs[1]])).view(-1, udb.shape[0])
    _check_against_full(sim_from_u(u_src.cpu(), u_tar.cpu()), idx, val, 50, label="fb cross " + algo)
    full = bo.full_sim_matrix(T(g["z_src"]), T(g["z_tar"]), W, "cosine")
    _check_against_full(full, idx, val, 50, tol_val=2e-5, label="fb cross (reference operands) " + algo, near_tie=2e-5)
    ref_sets = [set(r.tolist()) for r in g["cross_idx"]]
    tie_rows = set(torch.nonzero(bo.near_tie_rows(full, 50, 2e-5)).view(-1).tolist())
    differ = {r for r in range(400) if ref_sets[r] != set(idx[r].tolist())}
    assert differ <= tie_rows
    idx, val, gap, _ = ops.knn_cosine(u_tar, u_tar, 5, algo=algo)
    _check_against_full(sim_from_u(u_tar.cpu(), u_tar.cpu()), idx, val, 5, label="fb within " + algo)


# ------------------------------------------------------------------ seeded random inputs vs the oracle
@pytest.mark.parametrize("nq,ndb,d,k", [(1, 1, 3, 1), (5, 7, 128, 7), (130, 1000, 100, 20), (257, 2049, 128, 3),
                                        (64, 4096, 256, 32), (33, 300, 17, 60)])
@pytest.mark.parametrize("algo", ["simt", "tc3", "tc1", "f16"])
def test_cosine_knn_random_vs_oracle(nq, ndb, d, k, algo):
    ops = _ops()
    gq = torch.Generator().manual_seed(1000 + nq + ndb)
    q, db = torch.randn(nq, d, generator=gq), torch.randn(ndb, d, generator=gq)
    idx, val, gap, stats = ops.knn_cosine(q.cuda(), db.cuda(), k, algo=algo)
    pairs = bo.pair_enumeration(torch.arange(ndb).unsqueeze(-1), torch.arange(nq).unsqueeze(-1)).t()
    sim = torch.sigmoid(torch.nn.CosineSimilarity(dim=1)(db[pairs[0]], q[pairs[1]])).view(-1, ndb)
    _check_against_full(sim, idx, val, k, label="random %s" % algo)
    if ndb > k:
        v, _ = torch.sort(sim, dim=1, descending=True, stable=True)
        assert float((gap.cpu() - (v[:, k - 1] - v[:, k])).abs().max()) < 4e-6
    else:
        assert bool(torch.isinf(gap).all())


@pytest.mark.parametrize("algo", ["tc3", "tc1", "f16"])
@pytest.mark.parametrize("nq,ndb,d,k", [(700, 9000, 128, 20), (300, 5000, 96, 50), (129, 70000, 256, 8)])
def test_tensor_core_path_is_bit_identical_to_cuda_core_path(algo, nq, ndb, d, k):
    """The tcgen05 sweep only nominates; after exact re-scoring + certification (+ exact fallback) its
    output must equal the CUDA-core sweep bit for bit -- indices, values and gaps."""
    ops = _ops()
    gq = torch.Generator().manual_seed(7)
    # clustered data: many close similarities around the k-th place stress the certification
    cent = torch.randn(16, d, generator=gq)
    q = (cent[torch.randint(0, 16, (nq,), generator=gq)] + 0.7 * torch.randn(nq, d, generator=gq)).cuda()
    db = (cent[torch.randint(0, 16, (ndb,), generator=gq)] + 0.7 * torch.randn(ndb, d, generator=gq)).cuda()
    i0, v0, g0, _ = ops.knn_cosine(q, db, k, algo="simt")
    i1, v1, g1, st = ops.knn_cosine(q, db, k, algo=algo)
    assert torch.equal(i0, i1)
    assert torch.equal(v0, v1)
    assert torch.equal(g0, g1)
    assert int(st[0]) <= nq


_VARIANT_SCRIPT = r"""
import sys, torch
sys.path.insert(0, %r)
from bridged_gnn_b200 import ops
g = torch.Generator().manual_seed(21)
for nq, ndb, d, k in ((700, 9000, 128, 20), (1000, 70000, 256, 32), (129, 5000, 96, 8)):
    cent = torch.randn(16, d, generator=g)
    q = (cent[torch.randint(0, 16, (nq,), generator=g)] + 0.7 * torch.randn(nq, d, generator=g)).cuda()
    db = (cent[torch.randint(0, 16, (ndb,), generator=g)] + 0.7 * torch.randn(ndb, d, generator=g)).cuda()
    i0, v0, g0, _ = ops.knn_cosine(q, db, k, algo="simt")
    i1, v1, g1, st = ops.knn_cosine(q, db, k, algo="f16")
    assert torch.equal(i0, i1) and torch.equal(v0, v1) and torch.equal(g0, g1), (nq, ndb, d, k)
    print("ok", nq, ndb, d, k, "fallback rows", int(st[0]))
"""


@pytest.mark.parametrize("env", [{"BGNN_F16_PAIR": "0"}, {"BGNN_F16_EW": "4"}, {"BGNN_F16_PAIR": "0", "BGNN_F16_BN": "128"}])
def test_f16_sweep_variants_are_bit_identical(env):
    """The fp16 sweep's shape knobs are read once per process: single CTAs instead of CTA pairs (cta_group::2), four
    epilogue warps / lists per TMEM lane quarter, 128-column tiles.  Every variant must return the CUDA-core result."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", _VARIANT_SCRIPT % root], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 3


@pytest.mark.parametrize("algo", ["tc3", "f16"])
@pytest.mark.parametrize("n_hard", [1, 5, 8, 9, 40])
def test_uncertified_rows_few_and_many_take_the_exact_path(algo, n_hard):
    """Rows whose neighbourhood is a tight cluster (hundreds of db rows within the approximation error of each
    other) cannot be certified by the tensor-core sweep.  Up to 8 of them go through the few-rows exact sweep,
    more through the tiled one; either way the output equals the CUDA-core path bit for bit."""
    ops = _ops()
    g = torch.Generator().manual_seed(11 + n_hard)
    nq, ndb, d, k = 600, 40000, 128, 20
    q = torch.randn(nq, d, generator=g)
    db = torch.randn(ndb, d, generator=g)
    v = torch.randn(d, generator=g)
    db[1000:1400] = v + 1e-4 * torch.randn(400, d, generator=g)          # 400 near-duplicates
    hard = torch.randperm(nq, generator=g)[:n_hard]
    q[hard] = v + 1e-3 * torch.randn(n_hard, d, generator=g)             # queries sitting on the cluster
    q, db = q.cuda(), db.cuda()
    i0, v0, g0, _ = ops.knn_cosine(q, db, k, algo="simt")
    i1, v1, g1, st = ops.knn_cosine(q, db, k, algo=algo)
    assert int(st[0]) >= n_hard                                           # the hard rows were not certified
    assert torch.equal(i0, i1) and torch.equal(v0, v1) and torch.equal(g0, g1)
    if n_hard in (1, 5):      # (a stray random row may join the hard ones: leave room below the limit of 8)
        assert int(st[0]) <= 8, "expected the few-rows path for this case (got %d uncertified rows)" % int(st[0])


def test_exact_ties_resolve_to_lowest_index():
    """Duplicate db rows give exactly equal similarities (office has 52 duplicate source rows);
    the kept members of a tied group must be its lowest indices, i.e. the canonical order exactly."""
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    base = torch.randn(40, 64, generator=g)
    db = base[torch.randint(0, 40, (600,), generator=g)]          # every row repeated ~15 times
    q = torch.randn(50, 64, generator=g)
    pairs = bo.pair_enumeration(torch.arange(600).unsqueeze(-1), torch.arange(50).unsqueeze(-1)).t()
    sim = torch.sigmoid(torch.nn.CosineSimilarity(dim=1)(db[pairs[0]], q[pairs[1]])).view(-1, 600)
    cv, ci = bo.canonical_topk(sim, 10)
    for algo in ("simt", "tc3", "tc1", "f16"):
        idx, val, gap, _ = ops.knn_cosine(q.cuda(), db.cuda(), 10, algo=algo)
        assert torch.equal(idx.cpu(), ci), algo
        assert bool((gap.cpu() >= 0).all())
        assert int((gap.cpu() == 0).sum()) > 0      # some rows cut a tied group at the k-th place


def test_within_domain_keeps_self_and_aliasing():
    ops = _ops()
    x = torch.randn(500, 128, generator=torch.Generator().manual_seed(5)).cuda()
    for algo in ("simt", "tc3", "f16"):
        idx, val, _, _ = ops.knn_cosine(x, x, 4, algo=algo)
        assert torch.equal(idx[:, 0], torch.arange(500, device="cuda"))       # self is rank 1, not excluded
        idx2, val2, _, _ = ops.knn_cosine(x, x.clone(), 4, algo=algo)        # non-aliased path, same result
        assert torch.equal(idx, idx2) and torch.equal(val, val2)


def test_k_equals_ndb_and_invalid_k():
    ops = _ops()
    q, db = torch.randn(9, 32).cuda(), torch.randn(6, 32).cuda()
    idx, val, gap, _ = ops.knn_cosine(q, db, 6, algo="simt")
    assert sorted(idx[0].tolist()) == list(range(6)) and bool(torch.isinf(gap).all())
    with pytest.raises(ValueError):
        ops.knn_cosine(q, db, 7)
    with pytest.raises(RuntimeError):
        ops.knn_cosine(q.cpu(), db.cpu(), 2)


def test_zero_rows_use_cosine_eps():
    """A zero vector has cos = 0 with everything (denominator clamps at 1e-8): sim = sigmoid(0) = 0.5."""
    ops = _ops()
    db = torch.randn(100, 16).cuda()
    q = torch.zeros(3, 16).cuda()
    idx, val, _, _ = ops.knn_cosine(q, db, 5, algo="simt")
    assert bool((val == 0.5).all()) and idx[0].tolist() == [0, 1, 2, 3, 4]


@pytest.mark.parametrize("nq,ndb,h,k", [(100, 1000, 128, 20), (65, 129, 33, 3), (7, 5000, 64, 50)])
def test_addrelu_knn_random_vs_oracle(nq, ndb, h, k):
    ops = _ops()
    g = torch.Generator().manual_seed(11)
    Uq, Udb = torch.randn(nq, h, generator=g), torch.randn(ndb, h, generator=g)
    w2, b2 = torch.randn(h, generator=g) * 0.3, 0.1
    idx, val, gap = ops.knn_addrelu(Uq.cuda(), Udb.cuda(), w2.cuda(), b2, k)
    sim = torch.sigmoid((torch.relu(Uq[:, None, :] + Udb[None, :, :]) * w2).sum(-1) + b2)
    _check_against_full(sim, idx, val, k, tol_val=4e-6, label="addrelu")


# ------------------------------------------------------------------ full-size properties (configs 4 / 5 shapes)
def test_sync_1m_shape_sampled_rows_exact():
    """Config 4 shape: db = 786 432 source rows, d=128, k=20.  4096 sampled target rows through the
    tensor-core path must equal the exact CUDA-core path bit for bit; 64 of them are also checked
    against the pair-materialising oracle on the CPU."""
    ops = _ops()
    from bench import make_sync_embeddings
    u_src, u_tar, _, _ = make_sync_embeddings(786432, 4096, 128, torch.device("cuda:0"), seed=0)
    i0, v0, g0, _ = ops.knn_cosine(u_tar, u_src, 20, algo="simt")
    for algo in ("f16", "tc3"):
        i1, v1, g1, st = ops.knn_cosine(u_tar, u_src, 20, algo=algo)
        assert torch.equal(i0, i1) and torch.equal(v0, v1) and torch.equal(g0, g1), algo
        print(algo, "exact-fallback rows:", int(st[0]), "of 4096")
        assert int(st[0]) < 4096 // 4          # certification must hold for the bulk of the rows
    rows = torch.arange(0, 4096, 64)
    v, i, tie = bo.cosine_knn_rows(u_src.cpu(), u_tar.cpu(), 20, rows=rows, chunk=4)
    for n, r in enumerate(rows.tolist()):
        if set(i[n].tolist()) != set(i1[r].tolist()):
            assert bool(tie[n]), r
    assert float((v - v1.cpu()[rows]).abs().max()) < 2e-6


def test_sync_16m_shape_sampled_rows_exact():
    """Config 5 shape: db = 12 582 912 source rows, d=256, k=32 (12.9 GB of fp32 embeddings on one GPU).
    2048 sampled target rows: tensor-core path == exact CUDA-core path bit for bit."""
    ops = _ops()
    from bench import make_sync_embeddings
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * (1 << 30):
        pytest.skip("needs ~45 GB of free HBM")
    u_src, u_tar, _, _ = make_sync_embeddings(12582912, 2048, 256, torch.device("cuda:0"), seed=0)
    i1, v1, g1, st = ops.knn_cosine(u_tar, u_src, 32, algo="f16")
    i0, v0, g0, _ = ops.knn_cosine(u_tar, u_src, 32, algo="simt")
    assert torch.equal(i0, i1) and torch.equal(v0, v1) and torch.equal(g0, g1)
    assert int(st[0]) < 256
    assert bool((i1 >= 0).all()) and int(i1.max()) < 12582912
    assert bool((v1[:, :-1] >= v1[:, 1:]).all())
