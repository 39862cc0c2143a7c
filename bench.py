#!/usr/bin/env python
"""Benchmark of the Bridged-GNN hot path on B200 (BASELINE.json metric:
"bridged-graph kNN build ms & KT-GNN message-passing GEdges/s").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[3]): synthetic Sync-RD_intra-and-inter, 2^20 nodes per GPU
(786 432 source + 262 144 target), dim 128, 2 classes, k_cross = 20, 70 % homophily random edges.
  phase A  bridged-graph build: cosine-head kNN of every target row against all source rows
           (fused tcgen05 similarity + top-k, exact after re-scoring) + edge assembly      -> knn_build.ms
  phase B  message passing: one KT-GNN (2 layers, hidden 64) training forward + backward over
           the bridged graph (random edges U kNN edges, undirected, self loops)             -> value, GEdges/s
A "step" is one phase-B pass; `value` = E_mp * 8 conv passes (4 forward + 4 backward) / step time.
Multi-GPU (torchrun): the headline line is weak scaling -- every rank owns its own 2^20-node replica and `value` sums
the ranks' edges over the max time.  Beside it, at N > 1 the line carries the paths that really shard:
  `sharded_build_1m`  ONE sync-1M build with the target rows split over the ranks + NCCL all-gather (verified bit for
                      bit against the 1-GPU lists)
  `partitioned_1m`    ONE sync-1M KT-GNN step with the destination rows partitioned over the ranks (domain-aware halo
                      exchange of H, rank-combined BatchNorm), against the 1-GPU step of the same graph
  `sync16m`           BASELINE configs[4]: 2^24 nodes, dim 256, k_cross 32 -- row-sharded build, partitioned KT-GNN
                      step and the panel-pipelined partitioned SpMM (F = 256); also run at N = 1 when it fits
At N = 1 the line also carries `spmm` (K2 at F = 64 / 128 / 256), `addrelu` (K1b) and `configs_1_3` (office / fb).

--impl reference times the oracle restatement of the reference's own op sequence on the host CPU
(PyG-free port; the reference itself needs torch_geometric, which is not installable here) on a
bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NS, NT, DIM, K_CROSS, N_CLASS, HIDDEN = 786432, 262144, 128, 20, 2, 64
RAND_EDGES_PER_NODE, HOMOPHILY = 5, 0.7
METRIC = "bridged-graph kNN build ms & KT-GNN message-passing GEdges/s"
# identical in both arms (the driver compares it): only what names the workload
CONFIG = {"workload": "sync-1M per GPU (BASELINE configs[3]): Ns=%d Nt=%d dim=%d k_cross=%d classes=%d hidden=%d; "
                      "step = KT-GNN train fwd+bwd (4 AdaptedConv fwd + 4 bwd) over the bridged graph (5N random 70%%-homophily "
                      "edges U kNN edges, undirected, self loops)" % (NS, NT, DIM, K_CROSS, N_CLASS, HIDDEN),
          "conv_passes_per_step": 8,
          "l2": "inputs exceed L2 (features 537 MB, db 403 MB of fp32)"}
# BASELINE configs[4]
NS16, NT16, DIM16, K16 = 12582912, 4194304, 256, 32


# ----------------------------------------------------------------------------- synthetic workload
def make_sync_embeddings(ns, nt, d, device, seed=0, n_class=N_CLASS, tar_seed=None):
    """SURVEY 8d config 4 generator: class means mu_s[c] ~ N(0, I), mu_t[c] = 0.8 mu_s[c] + 0.3 + 0.2 N(0, I);
    u = mu[y] + N(0, I).  Means come from `seed`, source rows from seed+1, target rows from `tar_seed`
    (default seed+2) so that shards of target rows share one source set.
    Returns (u_src [ns,d], u_tar [nt,d], y_src, y_tar)."""
    g = torch.Generator(device=device).manual_seed(seed)
    mu_s = torch.randn(n_class, d, generator=g, device=device)
    mu_t = 0.8 * mu_s + 0.3 + 0.2 * torch.randn(n_class, d, generator=g, device=device)
    gs = torch.Generator(device=device).manual_seed(seed + 1)
    y_s = torch.randint(0, n_class, (ns,), generator=gs, device=device)
    u_s = mu_s[y_s] + torch.randn(ns, d, generator=gs, device=device)
    gt = torch.Generator(device=device).manual_seed(seed + 2 if tar_seed is None else tar_seed)
    y_t = torch.randint(0, n_class, (nt,), generator=gt, device=device)
    u_t = mu_t[y_t] + torch.randn(nt, d, generator=gt, device=device)
    return u_s, u_t, y_s, y_t


def make_random_edges(y, per_node, homophily, device, seed=1):
    """per_node * N directed draws: dst uniform, src of the same class w.p. `homophily` else of another class."""
    n = y.shape[0]
    g = torch.Generator(device=device).manual_seed(seed)
    e = per_node * n
    dst = torch.randint(0, n, (e,), generator=g, device=device)
    same = torch.rand(e, generator=g, device=device) < homophily
    order = torch.argsort(y, stable=True)
    counts = torch.bincount(y, minlength=N_CLASS)
    starts = torch.cumsum(counts, 0) - counts
    cls = torch.where(same, y[dst], (y[dst] + 1) % N_CLASS)
    r = torch.rand(e, generator=g, device=device)
    pos = starts[cls] + (r * counts[cls]).long().clamp(max=n - 1)
    src = order[pos.clamp(max=n - 1)]
    return torch.stack((src, dst), 0)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                p = [t.strip() for t in line.split(",")]
                if len(p) < 9:
                    continue
                sm.append(float(p[1])); mx.append(float(p[2]))
                for nme, v in zip(names, p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            os.unlink(self.path)
        except Exception:
            pass
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


def load_traffic():
    """DRAM bytes (read + write) per launch of the dominant kernels, from `ncu --set full` captures of this very
    workload (tools/ncu_traffic.py -> profiles/traffic.json); {} when no capture has been summarised yet."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_knn_sample(rows=12, seed=0):
    """Oracle (reference op order: pair enumeration -> gather -> CosineSimilarity -> sigmoid -> top-k) on a
    few target rows against the full 786 432-row source set."""
    from oracle import build_oracle as bo
    cores = bo.set_threads()
    dev = torch.device("cpu")
    u_s, u_t, _, _ = make_sync_embeddings(NS, rows, DIM, dev, seed)
    t0 = time.perf_counter()
    bo.cosine_knn_rows(u_s, u_t, K_CROSS, chunk=4)
    dt = time.perf_counter() - t0
    return {"gpairs_per_s": rows * NS / dt / 1e9, "seconds": dt, "cores": cores,
            "sample": "%d target rows x %d source rows, d=%d (full build = %d rows)" % (rows, NS, DIM, NT),
            "ms_full_build_extrapolated": dt / rows * NT * 1e3}


def _oracle_ktgnn_params(model):
    return {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}


def cpu_mp_sample(n_nodes=1 << 17, seed=0, reps=1):
    """Oracle KT-GNN (PyG-semantics gather / softmax / scatter-add, torch autograd) forward + backward on
    the same generator at a reduced node count."""
    from oracle import build_oracle as bo
    from oracle import mp_oracle as mo
    cores = bo.set_threads()
    dev = torch.device("cpu")
    ns, nt = n_nodes * 3 // 4, n_nodes // 4
    u_s, u_t, y_s, y_t = make_sync_embeddings(ns, nt, DIM, dev, seed)
    x = torch.cat((u_s, u_t), 0)
    y = torch.cat((y_s, y_t), 0)
    ei = make_random_edges(y, RAND_EDGES_PER_NODE, HOMOPHILY, dev)
    # stand-in for the kNN edges at this size: K_CROSS random same-class source neighbours per target row
    tar = torch.arange(ns, n_nodes).repeat_interleave(K_CROSS)
    srcn = torch.randint(0, ns, (tar.numel(),), generator=torch.Generator().manual_seed(2))
    ei = mo.to_undirected(torch.cat((ei, torch.stack((srcn, tar))), 1), n_nodes)
    cm = torch.zeros(n_nodes, dtype=torch.bool)
    cm[:ns] = True
    torch.manual_seed(0)
    P = _init_ktgnn_params(DIM, N_CLASS, HIDDEN)
    e1, e2, eall = mo.graph_partition(ei, cm)
    e_mp = eall.shape[1]
    best = None
    for _ in range(reps):
        Pg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in P.items()}
        t0 = time.perf_counter()
        lb, lt, ltt = mo.ktgnn_no_complement(x, ei, cm, Pg, training=True)
        loss = sum(torch.nn.functional.nll_loss(l[cm], y[cm]) for l in (lb, lt, ltt))
        loss.backward()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": e_mp * 8 / best / 1e9, "unit": "GEdges/s", "cores": cores, "kind": "port", "seconds": best,
            "sample": "same generator at N=%d nodes (E_mp=%d); oracle KT-GNN fwd+bwd, best of %d" % (n_nodes, e_mp, reps)}


def _init_ktgnn_params(f_in, n_class, hidden):
    """state_dict-shaped random parameters for the oracle (shapes of KTGNN_no_complement, layer_num=2)."""
    g = torch.Generator().manual_seed(0)
    P = {}

    def conv(prefix, i, o):
        for nme, shape in (("lin_s.weight", (o, i)), ("lin_s.bias", (o,)), ("lin_t.weight", (o, i)), ("lin_t.bias", (o,)),
                           ("a_g_s2t.weight", (1, 2 * i)), ("a_g_t2s.weight", (1, 2 * i)), ("a_f_s2t.weight", (1, o)),
                           ("a_f_t2s.weight", (1, o))):
            P[prefix + nme] = torch.randn(shape, generator=g) * (1.0 / max(shape[-1], 1)) ** 0.5
    conv("convs.0.", f_in, hidden)
    conv("clf_base.", hidden, n_class)
    conv("clf_target.", hidden, n_class)
    for bn in ("bns.0", "clf_transformer.1"):
        P[bn + ".weight"], P[bn + ".bias"] = torch.ones(hidden), torch.zeros(hidden)
        P[bn + ".running_mean"], P[bn + ".running_var"] = torch.zeros(hidden), torch.ones(hidden)
    for lin in ("clf_transformer.0", "clf_transformer.3"):
        P[lin + ".weight"] = torch.randn(hidden, hidden, generator=g) / hidden ** 0.5
        P[lin + ".bias"] = torch.zeros(hidden)
    return P


# ----------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's CPU path (oracle port: torch_geometric is not installable) on the box's host cores.
    Every step is a BOUNDED SAMPLE of the workload: the same generator at N = 2^18 nodes (a quarter of the nodes, same
    degree distribution), one KT-GNN training forward + backward; `value` is per-edge throughput, so it is comparable
    with the GPU arm's.  K steps are really run (after W warm-up steps)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.perf_counter()
    n_nodes = 1 << 18
    K, W = max(1, args.steps), max(0, args.warmup)
    # keep the whole run within a few minutes: a step takes ~2.5 s on 16 cores
    budget_steps = 60
    if K + W > budget_steps:
        W = min(W, 3)
        K = max(1, budget_steps - W)
    runs = [cpu_mp_sample(n_nodes=n_nodes, reps=1) for _ in range(W + K)]
    timed = runs[W:]
    secs = sum(r["seconds"] for r in timed) / len(timed)
    e_mp = timed[0]["value"] * timed[0]["seconds"] * 1e9 / 8
    value = e_mp * 8 / secs / 1e9
    knn = cpu_knn_sample(rows=48)
    sample = "%d timed steps (+%d warm-up) of: %s" % (len(timed), W, timed[0]["sample"].replace(", best of 1", ""))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GEdges/s", "n_gpus": args.gpus,
        "steps": len(timed), "warmup": W, "ms_per_step": secs * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": CONFIG,
        "cpu_baseline": {"value": value, "unit": "GEdges/s", "cores": timed[0]["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "GEdges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "knn_build": {"ms": knn["ms_full_build_extrapolated"], "gpairs_per_s": knn["gpairs_per_s"], "cores": knn["cores"],
                      "sample": knn["sample"], "note": "ms extrapolated linearly in rows from the sample"},
        "requested": {"steps": args.steps, "warmup": args.warmup}, "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm
class Ctx:
    """Process-wide state of the GPU arm: device, ranks, timing helpers."""

    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU oracle)")
        self.dev = torch.device("cuda", self.local)
        torch.cuda.set_device(self.dev)
        if self.world > 1:
            import datetime
            # a rank that fails must not leave the others waiting in a collective for NCCL's default 10 minutes
            dist.init_process_group("nccl", device_id=self.dev, timeout=datetime.timedelta(seconds=300))
        self.K, self.W = args.steps, max(args.warmup, 3)
        self.args = args

    def hp_group(self):
        """All ranks on a communicator whose NCCL kernels run on a HIGH-PRIORITY stream: the exchanges of the partitioned
        paths then make progress while a gather kernel fills the SMs (tools/dist_spmm_check.py: the column-panel
        pipeline overlaps only with it)."""
        if self.world == 1:
            return None
        if getattr(self, "_hp", None) is None:
            opts = self.dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            self._hp = self.dist.new_group(pg_options=opts)
        return self._hp

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = torch.tensor([v], device=self.dev, dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        if self.world == 1:
            return v
        t = torch.tensor([v], device=self.dev, dtype=torch.float64)
        self.dist.all_reduce(t)
        return float(t.item())

    def timed(self, fn, steps, warm):
        """ms per call: `warm` untimed calls, then `steps` calls between barrier + synchronize, CUDA events on the
        launching stream, max over ranks."""
        for _ in range(warm):
            fn()
        self.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        self.barrier()
        return self.max_over_ranks(a.elapsed_time(b)) / steps

    def timed_once(self, fn):
        """(ms, result) of ONE call (for calls that take seconds: the caller warms the kernels up on a slice)."""
        self.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        self.barrier()
        return self.max_over_ranks(a.elapsed_time(b)), out


def nll_weighted(y, mask):
    """nll_loss(out[mask], y[mask]) (main_graph_knowledge_transfer.py:57-59) as gather * mask / count: no boolean-mask
    compaction and none of ATen's single-block nll_loss reductions (0.5 ms each at 1 M rows)."""
    w = mask.to(torch.float32) / mask.sum()
    y_col = y.unsqueeze(1)

    def nll(lp):
        return -(lp.gather(1, y_col).squeeze(1) * w).sum()
    return nll


def gat_bytes(e_mp, n, c, bwd):
    """Algorithmic bytes of one launch of the fused AdaptedConv aggregation."""
    fwd_b = e_mp * (4 + 4 * c) + n * (4 + 4 * c + 4 * c + 1 + 8)
    rec = 16 if c <= 64 else 8 + 4 * ((c + 31) // 32)
    # pass A: col + H[src] gather + record write; pass B: t_col + slot map + record + gout[dst] gather; 7 row-sized node passes
    bwd_b = e_mp * (4 + 4 * c + rec) + e_mp * (8 + rec + 4 * c) + n * 7 * 4 * c
    return bwd_b if bwd else fwd_b


def measure_spmm(cx, graph, n, peaks):
    """K2: CSR SpMM mean, forward and backward (= the same kernel on the transposed CSR), on the bench graph
    at F = 64 / 128 / 256.  Two byte models: the per-edge one, E (4 + 4F) + N (4 + 4F) -- the contract for graphs whose X
    does not fit L2 -- and the compulsory one, 8 N F + 4 E, reached only if every row of X were fetched once."""
    from bridged_gnn_b200 import ops
    out = {}
    e = graph.e
    _ = graph.t
    for f in (64, 128, 256):
        x = torch.randn(n, f, device=cx.dev, requires_grad=True)
        gy = torch.randn(n, f, device=cx.dev)
        fwd = cx.timed(lambda: ops.spmm(graph, x.detach(), "mean"), 5, 3)
        y = ops.spmm(graph, x, "mean")

        def bwd_only():
            x.grad = None
            y.backward(gy, retain_graph=True)
        bwd = cx.timed(bwd_only, 5, 3)
        per_edge = e * (4 + 4 * f) + n * (4 + 4 * f)
        comp = 8 * n * f + 4 * e
        out["F=%d" % f] = {"fwd_ms": fwd, "bwd_ms": bwd, "gedges_per_s_fwd": e / fwd / 1e6,
                           "per_edge_model": {"bytes": per_edge, "fwd_gbs": per_edge / fwd / 1e6, "fwd_frac": per_edge / fwd / 1e6 / peaks["hbm_gbs"],
                                              "bwd_gbs": per_edge / bwd / 1e6, "bwd_frac": per_edge / bwd / 1e6 / peaks["hbm_gbs"]},
                           "compulsory_model": {"bytes": comp, "fwd_gbs": comp / fwd / 1e6, "fwd_frac": comp / fwd / 1e6 / peaks["hbm_gbs"]}}
        del x, gy, y
    out["note"] = ("X of the bench graph is 268 / 537 / 1074 MB against 126 MB of L2: part of the gather is served by L2, so the "
                   "per-edge model can exceed the DRAM peak; peak = %.0f GB/s (%s)" % (peaks["hbm_gbs"], peaks["source"]))
    return out


def measure_addrelu(cx):
    """K1b: the add-ReLU head on CUDA cores: 3 Nq Ndb H lane-ops against 148 SM x 128 lanes x f_SM."""
    from bridged_gnn_b200 import ops
    g = torch.Generator(device=cx.dev).manual_seed(3)
    h, k = 128, 20
    w2 = torch.randn(h, generator=g, device=cx.dev) * 0.1
    res = {}
    for name, nq, ndb, steps in (("config1_cross_591x2817", 591, 2817, 20), ("sync1m_quarter_65536x786432", 65536, NS, 1)):
        uq, udb = torch.randn(nq, h, generator=g, device=cx.dev), torch.randn(ndb, h, generator=g, device=cx.dev)
        if steps == 1:
            ops.knn_addrelu(uq[:4096], udb, w2, 0.1, k)                      # warm-up on a slice
            ms = cx.timed(lambda: ops.knn_addrelu(uq, udb, w2, 0.1, k), 1, 0)
        else:
            ms = cx.timed(lambda: ops.knn_addrelu(uq, udb, w2, 0.1, k), steps, 3)
        lane_ops = 3.0 * nq * ndb * h
        res[name] = {"ms": ms, "gpairs_per_s": nq * ndb / ms / 1e6, "lane_tops": lane_ops / ms / 1e9,
                     "frac_of_37.2T_lane_ops": lane_ops / ms / 1e9 / 37.2}
        del uq, udb
    res["note"] = "ceiling 148 SM x 128 lanes x 1.965 GHz = 37.2 T lane-ops/s (add, max, fma per pair and hidden unit)"
    return res


def measure_small_configs(cx):
    """Wall time of BASELINE configs[0..2] (launch-latency bound; parity is in tests/): office A->D build through the
    reference entry points, one office KT-GNN epoch eager and as CUDA graphs, fb-shaped build + steps."""
    import numpy as np
    from bridged_gnn_b200 import ops
    from bridged_gnn_b200.data import Data, to_undirected
    from bridged_gnn_b200.main_bridged_graph import add_topk_sim_cross_domain_edges, add_topk_sim_within_domain_edges
    from bridged_gnn_b200.main_graph_knowledge_transfer import GraphedEpoch, get_each_clf_res, test, train
    from bridged_gnn_b200.models import Adversarial_Learner_v2, GraphSAGE, KTGNN_no_complement
    dev, out = cx.dev, {}
    path = os.path.join(ROOT, "tests", "golden", "office_a2d_build.npz")
    if os.path.exists(path):
        g = np.load(path)
        T = torch.from_numpy
        ns = 2817
        x, y, cm = T(g["x"]).to(dev), T(g["y"]).to(dev), T(g["central_mask"]).to(dev)
        src = Data(x=x[:ns].contiguous(), y=y[:ns], edge_index=torch.zeros((2, 0), dtype=torch.long, device=dev))
        tar = Data(x=x[ns:].contiguous(), y=y[ns:], edge_index=torch.zeros((2, 0), dtype=torch.long, device=dev))
        sim = Adversarial_Learner_v2(src, tar, dim_hidden=128, num_layer=2, source_clf=True, use_norm=True, norm_mode="None",
                                     norm_scale=1.0, backbone="mlp", sim_mode="mlp")
        sim.load_state_dict({k[5:]: T(g[k]) for k in g.files if k.startswith("ckpt.")}, strict=False)
        sim = sim.to(dev).eval()

        def build():
            add_topk_sim_cross_domain_edges(src, tar, sim, k=20, verbose=False)
            add_topk_sim_within_domain_edges(src, sim, k=3, domain="source", verbose=False)
            add_topk_sim_within_domain_edges(tar, sim, k=3, domain="target", verbose=False)
        out["config1_office_build"] = {"ms": cx.timed(build, 5, 3), "what": "cross k=20 (591 x 2817) + within-source k=3 (2817^2) + "
                                       "within-target k=3 (591^2), embeddings, mlp head, coalesce and the copies to the host included",
                                       "reference_cpu_s": "4.7-5.6 (cross) + 19.2 (within-source), SURVEY 6"}
        ei = to_undirected(T(g["edge_index"]).to(dev), x.shape[0])
        data = Data(x=x, edge_index=ei, y=y, central_mask=cm, train_mask=T(g["train_mask"]).to(dev), val_mask=T(g["val_mask"]).to(dev),
                    test_mask=T(g["test_mask"]).to(dev))
        torch.manual_seed(0)
        model = KTGNN_no_complement(256, 31, 2, 64, root_weight=False, use_bn=True, dim_share=256).to(dev)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-3)

        def epoch_eager():
            train(data, model, opt, gnn="KTGNN")
            test(data, model, gnn="KTGNN")
            get_each_clf_res(data, model)
        eager = cx.timed(epoch_eager, 10, 3)
        ge = GraphedEpoch(data, model, opt)

        def epoch_graphed():
            ge.train_step()
            ge.evaluate()
        graphed = cx.timed(epoch_graphed, 10, 3)
        out["config2_office_ktgnn_epoch"] = {"eager_ms": eager, "cuda_graph_ms": graphed, "E_mp": 37522,
                                             "what": "train step + test + get_each_clf_res (main_graph_knowledge_transfer.py:219-227), "
                                                     "device-side F1 included; the reference's AdaptedConv forward alone takes ~110 ms on 8 CPU cores"}
    # config 3: fb_hamilton2caltech shapes (data not shipped: seeded stand-ins), cosine head d = 128, k = 50
    g3 = torch.Generator(device=dev).manual_seed(0)
    ns, nt, din = 2314, 769, 1685
    u_s, u_t = torch.randn(ns, 128, generator=g3, device=dev), torch.randn(nt, 128, generator=g3, device=dev)
    out["config3_fb_build_kernel"] = {"ms": cx.timed(lambda: ops.knn_cosine(u_t, u_s, 50), 20, 3), "what": "769 x 2314, d=128, k=50 (CUDA-core sweep: too small for tcgen05)"}
    n = ns + nt
    xf = torch.zeros(n, din, device=dev)
    xf.scatter_(1, torch.randint(0, din, (n, 6), generator=g3, device=dev), 1.0)
    ei = to_undirected(torch.randint(0, n, (2, 113000), generator=g3, device=dev), n)
    y3 = torch.randint(0, 2, (n,), generator=g3, device=dev)
    cm3 = torch.arange(n, device=dev) < ns
    d3 = Data(x=xf, edge_index=ei, y=y3, central_mask=cm3, train_mask=torch.ones(n, dtype=torch.bool, device=dev))
    torch.manual_seed(0)
    kt = KTGNN_no_complement(din, 2, 2, 64, root_weight=False, use_bn=True, dim_share=din).to(dev)
    sg = GraphSAGE(type("D", (), {"num_features": din, "num_classes": 2})(), layer_num=2, hidden=64).to(dev)
    o1, o2 = torch.optim.Adam(kt.parameters(), lr=1e-3), torch.optim.Adam(sg.parameters(), lr=1e-3)
    out["config3_fb_train_step"] = {"ktgnn_ms": cx.timed(lambda: train(d3, kt, o1, gnn="KTGNN"), 10, 3),
                                    "graphsage_no_dtc_ms": cx.timed(lambda: train(d3, sg, o2, gnn="GraphSAGE"), 10, 3),
                                    "what": "N=3083, F=1685 one-hot-like, E=%d undirected; optimiser step included" % ei.shape[1]}
    return out


def build_bridged_graph(idx, y, ns, n, dev, seed):
    """Bridged graph of the synthetic configs: 5N random 70 %-homophily edges U kNN (source, target) edges, undirected."""
    from bridged_gnn_b200 import dist as bdist
    from bridged_gnn_b200.data import to_undirected
    rnd = make_random_edges(y, RAND_EDGES_PER_NODE, HOMOPHILY, dev, seed=seed)
    cross = bdist.edges_from_topk(idx) + torch.tensor([[0], [ns]], device=dev)
    return to_undirected(torch.cat((rnd, cross), 1), n)


def partitioned_step_fn(cx, model, x_full_fn, ei_u, cm, y):
    """(train_step, data_loc, part) of the destination-partitioned KT-GNN on this rank's rows.  The row blocks are cut
    at equal incoming-EDGE counts (target rows carry ~1.8x the edges of source rows in these graphs)."""
    from bridged_gnn_b200 import dist as bdist
    from bridged_gnn_b200.data import Data
    from bridged_gnn_b200.models import graph_partition
    _, _, ei_all = graph_partition(ei_u, cm)
    n = cm.shape[0]
    n_src_nodes = int(cm.sum())
    prefix = n_src_nodes if bool(cm[:n_src_nodes].all()) else None
    part = bdist.DstPartition(n, cx.hp_group(), bounds=bdist.DstPartition.balanced_bounds(ei_all[1], n, cx.world, n_prefix=prefix))
    ei_loc = part.local_edges(ei_all)
    del ei_all
    data_loc = Data(x=x_full_fn(part.r0, part.r1), edge_index=ei_loc, central_mask=part.pad_rows(cm), part=part)
    tm_loc, y_loc = part.local_rows(cm), part.local_rows(y)          # train mask = the source nodes, as in the headline
    w = tm_loc.to(torch.float32) / cm.sum()
    y_col = y_loc.unsqueeze(1)

    def nll(lp):
        return -(lp.gather(1, y_col).squeeze(1) * w).sum()

    def train_step():
        model.zero_grad(set_to_none=True)
        lb, lt, ltt, _ = model(data_loc)
        loss = nll(lb) + nll(lt) + nll(ltt)
        loss.backward()
        part.sync_grads(model)
        return loss
    return train_step, data_loc, part


def run_sync16m(cx, peaks):
    """BASELINE configs[4]: N = 2^24 nodes (12 582 912 source + 4 194 304 target), dim 256, k_cross 32.
    Row-sharded build (source set replicated, target rows split, lists all-gathered), then message passing on the
    destination-partitioned bridged graph: one KT-GNN training step (F_in 256, hidden 64) and the SAGE-style mean SpMM at
    F = 256 with the column-panel pipelined halo.  Strong scaling: the problem is fixed, N ranks share it."""
    from bridged_gnn_b200 import dist as bdist
    from bridged_gnn_b200 import ops
    from bridged_gnn_b200.models import KTGNN_no_complement
    dev, world, rank = cx.dev, cx.world, cx.rank
    res = {"workload": "sync-16M (BASELINE configs[4]): Ns=%d Nt=%d dim=%d k_cross=%d" % (NS16, NT16, DIM16, K16), "scaling": "strong"}
    free, total = torch.cuda.mem_get_info()
    res["hbm_free_gb_at_start"] = round(free / 2**30, 1)
    if free < 120 * 2**30:
        res["skipped"] = "needs ~110 GB of free HBM per GPU"
        return res
    u_src, u_tar, y_src, y_tar = make_sync_embeddings(NS16, NT16, DIM16, dev, seed=0)
    s, e = bdist.row_shard(NT16, rank, world)
    # ---- build -------------------------------------------------------------------------------------------------
    def build():
        idx, val, gap, stats = ops.knn_cosine(u_tar[s:e], u_src, K16, algo="f16")
        if world > 1:
            idx = bdist.all_gather_rows(idx, NT16)
            val = bdist.all_gather_rows(val, NT16)
        return idx, val, stats
    ops.knn_cosine(u_tar[s:s + 8192], u_src, K16, algo="f16")            # warm-up on a slice (the full build takes seconds)
    ms, (idx, val, stats) = cx.timed_once(build)
    flops = 2.0 * NT16 * NS16 * DIM16
    res["knn_build"] = {"ms": ms, "tflops_all_gpus": flops / ms / 1e9, "frac_of_bf16_peak_per_gpu": flops / ms / 1e9 / world / peaks["bf16_tflops"],
                        "gpairs_per_s": NT16 * NS16 / ms / 1e6, "rows_per_rank": e - s, "exact_fallback_rows_this_rank": int(stats[0]),
                        "timed": "1 build after a warm-up on 8192 rows"}
    # sampled rows of the gathered lists against the exact CUDA-core sweep on this rank (rows of OTHER ranks too)
    rows = torch.arange(rank, NT16, NT16 // 256, device=dev)[:256]
    i0, v0, _, _ = ops.knn_cosine(u_tar[rows], u_src, K16, algo="simt")
    res["knn_build"]["gathered_lists_match_exact_sweep_on_256_sampled_rows"] = bool(torch.equal(idx[rows], i0) and torch.equal(val[rows], v0))
    del val, i0, v0
    # ---- bridged graph ---------------------------------------------------------------------------------------------
    n = NS16 + NT16
    y = torch.cat((y_src, y_tar))
    ei_u = build_bridged_graph(idx, y, NS16, n, dev, seed=1)
    del idx
    cm = torch.arange(n, device=dev) < NS16

    def x_rows(r0, r1):      # rows [r0, r1) of cat(u_src, u_tar) without materialising the concatenation
        parts = []
        if r0 < NS16:
            parts.append(u_src[r0:min(r1, NS16)])
        if r1 > NS16:
            parts.append(u_tar[max(r0, NS16) - NS16:r1 - NS16])
        return torch.cat(parts, 0).contiguous() if len(parts) > 1 else parts[0].contiguous()
    torch.manual_seed(0)
    model = KTGNN_no_complement(DIM16, N_CLASS, 2, HIDDEN, root_weight=False, use_bn=True, dim_share=DIM16, dropout=0.0).to(dev)
    model.train()
    if world > 1:
        step, data_loc, part = partitioned_step_fn(cx, model, x_rows, ei_u, cm, y)
        e_mp = cx.sum_over_ranks(float(data_loc.edge_index.shape[1]))
        res["row_blocks"] = {"bounds": part.bounds, "edges_this_rank": int(data_loc.edge_index.shape[1]), "needs_Hs_Ht_per_rank": part.needs(data_loc.central_mask)}
    else:
        from bridged_gnn_b200.data import Data
        data_loc = Data(x=x_rows(0, n), edge_index=ei_u, y=y, central_mask=cm)
        nll = nll_weighted(y, cm)

        def step():
            model.zero_grad(set_to_none=True)
            lb, lt, ltt, _ = model(data_loc)
            loss = nll(lb) + nll(lt) + nll(ltt)
            loss.backward()
            return loss
        step()
        e_mp = float(model.edge_index.shape[1])
    del u_src, u_tar
    torch.cuda.empty_cache()
    ms = cx.timed(step, 3, 2)
    res["ktgnn_step"] = {"ms": ms, "gedges_per_s": e_mp * 8 / ms / 1e6, "E_mp": int(e_mp), "loss": cx.sum_over_ranks(float(step())),
                         "what": "KT-GNN (F_in 256, hidden 64, 2 classes) train fwd+bwd%s" % (
                             ", destination-partitioned: domain-aware halo exchange, rank-combined BatchNorm, one flat gradient all-reduce" if world > 1 else "")}
    # ---- SpMM F = 256 (SAGE / GCN aggregation of configs[4]) ---------------------------------------------------------
    if world > 1:
        # equal row blocks here: the panels are all-gathered
        from bridged_gnn_b200.models import graph_partition
        upart = bdist.DstPartition(n, cx.hp_group())
        del data_loc, model, step
        torch.cuda.empty_cache()
        ei_loc = upart.local_edges(graph_partition(ei_u, cm)[2])
        part = upart
        graph = part.graph(ei_loc)
        x_loc = torch.randn(part.n_loc, DIM16, device=dev)        # (the features themselves were freed with the embeddings)
        t_pipe = cx.timed(lambda: bdist.partitioned_spmm(graph, x_loc, part, "mean", panels=4), 3, 2)
        t_mono = cx.timed(lambda: bdist.partitioned_spmm(graph, x_loc, part, "mean", panels=1), 3, 2)
        full = torch.empty((part.n_pad, DIM16), device=dev)
        t_kernel = cx.timed(lambda: ops._spmm_raw(graph.rowptr, graph.col, full, graph.n_rows, True), 3, 2)
        del full
        res["spmm_f256"] = {"ms_panel_pipelined": t_pipe, "ms_single_all_gather": t_mono, "ms_local_kernel_only": t_kernel,
                            "gedges_per_s": e_mp / t_pipe / 1e6, "halo_bytes_received_per_rank": (part.n_pad - part.n_loc) * DIM16 * 4,
                            "what": "mean aggregation of X [N, 256] over the partitioned graph: NCCL all-gather of X in 4 column panels, "
                                    "the gather kernel of panel p overlapping the transfer of panels p+1.."}
    else:
        graph = ops.cached_graph(model.edge_index, n)
        x_all = data_loc.x
        t = cx.timed(lambda: ops.spmm(graph, x_all, "mean"), 3, 2)
        b = e_mp * (4 + 4 * DIM16) + n * (4 + 4 * DIM16)
        res["spmm_f256"] = {"ms": t, "gedges_per_s": e_mp / t / 1e6, "per_edge_model_gbs": b / t / 1e6, "frac_of_hbm_peak": b / t / 1e6 / peaks["hbm_gbs"]}
    return res


def run_ours(args):
    from bridged_gnn_b200 import _lib, ops
    from bridged_gnn_b200 import dist as bdist
    from bridged_gnn_b200.data import Data
    from bridged_gnn_b200.models import KTGNN_no_complement

    cx = Ctx(args)
    dist, dev, rank, world = cx.dist, cx.dev, cx.rank, cx.world
    _lib.load()
    peaks = load_peaks()
    K, W = cx.K, cx.W
    timed = cx.timed

    # ---- data: every rank owns one sync-1M replica (weak scaling); the source set is shared -------------
    u_src, u_tar, y_src, y_tar = make_sync_embeddings(NS, NT, DIM, dev, seed=0, tar_seed=100 + rank)
    u_src_h, u_tar_h = u_src.cpu().pin_memory(), u_tar.cpu().pin_memory()

    # ---- phase A: bridged-graph build ------------------------------------------------------------------
    def build(us, ut):
        idx, val, gap, stats = ops.knn_cosine(ut, us, K_CROSS, normalize=True, apply_sigmoid=True, algo=args.knn_algo)
        if world > 1:   # per-replica kNN lists -> every rank (one NCCL all-gather each, over NVLink)
            bdist.all_gather_rows(idx, NT * world)
            bdist.all_gather_rows(val, NT * world)
        edges = bdist.edges_from_topk(idx)       # this rank's edge list (local target ids)
        return idx, val, gap, stats, edges

    clocks = ClockSampler(cx.local)
    clocks.start()
    launches0 = _lib.launches
    _lib.start_timing()
    knn_ms = timed(lambda: build(u_src, u_tar), K, W)
    knn_calls = _lib.stop_timing()
    knn_launches = (_lib.launches - launches0) * K // (K + W)
    idx, val, gap, stats, cross_edges = build(u_src, u_tar)
    n_fallback = int(stats[0].item())
    near_ties = int((gap < 1e-6).sum().item())

    edges_h = torch.empty(tuple(cross_edges.shape), dtype=cross_edges.dtype).pin_memory()

    # Serving-loop pipeline, as in the message-passing e2e below: the copy engine brings in the embeddings of build i+1
    # while build i runs (one input copy and one result copy per build, all inside the timed region).
    knn_copy_stream = torch.cuda.Stream(device=dev)

    def issue_embeddings():
        with torch.cuda.stream(knn_copy_stream):
            us, ut = u_src_h.to(dev, non_blocking=True), u_tar_h.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(knn_copy_stream)
        return us, ut, ev
    knn_pending = [issue_embeddings()]

    def build_e2e():
        cur = torch.cuda.current_stream(dev)
        us, ut, ev = knn_pending.pop()
        knn_pending.append(issue_embeddings())
        cur.wait_event(ev)
        for t in (us, ut):
            t.record_stream(cur)
        out = build(us, ut)
        edges_h.copy_(out[4], non_blocking=True)      # the edge list lands in a pinned host buffer
        cur.synchronize()
        return edges_h
    knn_e2e_ms = timed(build_e2e, max(3, K // 2), 3)
    knn_copy_stream.synchronize()
    del knn_pending
    calls, tot = knn_calls.get("bgnn_knn_cosine_f32", (1, 0.0))
    knn_call_ms = tot / max(calls, 1)
    flops = 2.0 * NT * NS * DIM
    half_rate = args.knn_algo in ("tc3", "tc1")
    tc_peak = peaks["bf16_tflops"] / (2.0 if half_rate else 1.0)
    traffic = load_traffic()
    knn_tr = traffic.get("knn_cosine_f16_kernel" if not half_rate else "knn_cosine_tc_kernel", {})
    knn_roof = {"bound": "tensor", "achieved": flops / (knn_call_ms * 1e-3) / 1e12, "peak": tc_peak, "unit": "TFLOP/s",
                "frac": flops / (knn_call_ms * 1e-3) / 1e12 / tc_peak, "traffic": knn_tr.get("bytes"),
                "traffic_source": knn_tr.get("source"),
                "kernel": ("knn_cosine_tc_kernel" if half_rate else "knn_cosine_f16_kernel")
                + " (timed: whole bgnn_knn_cosine_f32 call incl. prologue, merge/re-score and exact fallback)",
                "peak_source": peaks["source"] + (" bf16 dense / 2 (tcgen05 kind::tf32 runs at half the 16-bit rate)"
                                                  if half_rate else " dense bf16 == fp16 rate of tcgen05 kind::f16"),
                "algorithmic_flops_per_launch": flops}

    # ---- N > 1: ONE sync-1M build, target rows sharded over the ranks, lists all-gathered and VERIFIED ----
    sharded = None
    idx_global = None
    if world > 1:
        u_tar0 = u_tar if rank == 0 else make_sync_embeddings(NS, NT, DIM, dev, seed=0, tar_seed=100)[1]
        y_tar0 = y_tar if rank == 0 else make_sync_embeddings(NS, NT, DIM, dev, seed=0, tar_seed=100)[3]
        local_fn = lambda q, db, k: ops.knn_cosine(q, db, k, algo=args.knn_algo)[:3]        # noqa: E731
        sh_ms = timed(lambda: bdist.sharded_topk(u_tar0, u_src, K_CROSS, local_fn), K, W)
        idx_global, val_global, _ = bdist.sharded_topk(u_tar0, u_src, K_CROSS, local_fn)
        ok = torch.tensor([1.0], device=dev)
        if rank == 0:      # rank 0's own replica IS this problem: the 1-GPU lists are at hand
            ok[0] = float(torch.equal(idx_global, idx) and torch.equal(val_global, val))
        dist.broadcast(ok, 0)
        sharded = {"ms": sh_ms, "speedup_vs_1gpu_build": knn_ms / sh_ms, "rows_per_rank": NT // world,
                   "gathered_lists_bit_identical_to_1gpu": bool(ok.item()),
                   "what": "262 144 target rows split over %d ranks, source set replicated, idx + val all-gathered" % world}
        del val_global

    # ---- phase B: message passing over the bridged graph ------------------------------------------------
    n = NS + NT
    y = torch.cat((y_src, y_tar), 0)
    ei = build_bridged_graph(idx, y, NS, n, dev, seed=1 + rank)
    cm = torch.zeros(n, dtype=torch.bool, device=dev)
    cm[:NS] = True
    data = Data(x=torch.cat((u_src, u_tar), 0).contiguous(), edge_index=ei, y=y, central_mask=cm)
    del cross_edges, idx, val
    torch.manual_seed(0)
    model = KTGNN_no_complement(DIM, N_CLASS, 2, HIDDEN, root_weight=False, use_bn=True, dim_share=DIM,
                                need_complement=False, dropout=0.0).to(dev)
    model.train()
    nll = nll_weighted(y, cm)

    def train_step():
        model.zero_grad(set_to_none=True)
        lb, lt, ltt, _ = model(data)
        loss = nll(lb) + nll(lt) + nll(ltt)
        loss.backward()
        return loss

    def fwd_step():
        with torch.no_grad():
            return model(data)

    loss0 = float(train_step())        # builds + caches partition / CSR / transposed CSR
    e_mp = int(model.edge_index.shape[1])
    deg = torch.bincount(model.edge_index[1], minlength=n)
    deg_stats = {"mean": float(deg.float().mean()), "p99": int(torch.quantile(deg[::64].float(), 0.99)), "max": int(deg.max())}
    launches0 = _lib.launches
    _lib.start_timing()
    step_ms = timed(train_step, K, W)
    mp_calls = _lib.stop_timing()
    mp_launches = (_lib.launches - launches0) * K // (K + W)
    fwd_ms = timed(fwd_step, K, W)
    clk = clocks.stop()
    total_edges = cx.sum_over_ranks(float(e_mp))
    value = total_edges * 8 / (step_ms * 1e-3) / 1e9

    # e2e: host buffers in, log-probs out, through the public model API (fresh tensors -> CSR rebuilt too).  The edge
    # list crosses PCIe as int32 (node ids < 2^31; widened to the int64 the PyG-style API takes on the device).
    x_h, ei_h, cm_h = data.x.cpu().pin_memory(), ei.to(torch.int32).cpu().pin_memory(), cm.cpu().pin_memory()
    h2d = x_h.numel() * 4 + ei_h.numel() * 4 + cm_h.numel()
    d2h = 3 * n * N_CLASS * 4 + 4

    copy_stream = torch.cuda.Stream(device=dev)
    out_h = [torch.empty((n, N_CLASS), dtype=torch.float32).pin_memory() for _ in range(3)]
    loss_h = torch.empty((), dtype=torch.float32).pin_memory()

    def issue_inputs():
        """H2D of one step's inputs on the copy stream (graph first: the CSR build needs no features); returns the device
        tensors and the events that mark their arrival."""
        with torch.cuda.stream(copy_stream):
            ei_d = ei_h.to(dev, non_blocking=True)
            cm_d = cm_h.to(dev, non_blocking=True)
            ev_graph = torch.cuda.Event()
            ev_graph.record(copy_stream)
            x_d = x_h.to(dev, non_blocking=True)
            ev_x = torch.cuda.Event()
            ev_x.record(copy_stream)
        return ei_d, cm_d, x_d, ev_graph, ev_x

    pending = [issue_inputs()]

    def e2e_step():
        # Serving-loop pipeline: while step i computes, the copy engine already brings in the inputs of step i+1 (one
        # input copy and one result copy per step, all inside the timed region).  Within a step the graph arrives
        # first, so the partition / CSR / transposed CSR / row orders are built while the features are still on the wire.
        cur = torch.cuda.current_stream(dev)
        ei_d, cm_d, x_d, ev_graph, ev_x = pending.pop()
        pending.append(issue_inputs())
        cur.wait_event(ev_graph)
        d = Data(x=None, edge_index=ei_d.to(torch.int64), central_mask=cm_d)
        model.edge_index = None        # a new graph arrives: re-partition, rebuild CSR
        model.zero_grad(set_to_none=True)
        model.prepare_graph(d)
        cur.wait_event(ev_x)
        d.x = x_d
        for t in (ei_d, cm_d, x_d):
            t.record_stream(cur)
        lb, lt, ltt, _ = model(d)
        loss = nll(lb) + nll(lt) + nll(ltt)
        loss.backward()
        # results come back into pinned host buffers (one synchronisation for the three log-prob matrices and the loss)
        for buf, t in zip(out_h, (lb, lt, ltt)):
            buf.copy_(t.detach(), non_blocking=True)
        loss_h.copy_(loss.detach(), non_blocking=True)
        cur.synchronize()
        return out_h, float(loss_h)
    e2e_ms = timed(e2e_step, max(3, K // 2), 6)   # fresh tensors every step: the caching allocator keeps growing for ~5 steps
    e2e_value = total_edges * 8 / (e2e_ms * 1e-3) / 1e9
    model.edge_index = None
    train_step()                       # back on the resident graph
    # the same step replayed as one CUDA graph (bridged_gnn_b200.graphs.GraphedStep): reported beside the eager headline
    graphed = None
    try:
        from bridged_gnn_b200.graphs import GraphedStep
        gstep = GraphedStep(train_step)
        g_ms = timed(gstep, K, W)
        graphed = {"ms_per_step": g_ms, "gedges_per_s": total_edges * 8 / (g_ms * 1e-3) / 1e9, "loss": float(gstep()), "eager_loss": loss0}
        del gstep
    except Exception as ex:
        graphed = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:200])}
    torch.cuda.empty_cache()

    # roofline of the dominant message-passing kernel (largest share of the step among our kernels)
    shares = {k: v[1] / (K + W) for k, v in mp_calls.items()}       # ms per step (events span warm-up + timed steps)
    top = max(shares, key=shares.get) if shares else None
    roof = None
    if top:
        c = int(top.split("c=")[1].rstrip("]")) if "c=" in top else HIDDEN
        calls_per_step = mp_calls[top][0] / (K + W)
        dur_ms = shares[top] / max(calls_per_step, 1e-9)
        b = gat_bytes(e_mp, n, c, "bwd" in top)
        roof = {"bound": "hbm", "achieved": b / (dur_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": b / (dur_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": traffic.get(top, {}).get("bytes"),
                "traffic_source": traffic.get(top, {}).get("source"), "kernel": top,
                "avg_launch_ms": dur_ms, "share_of_step": shares[top] / step_ms, "algorithmic_bytes_per_launch": b,
                "peak_source": peaks["source"], "kernel_ms_per_step": {k: round(v, 4) for k, v in shares.items()}}

    extras = {}
    if world == 1 and not args.no_extras:
        for name, fn in (("spmm", lambda: measure_spmm(cx, ops.cached_graph(model.edge_index, n), n, peaks)),
                         ("addrelu", lambda: measure_addrelu(cx))):
            try:
                extras[name] = fn()
            except Exception as ex:      # an auxiliary measurement must not cost the headline line
                extras[name] = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:300])}
            torch.cuda.empty_cache()

    # ---- N > 1: ONE sync-1M KT-GNN step, destination rows partitioned over the ranks ----------------------
    partitioned = None
    if world > 1:
        try:
            y0 = torch.cat((y_src, y_tar0), 0)
            x0 = data.x if rank == 0 else torch.cat((u_src, u_tar0), 0)
            ei0 = ei if rank == 0 else build_bridged_graph(idx_global, y0, NS, n, dev, seed=1)
            torch.manual_seed(0)
            pmodel = KTGNN_no_complement(DIM, N_CLASS, 2, HIDDEN, root_weight=False, use_bn=True, dim_share=DIM,
                                         need_complement=False, dropout=0.0).to(dev)
            pmodel.train()
            pstep, pdata, part = partitioned_step_fn(cx, pmodel, lambda r0, r1: x0[r0:r1].contiguous(), ei0, cm, y0)
            ploss = pstep().detach().clone()
            dist.all_reduce(ploss)
            ref = torch.tensor([loss0], device=dev)
            dist.broadcast(ref, 0)      # rank 0's replica is this very graph and model initialisation
            p_ms = timed(pstep, K, W)
            # the same step replayed as ONE CUDA graph (NCCL exchanges included): removes the host's issue time, which
            # does not shrink with the number of ranks
            graphed_ms, graph_err = None, None
            try:
                from bridged_gnn_b200.graphs import GraphedStep
                gstep = GraphedStep(pstep)
                gl = gstep().detach().clone()
                dist.all_reduce(gl)
                graphed_ms = timed(gstep, K, W)
                graph_loss = float(gl)
            except Exception as ex:
                import traceback
                graph_err = "%s: %s || %s" % (type(ex).__name__, str(ex)[:200], " | ".join(
                    ln.strip() for ln in traceback.format_exc().splitlines() if "File" in ln)[-900:])
            # where the step goes: our kernels (CUDA events around every C-ABI call), the host's issue time (a step
            # enqueued without synchronising), the rest = NCCL exchanges + torch glue
            _lib.start_timing()
            for _ in range(3):
                pstep()
            pk = _lib.stop_timing()
            kernel_ms = sum(v[1] for v in pk.values()) / 3
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                pstep()
            host_ms = (time.perf_counter() - t0) / 3 * 1e3
            torch.cuda.synchronize()
            ref_step = torch.tensor([step_ms], device=dev)
            e0 = torch.tensor([float(e_mp)], device=dev)
            dist.broadcast(e0, 0)
            partitioned = {"ms_per_step": p_ms, "one_gpu_ms_per_step": step_ms, "ratio_to_1gpu_step": p_ms / step_ms,
                           "cuda_graph_ms_per_step": graphed_ms, "cuda_graph_ratio_to_1gpu_step": (graphed_ms / step_ms) if graphed_ms else None,
                           "cuda_graph_loss": graph_loss if graphed_ms else None, "cuda_graph_error": graph_err,
                           "gedges_per_s": float(e0) * 8 / p_ms / 1e6, "loss_partitioned": float(ploss), "loss_1gpu": float(ref),
                           "loss_rel_err": abs(float(ploss) - float(ref)) / max(abs(float(ref)), 1e-12),
                           "edges_this_rank": int(pdata.edge_index.shape[1]), "row_bounds": part.bounds,
                           "halo": "domain-aware point-to-point exchange of H; row blocks cut at equal edge counts",
                           "rank0_kernel_ms_per_step": kernel_ms, "rank0_host_issue_ms_per_step": host_ms,
                           "rank0_kernels": {k: round(v[1] / 3, 3) for k, v in pk.items()},
                           "what": "the 2^20-node graph of rank 0 with destination rows split over %d ranks: strong scaling of ONE step" % world}
            del pmodel, pdata, x0, ei0
        except Exception as ex:
            partitioned = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:300])}
        torch.cuda.empty_cache()

    # ---- configs[4] ------------------------------------------------------------------------------------------
    sync16m = None
    if not args.no_sync16m:
        del data, model, u_src, u_tar, x_h, ei_h, u_src_h, u_tar_h, ei
        torch.cuda.empty_cache()
        try:
            sync16m = run_sync16m(cx, peaks)
        except Exception as ex:
            sync16m = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:300])}
            # every rank must leave the collective sequence together: a failure on one rank ends the run for all
        torch.cuda.empty_cache()

    if world == 1 and not args.no_extras:      # last: CUDA-graph capture lives here
        try:
            extras["configs_1_3"] = measure_small_configs(cx)
        except Exception as ex:
            extras["configs_1_3"] = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:300])}

    if rank == 0:
        cpu = cpu_mp_sample() if world == 1 and not args.no_cpu_baseline else None
        cpu_knn = cpu_knn_sample(rows=24) if world == 1 and not args.no_cpu_baseline else None
        line = {
            "metric": METRIC, "value": value, "unit": "GEdges/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": CONFIG,
            "run": {"E_mp": e_mp, "in_degree": deg_stats, "knn_algo": args.knn_algo, "parallelism": "replicas x%d (headline); "
                    "row-sharded build / destination-partitioned step in sharded_build_1m, partitioned_1m, sync16m" % world},
            "fwd_only": {"ms": fwd_ms, "gedges_per_s": total_edges * 4 / (fwd_ms * 1e-3) / 1e9},
            "step_cuda_graph": graphed,
            "e2e": {"value": e2e_value, "unit": "GEdges/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "roofline": roof,
            "knn_build": {"ms": knn_ms, "gpairs_per_s": NT * NS * world / (knn_ms * 1e-3) / 1e9,
                          "tflops": flops * world / (knn_ms * 1e-3) / 1e12, "kernel_call_ms": knn_call_ms,
                          "roofline": knn_roof, "exact_fallback_rows": n_fallback, "near_tie_rows": near_ties,
                          "e2e": {"ms": knn_e2e_ms, "h2d_bytes_per_step": (NS + NT) * DIM * 4,
                                  "d2h_bytes_per_step": 2 * NT * K_CROSS * 8},
                          "cpu_baseline": cpu_knn, "gpu_launches": knn_launches},
            "cpu_baseline": cpu, "gpu_launches": mp_launches, "clocks": clk,
        }
        if sharded is not None:
            line["sharded_build_1m"] = sharded
        if partitioned is not None:
            line["partitioned_1m"] = partitioned
        if sync16m is not None:
            line["sync16m"] = sync16m
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL work captured in CUDA graphs makes ProcessGroupNCCL's teardown wait for minutes (measured: the process
        # sat in destroy_process_group until the launcher's timeout): leave together, without the teardown
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--knn-algo", dest="knn_algo", default="f16", choices=["f16", "tc3", "tc1", "simt"])
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--no-extras", dest="no_extras", action="store_true", help="skip the SpMM / add-ReLU / configs 1-3 blocks (N = 1)")
    ap.add_argument("--no-sync16m", dest="no_sync16m", action="store_true", help="skip BASELINE configs[4]")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
