"""Inference-side mirror of the reference's similarity learner (models/models.py): the module tree,
constructor arguments and state_dict keys of ``Adversarial_Learner`` (v1: SAGE backbone + cosine head,
:815-844) and ``Adversarial_Learner_v2`` (:1110-1142: mlp|gnn backbone, cosine|mlp head), so the
shipped ``ckpt/model_AdvLearner_*_best.ckpt`` load with ``strict=True``.

Only what the bridged-graph build consumes is evaluated here: node embeddings ``z`` (backbone /
``target_learner.encode``), the node-wise operands of the pair similarity, and the node classifier.
The pair similarity itself is never evaluated pair by pair: ``main_bridged_graph`` hands the operands to
the fused kNN kernels.  Adversarial training (scripts.py) is outside the accelerated path; the decoder
and discriminator exist only so that checkpoints load.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .backbones import SAGEConv


def _bn_fold(bn):
    """(scale, shift) of an eval-mode BatchNorm1d in float64: y = x * scale + shift."""
    f64 = torch.float64
    s = bn.weight.detach().to(f64) / torch.sqrt(bn.running_var.to(f64) + bn.eps)
    return s, bn.bias.detach().to(f64) - bn.running_mean.to(f64) * s


class PairNorm(nn.Module):
    """models/models.py:29-64."""

    def __init__(self, mode="PN", scale=10):
        super().__init__()
        assert mode in ("None", "PN", "PN-SI", "PN-SCS")
        self.mode, self.scale = mode, scale

    def forward(self, x):
        if self.mode == "None":
            return x
        col_mean = x.mean(dim=0)
        if self.mode == "PN":
            x = x - col_mean
            return self.scale * x / (1e-6 + x.pow(2).sum(dim=1).mean()).sqrt()
        if self.mode == "PN-SI":
            x = x - col_mean
            return self.scale * x / (1e-6 + x.pow(2).sum(dim=1, keepdim=True)).sqrt()
        return self.scale * x / (1e-6 + x.pow(2).sum(dim=1, keepdim=True)).sqrt() - col_mean


def _glorot_linear(i, o, bias=True):
    lin = nn.Linear(i, o, bias=bias)
    nn.init.xavier_uniform_(lin.weight)
    if bias:
        nn.init.zeros_(lin.bias)
    return lin


def _dims(dim_in, dim_hidden, dim_out, layer_num):
    if layer_num == 1:
        return [(dim_in, dim_out)]
    return [(dim_in if n == 0 else dim_hidden, dim_out if n == layer_num - 1 else dim_hidden) for n in range(layer_num)]


class MLP(nn.Module):
    """models/models.py:852-893; ``edge_index`` is accepted and ignored like the reference."""

    def __init__(self, dim_in, dim_out, dim_hidden=64, layer_num=2, root_weight=True, use_norm=False,
                 norm_mode="PN-SCS", norm_scale=1, log_softmax=False):
        super().__init__()
        self.layers = nn.ModuleList(_glorot_linear(i, o) for i, o in _dims(dim_in, dim_hidden, dim_out, layer_num))
        self.use_norm, self.log_softmax = use_norm, log_softmax
        if use_norm:
            self.norm = PairNorm(norm_mode, norm_scale)

    def forward(self, x, edge_index=None):
        last = len(self.layers) - 1
        plain_norm = not self.use_norm or self.norm.mode == "None"
        for ind, layer in enumerate(self.layers):
            if not self.training and not torch.is_grad_enabled() and ops.dense_supported(x, layer.in_features, layer.out_features) \
                    and (ind == last or plain_norm):
                # inference on the device (the bridged-graph build): one tcgen05 row-panel GEMM per layer, bias and
                # ReLU applied to the accumulator tile
                x = ops.rowpanel_gemm(x, layer.weight, layer.bias, act=None if ind == last else "relu")
                continue
            x = layer(x)
            if ind != last:
                if self.use_norm:
                    x = self.norm(x)
                x = F.dropout(F.relu(x), p=0.5, training=self.training)
        return F.log_softmax(x, dim=1) if self.log_softmax else x


class GraphEncoder(nn.Module):
    """models/models.py:220-263: stacked SAGEConv(mean) over a Tensor edge_index (CSR SpMM kernel)."""

    def __init__(self, dim_in, dim_out, dim_hidden=64, layer_num=2, root_weight=True, norm_mode="PN-SCS",
                 norm_scale=1, log_softmax=False):
        super().__init__()
        self.convs = nn.ModuleList(SAGEConv(i, o, root_weight=root_weight)
                                   for i, o in _dims(dim_in, dim_hidden, dim_out, layer_num))
        self.norm = PairNorm(norm_mode, norm_scale)
        self.log_softmax = log_softmax

    def forward(self, x, edge_index):
        for ind, conv in enumerate(self.convs):
            x = conv(x, edge_index)
            if ind != len(self.convs) - 1:
                x = F.dropout(F.relu(self.norm(x)), p=0.5, training=self.training)
        return F.log_softmax(x, dim=1) if self.log_softmax else x


class _LayerStack(nn.Module):
    """Parameter holder for Decoder (:653) / Discriminator (:753): ``layers.{i}.weight|bias``."""

    def __init__(self, dims):
        super().__init__()
        self.layers = nn.ModuleList(nn.Linear(i, o) for i, o in dims)

    def forward(self, *a, **k):
        raise NotImplementedError("decoder / discriminator belong to adversarial training, not to the graph build")


class Similar_v2(nn.Module):
    """Pair-similarity head (models/models.py:895-977; v1 ``Similar`` :67-169 is the cosine mode).

    cosine: sim(i,j) = sigmoid(cos(u_i, u_j)), u = lin_self(z) + biasatt(lin_self(z))
    mlp   : sim(i,j) = sigmoid(lin_self(cat(z_i, z_j))) with lin_self = BN, Linear, BN, ReLU, Linear
    """

    def __init__(self, in_channels, num_clf_classes, dropout=0.6, use_clf=True, mode="cosine"):
        super().__init__()
        self.mode, self.use_clf, self.dropout = mode, use_clf, dropout
        if mode == "cosine":
            self.biasatt = nn.Sequential(nn.Linear(128, 64), nn.Tanh(), nn.Linear(64, 128))
            for m in self.biasatt:
                if isinstance(m, nn.Linear):
                    nn.init.kaiming_normal_(m.weight)
                    nn.init.constant_(m.bias, 0)
            self.lin_self = nn.Sequential(nn.BatchNorm1d(in_channels), _glorot_linear(in_channels, 64, False),
                                          nn.BatchNorm1d(64), nn.Tanh(), _glorot_linear(64, 128, False))
        elif mode == "mlp":
            self.lin_self = nn.Sequential(nn.BatchNorm1d(in_channels * 2), _glorot_linear(in_channels * 2, 128),
                                          nn.BatchNorm1d(128), nn.ReLU(), _glorot_linear(128, 1))
        else:
            raise NotImplementedError("Not Supported Mode:{}".format(mode))
        if use_clf:
            self.lin_clf = _glorot_linear(in_channels, num_clf_classes)

    def clf_log_probs(self, z):
        """:137-140 / :961-964: log_softmax(lin_clf(dropout(relu(z))))."""
        if not self.use_clf:
            return None
        return F.log_softmax(self.lin_clf(F.dropout(F.relu(z), p=self.dropout, training=self.training)), dim=-1)

    def cosine_operand(self, z):
        """Node-wise vector fed to CosineSimilarity (:125-128, :946-948): u = z' + biasatt(z'), z' = lin_self(z).
        Inference on the device: four row-panel GEMMs -- the two eval-mode BatchNorms folded into the first layer's
        weights (float64 fold of the 64 x h matrix) and epilogue scale / shift, Tanh and the residual fused."""
        bn1, lin1, bn2, _, lin2 = self.lin_self
        if (not self.training and not torch.is_grad_enabled() and ops.dense_supported(z, lin1.in_features, lin1.out_features)
                and lin1.bias is None and lin2.bias is None):
            s1, t1 = _bn_fold(bn1)
            s2, t2 = _bn_fold(bn2)
            w1 = lin1.weight.detach().to(torch.float64)
            y1 = ops.rowpanel_gemm(z, (w1 * s1).float(), bias=((w1 @ t1) * s2 + t2).float(), scale=s2.float(), act="tanh")
            zz = ops.rowpanel_gemm(y1, lin2.weight)
            b1 = ops.rowpanel_gemm(zz, self.biasatt[0].weight, self.biasatt[0].bias, act="tanh")
            return ops.rowpanel_gemm(b1, self.biasatt[2].weight, self.biasatt[2].bias, res=zz)
        zz = self.lin_self(z)
        return zz + self.biasatt(zz)

    def mlp_operands(self, z_db, z_q):
        """Eval-mode fold of the mlp head into node-wise operands:
            logit(i, j) = sum_h w2[h] * relu(U_db[i,h] + U_q[j,h]) + b2
        with cat order (z_db[idx1], z_q[idx2]) as in :949-954.  Folded in fp64, rounded once to fp32."""
        if self.training:
            raise RuntimeError("the BatchNorm fold needs eval mode (running statistics)")
        bn1, lin1, bn2, _, lin2 = self.lin_self
        d = z_db.shape[1]
        f64 = torch.float64
        s1, t1 = _bn_fold(bn1)
        s2, t2 = _bn_fold(bn2)
        W = lin1.weight.detach().to(f64)
        Wa, Wb = W[:, :d], W[:, d:]
        # everything that is not an [N, d] operand is folded in float64 on the 128 x d weight blocks:
        #   U_db = z_db (s2 (.) Wa (.) s1a)^T + s2 (.) (Wa t1a),   U_q = z_q (s2 (.) Wb (.) s1b)^T + s2 (.) (Wb t1b + b1) + t2
        w_db, b_db = (Wa * s1[:d]) * s2[:, None], (Wa @ t1[:d]) * s2
        w_q, b_q = (Wb * s1[d:]) * s2[:, None], (Wb @ t1[d:] + lin1.bias.detach().to(f64)) * s2 + t2
        w2, b2 = lin2.weight.detach().view(-1).float(), float(lin2.bias.item())
        if ops.dense_supported(z_db, d, lin1.out_features) and ops.dense_supported(z_q, d, lin1.out_features):
            # the two [N, d] x [d, 128] products on the tensor cores (3 x TF32, fp32-grade), bias in the epilogue
            return (ops.rowpanel_gemm(z_db, w_db.float(), b_db.float()), ops.rowpanel_gemm(z_q, w_q.float(), b_q.float()), w2, b2)
        U_db = z_db.to(f64) @ w_db.t() + b_db
        U_q = z_q.to(f64) @ w_q.t() + b_q
        return U_db.float().contiguous(), U_q.float().contiguous(), w2, b2


class Similar(Similar_v2):
    def __init__(self, in_channels, num_clf_classes, dropout=0.6, use_clf=True):
        super().__init__(in_channels, num_clf_classes, dropout, use_clf, mode="cosine")


def _num_classes(data):
    return int(data.y.max().item()) + 1


class Source_Learner(nn.Module):
    def __init__(self, data, dim_hidden=64, norm_mode="None", norm_scale=1, use_clf=True):
        super().__init__()
        self.dim_in, self.num_classes, self.dim_hidden = data.num_features, _num_classes(data), dim_hidden
        self.backbone = GraphEncoder(self.dim_in, dim_hidden, dim_hidden, 2, True, norm_mode, norm_scale)
        self.sim_net = Similar(dim_hidden, self.num_classes, 0.6, use_clf)


class Source_Learner_v2(nn.Module):
    def __init__(self, data, dim_hidden=64, norm_mode="None", norm_scale=1, use_clf=True, use_norm=True,
                 backbone="mlp", mode="cosine"):
        super().__init__()
        self.dim_in, self.num_classes, self.dim_hidden = data.num_features, _num_classes(data), dim_hidden
        if backbone == "gnn":
            # (the reference passes use_norm= to GraphEncoder here, which it does not accept: models.py:1007-1009)
            self.backbone = GraphEncoder(self.dim_in, dim_hidden, dim_hidden, 2, True, norm_mode, norm_scale)
        elif backbone == "mlp":
            self.backbone = MLP(self.dim_in, dim_hidden, dim_hidden, 2, True, use_norm, norm_mode, norm_scale)
        else:
            raise NotImplementedError("Not Implemented Backbone:{}".format(backbone))
        self.sim_net = Similar_v2(dim_hidden, self.num_classes, 0.6, use_clf, mode)


class _TargetBase(nn.Module):
    def encode(self, data):
        """models/models.py:735-739 / 1092-1096.  Inference on the device: Linear + Tanh of equavilent_trans_layer as one
        row-panel GEMM (PairNorm mode 'None' is the identity)."""
        lin, norm, _ = self.equavilent_trans_layer
        x = data.x
        if (not self.training and not torch.is_grad_enabled() and norm.mode == "None"
                and ops.dense_supported(x, lin.in_features, lin.out_features)):
            h0 = ops.rowpanel_gemm(x, lin.weight, lin.bias, act="tanh")
        else:
            h0 = self.equavilent_trans_layer(x)
        return self.encoder(h0, getattr(data, "edge_index", None)), h0


class Target_Learner_AE(_TargetBase):
    def __init__(self, data, dim_eq_trans=128, dim_hidden=64, norm_mode="None", norm_scale=1):
        super().__init__()
        self.dim_in, self.dim_eq_trans, self.dim_hidden = data.num_features, dim_eq_trans, dim_hidden
        self.equavilent_trans_layer = nn.Sequential(nn.Linear(self.dim_in, dim_eq_trans),
                                                    PairNorm(norm_mode, norm_scale), nn.Tanh())
        self.encoder = GraphEncoder(dim_eq_trans, dim_hidden, dim_hidden, 2, True, norm_mode, norm_scale)
        self.decoder = _LayerStack(_dims(dim_hidden, dim_hidden, dim_eq_trans, 2))


class Target_Learner_AE_v2(_TargetBase):
    def __init__(self, data, dim_eq_trans=128, dim_hidden=64, use_norm=True, norm_mode="None", norm_scale=1,
                 backbone="mlp"):
        super().__init__()
        self.dim_in, self.dim_eq_trans, self.dim_hidden = data.num_features, dim_eq_trans, dim_hidden
        self.equavilent_trans_layer = nn.Sequential(nn.Linear(self.dim_in, dim_eq_trans),
                                                    PairNorm(norm_mode, norm_scale), nn.Tanh())
        if backbone == "gnn":
            self.encoder = GraphEncoder(dim_eq_trans, dim_hidden, dim_hidden, 2, True, norm_mode, norm_scale)
        elif backbone == "mlp":
            self.encoder = MLP(dim_eq_trans, dim_hidden, dim_hidden, 2, True, use_norm, norm_mode, norm_scale)
        else:
            raise NotImplementedError("Not Implemented Backbone:{}".format(backbone))
        self.decoder = _LayerStack(_dims(dim_hidden, dim_hidden, dim_eq_trans, 2))


class _AdvBase(nn.Module):
    """Node-wise half of get_probs_cross_domain / get_probs_within_domain (:824-844, :1122-1142): the
    embeddings and classifier outputs.  The pair half is the fused kNN kernel's job."""

    def embed_source(self, data):
        return self.source_learner.backbone(data.x, getattr(data, "edge_index", None))

    def embed_target(self, data):
        return self.target_learner.encode(data)[0]

    def clf_probs(self, z):
        lp = self.source_learner.sim_net.clf_log_probs(z)
        return None if lp is None else lp.exp()


class Adversarial_Learner(_AdvBase):
    def __init__(self, data_src, data_tar, dim_hidden=64, num_layer=2, source_clf=True, norm_mode="PN", norm_scale=1.0):
        super().__init__()
        self.num_layer, self.source_clf = num_layer, source_clf
        self.source_learner = Source_Learner(data_src, dim_hidden, norm_mode, norm_scale, source_clf)
        self.target_learner = Target_Learner_AE(data_tar, 128, dim_hidden, norm_mode, norm_scale)
        self.discriminator = _LayerStack([(dim_hidden, dim_hidden), (dim_hidden, 1)])


class Adversarial_Learner_v2(_AdvBase):
    def __init__(self, data_src, data_tar, dim_hidden=64, num_layer=2, source_clf=True, use_norm=True, norm_mode="PN",
                 norm_scale=1.0, backbone="mlp", sim_mode="cosine"):
        super().__init__()
        self.num_layer, self.source_clf = num_layer, source_clf
        self.source_learner = Source_Learner_v2(data_src, dim_hidden, norm_mode, norm_scale, source_clf, use_norm,
                                                backbone, sim_mode)
        self.target_learner = Target_Learner_AE_v2(data_tar, 128, dim_hidden, use_norm, norm_mode, norm_scale, backbone)
        self.discriminator = _LayerStack([(dim_hidden, dim_hidden), (dim_hidden, 1)])
