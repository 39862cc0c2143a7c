from .KTGNN import AdaptedConv, KTGNN_no_complement, KTGNN_noDTC, graph_partition  # noqa: F401
from .backbones import GCNConv, GCNNet, GraphSAGE, SAGEConv  # noqa: F401
from .models import Adversarial_Learner, Adversarial_Learner_v2, Similar, Similar_v2  # noqa: F401
