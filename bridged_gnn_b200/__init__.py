"""bridged_gnn_b200 -- B200-native (sm_100a) implementation of Bridged-GNN's data-parallel hot path:
bridged-graph construction (fused similarity + top-k) and message passing over the resulting graph.

The compute lives in ``libbgnn_b200.so`` (hand-written CUDA behind the C ABI of ``include/bgnn_b200.h``);
this package is the Python host that mirrors the reference's function / layer API.
"""
from . import ops  # noqa: F401
from .data import Data, load_pyg_dat, to_undirected  # noqa: F401

__version__ = "0.1.0"
