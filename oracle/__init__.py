"""CPU oracle for the Bridged-GNN hot path.  TEST INFRASTRUCTURE ONLY.

A restatement, in plain torch-on-CPU / numpy, of the reference's op sequence for
(1) bridged-graph construction and (2) message passing, each function citing
the reference file:line it follows (paths relative to the reference checkout,
``Bridged-GNN/...``).  Semantics of the un-vendored third-party calls
(torch_geometric ``softmax``/``coalesce``/``add_self_loops``/``propagate``/
``SAGEConv``/``GCNConv``, torch_sparse ``matmul``; versions unpinned upstream,
PyG 2.1-2.3 era) are restated from their published behaviour.

Pinning: ``tests/test_oracle_golden.py`` checks every function here against
``tests/golden/*.npz``, which were produced by running the reference's own
unmodified Python (``tests/golden/make_golden.py``) on the shipped office
checkpoint + bridged graph and on seeded inputs with the shipped fb checkpoint.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product
(``bridged_gnn_b200``) never does.
"""
from . import build_oracle, mp_oracle  # noqa: F401
