"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/bgnn_b200.h declares, and the host-side argument validation that needs no GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from bridged_gnn_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build(verbose=False)
    return _lib.load()


def _declared():
    src = open(os.path.join(ROOT, "include", "bgnn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bgnn_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n


def test_ctypes_signatures_cover_header(lib):
    from bridged_gnn_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_version_and_error_strings(lib):
    assert lib.bgnn_version() >= 100
    assert lib.bgnn_error_string(0) == b"ok"
    assert b"invalid" in lib.bgnn_error_string(-1)
    assert b"workspace" in lib.bgnn_error_string(-2)


def test_workspace_queries_are_host_only(lib):
    # pure host arithmetic: callable without a GPU
    assert lib.bgnn_knn_cosine_workspace_bytes(591, 2817, 128, 20, 0) > 0
    assert lib.bgnn_knn_cosine_workspace_bytes(262144, 786432, 128, 20, 1) > 262144 * 128 * 4
    assert lib.bgnn_knn_cosine_workspace_bytes(10, 5, 16, 6, 0) == 0      # k > ndb
    assert lib.bgnn_knn_addrelu_workspace_bytes(591, 2817, 128, 20) > 0
    assert lib.bgnn_edges_to_csr_workspace_bytes(37522) > 37522 * 8
    assert lib.bgnn_gatv2_bwd_workspace_bytes(3408, 37522, 64) > 37522 * 16


def test_invalid_arguments_rejected_before_any_launch(lib):
    null = ctypes.c_void_p(None)
    # NULL inputs / k > ndb -> BGNN_ERR_INVALID_ARG (-1) without touching the device
    assert lib.bgnn_knn_cosine_f32(null, 4, null, 4, 8, 2, 1, 1, 0, null, null, null, null, null, 0, null) == -1
    assert lib.bgnn_knn_addrelu_f32(null, 4, null, 4, 8, null, 0.0, 9, 1, null, null, null, null, 0, null) == -1
    assert lib.bgnn_gatv2_fwd_f32(null, null, null, null, null, null, null, 0.1, 5, 0, null, null, null, null) == -1


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "bridged_gnn_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dp, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from bridged_gnn_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_cpu_tensors_rejected():
    import torch
    from bridged_gnn_b200 import _lib
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        _lib.ptr(torch.zeros(4))


def test_dense_kernel_plans_are_host_only(lib):
    """Shape predicates and workspace sizes of the node-wise dense kernels: host arithmetic, no GPU needed."""
    # row-panel GEMM: K, NO <= 256, 16-byte rows
    assert lib.bgnn_rowpanel_gemm_supported(128, 128, 130) == 1
    assert lib.bgnn_rowpanel_gemm_supported(130, 132, 128) == 1
    assert lib.bgnn_rowpanel_gemm_supported(128, 130, 64) == 0        # 520-byte rows: TMA cannot address them
    assert lib.bgnn_rowpanel_gemm_supported(64, 64, 257) == 0
    # wide / narrow AdaptedConv transforms
    assert lib.bgnn_adapted_wide_supported(64, 128) == 1 and lib.bgnn_adapted_wide_supported(32, 256) == 1
    assert lib.bgnn_adapted_wide_supported(48, 128) == 0 and lib.bgnn_adapted_wide_supported(160, 128) == 0
    assert lib.bgnn_adapted_wide_supported(64, 130) == 0
    assert lib.bgnn_adapted_skinny_heads_supported(2, 64, 2) == 1 and lib.bgnn_adapted_skinny_heads_supported(2, 64, 3) == 0
    assert lib.bgnn_adapted_skinny_heads_supported(5, 64, 1) == 0
    assert lib.bgnn_adapted_skinny_tc_supported(2, 64, 2) == 1 and lib.bgnn_adapted_skinny_tc_supported(4, 256, 2) == 1
    assert lib.bgnn_adapted_skinny_tc_supported(2, 66, 1) == 0
    # weight-gradient GEMM: d <= 128, no <= 256, workspace = one [128, nop] partial per SM
    assert lib.bgnn_wgrad_gemm_supported(128, 128, 130, 132) == 1 and lib.bgnn_wgrad_gemm_supported(129, 132, 64, 64) == 0
    assert lib.bgnn_wgrad_gemm_supported(64, 64, 64, 66) == 0
    assert lib.bgnn_wgrad_gemm_workspace_bytes(130) >= 148 * 128 * 160 * 4
    assert lib.bgnn_wgrad_gemm_workspace_bytes(64) < lib.bgnn_wgrad_gemm_workspace_bytes(130)
    # BatchNorm + ReLU: c % 4 == 0, c <= 1024
    assert lib.bgnn_bn_relu_supported(64) == 1 and lib.bgnn_bn_relu_supported(6) == 0 and lib.bgnn_bn_relu_supported(2048) == 0
    assert lib.bgnn_bn_relu_workspace_bytes(64) > 148 * 8 * 128 * 4


def test_dense_kernels_reject_invalid_arguments(lib):
    null = ctypes.c_void_p(None)
    assert lib.bgnn_rowpanel_gemm_f32(null, 4, 8, 8, null, null, null, 8, null, 8, null) == -1
    assert lib.bgnn_wgrad_gemm_f32(null, 8, 8, null, 8, 8, 4, null, 8, null, null, 0, null) == -1
    assert lib.bgnn_bn_relu_fwd_f32(null, 4, 8, null, null, 1e-5, 0.1, null, null, 1, null, null, null, 0, null) == -1
    assert lib.bgnn_adapted_wide_fwd_f32(null, 4, 8, null, null, 32, null, null, null, null, null, null, null, null) == -1
    assert lib.bgnn_tf32_planes_f32(null, 4, 4, 4, 1, 16, 32, null, null, null) == -1


def test_round2_entry_points_reject_invalid_arguments(lib):
    """Host-side validation of the entry points added in round 2 (no launch happens on these paths)."""
    null = ctypes.c_void_p(None)
    nan = float("nan")
    # kNN with the epsilon epilogue: NULL inputs / k > ndb
    assert lib.bgnn_knn_cosine_eps_f32(null, 4, null, 4, 8, 2, 1, 1, 0, nan, null, null, null, null, null, null, 0, null) == -1
    assert lib.bgnn_knn_addrelu_eps_f32(null, 4, null, 4, 8, null, 0.0, 9, 1, 0.5, null, null, null, null, null, 0, null) == -1
    # quantile: empty input, rank out of range, weight outside [0, 1]
    assert lib.bgnn_quantile_workspace_bytes() >= 1024
    assert lib.bgnn_quantile_f32(null, 0, 0, 0.0, null, null, 0, null) == -1
    one = ctypes.c_void_p(16)          # a non-NULL pointer that is never dereferenced on these paths
    assert lib.bgnn_quantile_f32(one, 10, 10, 0.0, one, one, 4096, null) == -1
    assert lib.bgnn_quantile_f32(one, 10, 3, 1.5, one, one, 4096, null) == -1
    assert lib.bgnn_quantile_f32(one, 10, 3, 0.5, one, null, 0, null) == -2          # workspace
    assert lib.bgnn_edge_validity_f32(null, null, 5, null, null, null, null, null, null, null, null, null, null, 8, 0.0, null, one, null) == -1
    # destination-partitioned aggregation: the local rows must lie inside the node range
    assert lib.bgnn_gatv2_fwd_part_f32(null, null, null, null, null, null, null, null, 0.1, 4, -1, 8, null, null, null, null, null) == -1
    assert lib.bgnn_gatv2_bwd_part_f32(*([null] * 7), 0, *([null] * 5), 0.1, 8, 4, 10, 8, *([null] * 10), 0, null) == -1     # 4 + 8 > 10
    assert lib.bgnn_gatv2_heads_bwd_part_f32(*([null] * 5), 0, *([null] * 5), 0.1, 8, 4, 10, 3, 2, *([null] * 9), 0, null) == -1
    # row-panel GEMM epilogue: unknown activation, residual narrower than the output
    assert lib.bgnn_rowpanel_gemm_act_f32(one, 4, 8, 8, one, one, null, null, 3, null, 0, 8, one, 8, null) == -1
    assert lib.bgnn_rowpanel_gemm_act_f32(one, 4, 8, 8, one, one, null, null, 1, one, 4, 8, one, 8, null) == -1
    # strided SpMM: strides shorter than the width
    assert lib.bgnn_spmm_csr_ld_f32(one, one, null, null, null, one, 4, 10, 8, 0, one, 8, null) == -1
    assert lib.bgnn_bn_relu_bwd_apply_f32(null, null, 4, 8, null, 1, null, null, null) == -1
    # the backward workspace now also holds scores and per-CTA partials of pass B: grows with n and e
    a, b = lib.bgnn_gatv2_bwd_workspace_bytes(1000, 20000, 64), lib.bgnn_gatv2_bwd_workspace_bytes(2000, 40000, 64)
    assert b > a > 20000 * 20
    # one-call graph preparation: NULL edge list with e > 0, the three transposed outputs all or none, workspace
    assert lib.bgnn_graph_prepare_workspace_bytes(37522, 3408, 1) > (37522 + 3408) * 24
    assert lib.bgnn_graph_prepare_workspace_bytes(-1, 3408, 1) == 0
    assert lib.bgnn_graph_prepare(null, null, 5, 10, 1, one, one, null, null, null, null, null, one, one, 1 << 20, null) == -1
    assert lib.bgnn_graph_prepare(one, one, 5, 10, 1, one, one, one, null, null, null, null, one, one, 1 << 20, null) == -1
    assert lib.bgnn_graph_prepare(one, one, 5, 10, 1, one, one, null, null, null, null, one, one, one, 1 << 20, null) == -1
    assert lib.bgnn_graph_prepare(one, one, 5, 10, 1, one, one, null, null, null, null, null, one, one, 64, null) == -2

