"""ctypes binding of libbgnn_b200.so (the C ABI in include/bgnn_b200.h).

This is the only place the package touches native code.  There is no CPU or pure-torch fallback: if
the library is missing, or a tensor is not a contiguous CUDA tensor of the expected dtype, the call
raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbgnn_b200.so")

KNN_SIMT_F32 = 0
KNN_TC_3XTF32 = 1
KNN_TC_1XTF32 = 2
KNN_TC_F16 = 3

_c = ctypes
_vp, _i64, _i32, _f32, _sz = _c.c_void_p, _c.c_int64, _c.c_int, _c.c_float, _c.c_size_t

# name -> (restype, argtypes); mirrors include/bgnn_b200.h one to one
SIGNATURES = {
    "bgnn_version": (_i32, []),
    "bgnn_error_string": (_c.c_char_p, [_i32]),
    "bgnn_knn_cosine_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32, _i32]),
    "bgnn_knn_cosine_f32": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "bgnn_knn_cosine_eps_f32": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "bgnn_knn_addrelu_eps_f32": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _f32, _i32, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "bgnn_knn_addrelu_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32]),
    "bgnn_knn_addrelu_f32": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _f32, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "bgnn_quantile_workspace_bytes": (_sz, []),
    "bgnn_quantile_f32": (_i32, [_vp, _i64, _i64, _f32, _vp, _vp, _sz, _vp]),
    "bgnn_edge_validity_f32": (_i32, [_vp, _vp, _i64] + [_vp] * 10 + [_i32, _f32, _vp, _vp, _vp]),
    "bgnn_edges_to_csr_workspace_bytes": (_sz, [_i64]),
    "bgnn_edges_to_csr": (_i32, [_vp, _vp, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "bgnn_graph_prepare_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "bgnn_graph_prepare": (_i32, [_vp, _vp, _i64, _i64, _i32] + [_vp] * 9 + [_sz, _vp]),
    "bgnn_spmm_csr_f32": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "bgnn_spmm_csr_ld_f32": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _i64, _vp]),
    "bgnn_gatv2_fwd_part_f32": (_i32, [_vp] * 8 + [_f32, _i64, _i64, _i32] + [_vp] * 5),
    "bgnn_gatv2_bwd_part_f32": (_i32, [_vp] * 7 + [_i64] + [_vp] * 5 + [_f32, _i64, _i64, _i64, _i32] + [_vp] * 10 + [_sz, _vp]),
    "bgnn_gatv2_heads_fwd_part_f32": (_i32, [_vp] * 7 + [_f32, _i64, _i64, _i32, _i32] + [_vp] * 4),
    "bgnn_gatv2_heads_bwd_part_f32": (_i32, [_vp] * 5 + [_i64] + [_vp] * 5 + [_f32, _i64, _i64, _i64, _i32, _i32] + [_vp] * 9 + [_sz, _vp]),
    "bgnn_bn_relu_bwd_reduce_f32": (_i32, [_vp, _vp, _i64, _i32, _vp, _i32, _vp, _vp, _sz, _vp]),
    "bgnn_bn_relu_bwd_apply_f32": (_i32, [_vp, _vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp]),
    "bgnn_gatv2_fwd_f32": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _i64, _i32, _vp, _vp, _vp, _vp]),
    "bgnn_gatv2_fwd_ord_f32": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _i64, _i32, _vp, _vp, _vp, _vp]),
    "bgnn_rows_by_degree_workspace_bytes": (_sz, [_i64]),
    "bgnn_rows_by_degree": (_i32, [_vp, _i64, _i32, _vp, _vp, _sz, _vp]),
    "bgnn_adapted_transform_fwd_f32": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "bgnn_adapted_transform_bwd_workspace_bytes": (_sz, [_i32]),
    "bgnn_adapted_transform_bwd_f32": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "bgnn_gatv2_heads_supported": (_i32, [_i32, _i32]),
    "bgnn_gatv2_heads_fwd_f32": (_i32, [_vp] * 7 + [_f32, _i64, _i32, _i32] + [_vp] * 4),
    "bgnn_gatv2_heads_bwd_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32]),
    "bgnn_gatv2_heads_bwd_f32": (_i32, [_vp] * 5 + [_i64] + [_vp] * 5 + [_f32, _i64, _i32, _i32] + [_vp] * 9 + [_sz, _vp]),
    "bgnn_adapted_skinny_supported": (_i32, [_i32, _i32]),
    "bgnn_adapted_skinny_fwd_f32": (_i32, [_vp] * 6 + [_i64, _i32, _i32] + [_vp] * 4),
    "bgnn_adapted_skinny_bwd_workspace_bytes": (_sz, [_i32, _i32]),
    "bgnn_adapted_skinny_bwd_f32": (_i32, [_vp] * 7 + [_i64, _i32, _i32] + [_vp] * 3 + [_sz, _vp]),
    "bgnn_rowpanel_gemm_supported": (_i32, [_i32, _i32, _i32]),
    "bgnn_rowpanel_gemm_act_f32": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _i32, _vp, _i32, _vp]),
    "bgnn_rowpanel_gemm_f32": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _i32, _vp]),
    "bgnn_adapted_transform_bwd_gates_f32": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _sz, _vp]),
    "bgnn_wgrad_gemm_cat_f32": (_i32, [_vp, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _i32, _i64, _vp, _i32, _vp, _vp, _sz, _vp]),
    "bgnn_wgrad_gemm_supported": (_i32, [_i32, _i32, _i32, _i32]),
    "bgnn_wgrad_gemm_workspace_bytes": (_sz, [_i32]),
    "bgnn_wgrad_gemm_f32": (_i32, [_vp, _i32, _i32, _vp, _i32, _i32, _i64, _vp, _i32, _vp, _vp, _sz, _vp]),
    "bgnn_bn_relu_supported": (_i32, [_i32]),
    "bgnn_bn_relu_workspace_bytes": (_sz, [_i32]),
    "bgnn_bn_relu_fwd_f32": (_i32, [_vp, _i64, _i32, _vp, _vp, _f32, _f32, _vp, _vp, _i32, _vp, _vp, _vp, _sz, _vp]),
    "bgnn_bn_relu_apply_f32": (_i32, [_vp, _i64, _i32, _vp, _i32, _vp, _vp]),
    "bgnn_bn_relu_bwd_f32": (_i32, [_vp, _vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _sz, _vp]),
    "bgnn_adapted_skinny_tc_supported": (_i32, [_i32, _i32, _i32]),
    "bgnn_adapted_skinny_heads_tc_fwd_f32": (_i32, [_vp, _i64, _i32, _vp, _vp, _i32, _i32] + [_vp] * 8),
    "bgnn_tf32_planes_f32": (_i32, [_vp, _i32, _i32, _i64, _i64, _i32, _i32, _vp, _vp, _vp]),
    "bgnn_adapted_wide_supported": (_i32, [_i32, _i32]),
    "bgnn_adapted_wide_fwd_f32": (_i32, [_vp, _i64, _i32, _vp, _vp, _i32] + [_vp] * 8),
    "bgnn_adapted_skinny_heads_supported": (_i32, [_i32, _i32, _i32]),
    "bgnn_adapted_skinny_heads_fwd_f32": (_i32, [_vp] * 6 + [_i64, _i32, _i32, _i32] + [_vp] * 4),
    "bgnn_adapted_skinny_heads_bwd_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "bgnn_adapted_skinny_heads_pre_f32": (_i32, [_vp] * 5 + [_i64, _i32, _i32, _vp, _vp, _sz, _vp]),
    "bgnn_adapted_skinny_heads_bwd_f32": (_i32, [_vp] * 8 + [_i64, _i32, _i32, _i32] + [_vp] * 3 + [_sz, _vp]),
    "bgnn_domain_colsum_workspace_bytes": (_sz, [_i32]),
    "bgnn_domain_colsum_f32": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _sz, _vp]),
    "bgnn_gatv2_bwd_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "bgnn_gatv2_bwd_f32": (_i32, [_vp] * 5 + [_i64] + [_vp] * 5 + [_f32, _i64, _i32] + [_vp] * 9 + [_sz, _vp]),
    "bgnn_gatv2_bwd_ord_f32": (_i32, [_vp] * 7 + [_i64] + [_vp] * 5 + [_f32, _i64, _i32] + [_vp] * 9 + [_sz, _vp]),
}

_lib = None


def load():
    """dlopen the library once; raises if it has not been built (python -m bridged_gnn_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "bridged_gnn_b200: %s is missing. Build it with `python -m bridged_gnn_b200.build` "
                "(needs nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(code):
    if code != 0:
        msg = load().bgnn_error_string(code).decode()
        raise RuntimeError("libbgnn_b200 error %d: %s" % (code, msg))


def ptr(t, dtype=None, allow_none=False):
    """Device pointer of a contiguous CUDA tensor (or NULL)."""
    if t is None:
        if allow_none:
            return None
        raise ValueError("tensor required")
    if not t.is_cuda:
        raise RuntimeError("bridged_gnn_b200 runs on CUDA tensors only (got %s); there is no CPU path" % t.device)
    if dtype is not None and t.dtype != dtype:
        raise TypeError("expected %s, got %s" % (dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return t.data_ptr()


def stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ---- instrumentation used by bench.py (off by default; no effect on results) ----------------------
# kernels launched per C-ABI call (hand-written kernels of this library only; CUB's sort/scan inside
# bgnn_edges_to_csr are not counted)
KERNELS_PER_CALL = {"bgnn_knn_cosine_f32": 13, "bgnn_knn_cosine_eps_f32": 13, "bgnn_knn_addrelu_eps_f32": 2, "bgnn_knn_cosine_f32[simt]": 4, "bgnn_knn_addrelu_f32": 2,
                    "bgnn_edges_to_csr": 4, "bgnn_graph_prepare": 8, "bgnn_quantile_f32": 11, "bgnn_edge_validity_f32": 2, "bgnn_spmm_csr_f32": 1, "bgnn_gatv2_fwd_f32": 1, "bgnn_gatv2_bwd_f32": 3,
                    "bgnn_gatv2_fwd_ord_f32": 1, "bgnn_gatv2_bwd_ord_f32": 3, "bgnn_gatv2_fwd_part_f32": 1, "bgnn_gatv2_bwd_part_f32": 3,
                    "bgnn_gatv2_heads_fwd_part_f32": 1, "bgnn_gatv2_heads_bwd_part_f32": 3, "bgnn_spmm_csr_ld_f32": 1,
                    "bgnn_bn_relu_bwd_reduce_f32": 2, "bgnn_bn_relu_bwd_apply_f32": 1, "bgnn_rows_by_degree": 1,
                    "bgnn_adapted_transform_fwd_f32": 1, "bgnn_adapted_transform_bwd_f32": 2,
                    "bgnn_gatv2_heads_fwd_f32": 1, "bgnn_gatv2_heads_bwd_f32": 3, "bgnn_adapted_skinny_fwd_f32": 1, "bgnn_adapted_skinny_bwd_f32": 2, "bgnn_domain_colsum_f32": 2, "bgnn_rowpanel_gemm_f32": 1, "bgnn_rowpanel_gemm_act_f32": 1, "bgnn_tf32_planes_f32": 1, "bgnn_adapted_skinny_heads_tc_fwd_f32": 1, "bgnn_wgrad_gemm_cat_f32": 2, "bgnn_adapted_transform_bwd_gates_f32": 2, "bgnn_adapted_skinny_heads_fwd_f32": 1, "bgnn_adapted_skinny_heads_pre_f32": 2,
                    "bgnn_adapted_skinny_heads_bwd_f32": 2, "bgnn_bn_relu_fwd_f32": 3, "bgnn_bn_relu_apply_f32": 1, "bgnn_bn_relu_bwd_f32": 3, "bgnn_wgrad_gemm_f32": 2,
                    "bgnn_adapted_wide_fwd_f32": 1}
launches = 0          # running count of kernels launched through the C ABI
_timing = None        # None, or {name: [(start_event, end_event), ...]}


def start_timing():
    global _timing
    _timing = {}


def stop_timing():
    """Returns {name: (calls, total_ms)} for the C-ABI calls issued since start_timing(); synchronises."""
    global _timing
    t, _timing = _timing, None
    torch.cuda.synchronize()
    return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in (t or {}).items()}


class call:
    """Context manager around one C-ABI call: counts its kernels and, when timing is on, brackets it with
    CUDA events on the launching stream."""

    def __init__(self, name, tag=None):
        self.name, self.key = name, (tag or name)

    def __enter__(self):
        global launches
        launches += KERNELS_PER_CALL.get(self.key, KERNELS_PER_CALL.get(self.name, 1))
        if _timing is not None:
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.ev[0].record()
        return self

    def __exit__(self, *exc):
        if _timing is not None:
            self.ev[1].record()
            _timing.setdefault(self.key, []).append(self.ev)
        return False
