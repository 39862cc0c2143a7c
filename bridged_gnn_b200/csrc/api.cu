// C ABI of libbgnn_b200 (declared in include/bgnn_b200.h): argument checking, workspace carving and
// the multi-kernel drivers.  No torch types, no allocation, no host synchronisation.
#include "bgnn_b200.h"

#include "common.cuh"
#include "kernels.h"

namespace bgnn {

__global__ void set_int_kernel(int* p, int v) { *p = v; }
__global__ void copy_int_kernel(const int* src, int* dst) { *dst = *src; }

// ---- cosine / add-ReLU kNN driver -------------------------------------------------------------
struct SimtPlan {
  int kc, nsplit, per_split;
};

static SimtPlan simt_plan(int nq, int ndb, int k, bool max_parallel) {
  SimtPlan p;
  p.kc = k + 1;                                    // the (k+1)-th value gives the near-tie gap
  const int tiles = (ndb + 63) / 64;
  const int qtiles = (nq + 63) / 64;
  int ns = max_parallel ? tiles : (2 * kNumSMs + qtiles - 1) / qtiles;
  ns = max(1, min(min(ns, tiles), BGNN_MERGE_MAX_CAND / p.kc));
  const int tiles_per = (tiles + ns - 1) / ns;
  p.per_split = tiles_per * 64;
  p.nsplit = (tiles + tiles_per - 1) / tiles_per;
  return p;
}

struct KnnPlan {
  int algo;        // resolved algorithm
  int passes;      // tf32 tensor-core passes (3 or 1), 0 otherwise
  bool f16;        // single-pass fp16 tensor-core sweep
  int ld;          // row stride of the normalised fp32 planes
  int ldh;         // row stride of the fp16 plane (f16 sweep)
  TcPlan tc;
  int seed_rows;   // sampled db rows of the threshold-seeding sweep (f16 only; 0 = none)
  TcPlan seed_tc;
  SimtPlan simt;   // main sweep (SIMT algo) or exact fallback (tensor-core algos)
  size_t cand_elems;
};

static bool knn_args_ok(int64_t nq, int64_t ndb, int d, int k) {
  return nq >= 0 && ndb >= 1 && nq < (1ll << 31) && ndb < (1ll << 31) && d >= 1 && k >= 1 && k <= ndb && k <= 255;
}

static KnnPlan knn_plan(int64_t nq, int64_t ndb, int d, int k, int algo) {
  KnnPlan p;
  p.algo = algo;
  p.passes = (algo == BGNN_KNN_TC_3XTF32) ? 3 : (algo == BGNN_KNN_TC_1XTF32 ? 1 : 0);
  p.tc.bn = 0;
  p.seed_rows = 0;
  p.f16 = false;
  p.ldh = 0;
  if (algo == BGNN_KNN_TC_F16) {
    p.tc = tc_plan_f16((int)nq, (int)ndb, d, k);
    p.f16 = p.tc.bn != 0;
    if (!p.f16) p.algo = BGNN_KNN_SIMT_F32;                           // d or k too large for the on-chip budget
    if (p.f16) {
      p.seed_rows = knn_seed_rows((int)ndb, k);
      if (p.seed_rows) {
        p.seed_tc = tc_plan_f16((int)nq, p.seed_rows, d, k, kSeedKc);
        if (p.seed_tc.bn == 0) p.seed_rows = 0;
      }
    }
  } else if (p.passes) {
    p.tc = tc_plan((int)nq, (int)ndb, d, k, p.passes);
    if (p.tc.bn == 0) { p.passes = 0; p.algo = BGNN_KNN_SIMT_F32; }   // k too large for the on-chip lists
  }
  p.ld = p.passes ? (d + 31) / 32 * 32 : (p.f16 ? (d + 3) / 4 * 4 : d);
  if (p.f16) p.ldh = (d + 63) / 64 * 64;
  const bool tcpath = p.passes != 0 || p.f16;
  p.simt = simt_plan((int)nq, (int)ndb, k, /*max_parallel=*/tcpath);
  size_t simt_c = (size_t)p.simt.nsplit * nq * p.simt.kc;
  size_t tc_c = tcpath ? (size_t)p.tc.nlists * nq * p.tc.kc : 0;
  p.cand_elems = simt_c > tc_c ? simt_c : tc_c;
  return p;
}

static size_t knn_ws_bytes(const KnnPlan& p, int64_t nq, int64_t ndb) {
  size_t b = 0;
  auto add = [&](size_t n) { b = align_up(b, 256) + n; };
  const int planes = p.passes ? 2 : 1;
  for (int i = 0; i < planes; ++i) { add((size_t)nq * p.ld * 4); add((size_t)ndb * p.ld * 4); }
  if (p.f16) { add((size_t)nq * p.ldh * 2); add((size_t)ndb * p.ldh * 2); }
  if (p.seed_rows) {
    add((size_t)p.seed_rows * p.ldh * 2);
    add((size_t)p.seed_tc.nlists * nq * p.seed_tc.kc * 4);
    add((size_t)p.seed_tc.nlists * nq * p.seed_tc.kc * 4);
    add((size_t)nq * 4);
  }
  add(p.cand_elems * 4);
  add(p.cand_elems * 4);
  add((size_t)nq * 4);   // fallback row list
  add(256);              // fallback row count
  add(knn_exact_rows_workspace_bytes(p.simt.kc - 1));   // per-slice lists of the few-rows exact sweep
  return align_up(b, 256) + 256;
}

}  // namespace bgnn

using namespace bgnn;

extern "C" {

int bgnn_version(void) { return 100; }

const char* bgnn_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case BGNN_ERR_INVALID_ARG: return "invalid argument";
    case BGNN_ERR_WORKSPACE: return "workspace too small";
    case BGNN_ERR_UNSUPPORTED: return "unsupported size or option";
    case BGNN_ERR_DRIVER: return "CUDA driver entry point unavailable";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}

size_t bgnn_knn_cosine_workspace_bytes(int64_t nq, int64_t ndb, int d, int k, int algo) {
  if (!knn_args_ok(nq, ndb, d, k)) return 0;
  return knn_ws_bytes(knn_plan(nq, ndb, d, k, algo), nq, ndb);
}

int bgnn_knn_cosine_f32(const float* q, int64_t nq, const float* db, int64_t ndb, int d, int k, int normalize,
                        int apply_sigmoid, int algo, int64_t* out_idx, float* out_val, float* out_gap,
                        int32_t* out_stats, void* workspace, size_t workspace_bytes, void* stream_) {
  return bgnn_knn_cosine_eps_f32(q, nq, db, ndb, d, k, normalize, apply_sigmoid, algo, nanf(""), out_idx, out_val, out_gap,
                                 nullptr, out_stats, workspace, workspace_bytes, stream_);
}

int bgnn_knn_cosine_eps_f32(const float* q, int64_t nq, const float* db, int64_t ndb, int d, int k, int normalize,
                            int apply_sigmoid, int algo, float eps, int64_t* out_idx, float* out_val, float* out_gap,
                            int32_t* out_count, int32_t* out_stats, void* workspace, size_t workspace_bytes,
                            void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!knn_args_ok(nq, ndb, d, k) || !q || !db || !out_idx || !out_val) return BGNN_ERR_INVALID_ARG;
  if (algo < BGNN_KNN_SIMT_F32 || algo > BGNN_KNN_TC_F16) return BGNN_ERR_INVALID_ARG;
  if (nq == 0) return BGNN_OK;
  const KnnPlan p = knn_plan(nq, ndb, d, k, algo);
  if (workspace_bytes < knn_ws_bytes(p, nq, ndb) || !workspace) return BGNN_ERR_WORKSPACE;
  Workspace w(workspace, workspace_bytes);
  const bool same = (q == db) && (nq == ndb);
  float* qhi = w.take<float>((size_t)nq * p.ld);
  float* dhi = w.take<float>((size_t)ndb * p.ld);
  float *qlo = nullptr, *dlo = nullptr;
  if (p.passes) { qlo = w.take<float>((size_t)nq * p.ld); dlo = w.take<float>((size_t)ndb * p.ld); }
  unsigned short *qh = nullptr, *dh = nullptr;
  if (p.f16) { qh = w.take<unsigned short>((size_t)nq * p.ldh); dh = w.take<unsigned short>((size_t)ndb * p.ldh); }
  unsigned short* sample = nullptr;
  float *seed_val = nullptr, *seed_thr = nullptr;
  int* seed_idx = nullptr;
  if (p.seed_rows) {
    sample = w.take<unsigned short>((size_t)p.seed_rows * p.ldh);
    seed_val = w.take<float>((size_t)p.seed_tc.nlists * nq * p.seed_tc.kc);
    seed_idx = w.take<int>((size_t)p.seed_tc.nlists * nq * p.seed_tc.kc);
    seed_thr = w.take<float>((size_t)nq);
  }
  float* cand_val = w.take<float>(p.cand_elems);
  int* cand_idx = w.take<int>(p.cand_elems);
  int* fb_rows = w.take<int>((size_t)nq);
  int* fb_count = w.take<int>(64);
  char* xr_ws = w.take<char>(knn_exact_rows_workspace_bytes(k));
  if (!w.ok()) return BGNN_ERR_WORKSPACE;
  int rc;
  // prologue: unit rows (plus the tf32 hi/lo split or the fp16 rounding for the tensor-core sweeps)
  if (p.f16) {
    if ((rc = launch_normalize_f16(db, ndb, d, p.ld, p.ldh, normalize, dhi, dh, stream)) != BGNN_OK) return rc;
    if (same) { qhi = dhi; qh = dh; }
    else if ((rc = launch_normalize_f16(q, nq, d, p.ld, p.ldh, normalize, qhi, qh, stream)) != BGNN_OK) return rc;
  } else {
    if ((rc = launch_normalize_split(db, ndb, d, p.ld, normalize, dhi, dlo, stream)) != BGNN_OK) return rc;
    if (same) { qhi = dhi; qlo = dlo; }
    else if ((rc = launch_normalize_split(q, nq, d, p.ld, normalize, qhi, qlo, stream)) != BGNN_OK) return rc;
  }

  if (!p.passes && !p.f16) {
    rc = launch_knn_simt(BGNN_PAIR_DOT, qhi, nullptr, (int)nq, dhi, nullptr, (int)ndb, d, p.ld, nullptr, 0.f,
                         apply_sigmoid, p.simt.kc, p.simt.nsplit, p.simt.per_split, nullptr, nullptr, 0, cand_val,
                         cand_idx, stream);
    if (rc != BGNN_OK) return rc;
    rc = launch_knn_merge(cand_val, cand_idx, p.simt.nsplit, p.simt.kc, (int)nq, k, 0, nullptr, nullptr, nullptr,
                          nullptr, d, p.ld, apply_sigmoid, -1.f, nullptr, nullptr, nullptr, 0, (long long*)out_idx, out_val,
                          out_gap, nullptr, nullptr, eps, out_count, stream);
    if (rc != BGNN_OK) return rc;
    if (out_stats) { set_int_kernel<<<1, 1, 0, stream>>>(out_stats, 0); BGNN_LAUNCH_CHECK(); }
    return BGNN_OK;
  }

  // tensor-core sweep nominates, merge re-scores + certifies, CUDA-core sweep redoes uncertified rows
  set_int_kernel<<<1, 1, 0, stream>>>(fb_count, 0);
  BGNN_LAUNCH_CHECK();
  if (p.f16) {
    if (p.seed_rows) {
      rc = launch_knn_seed_f16(qh, (int)nq, dh, (int)ndb, p.ldh, p.seed_rows, p.seed_tc, sample, seed_val, seed_idx,
                               seed_thr, stream);
      if (rc != BGNN_OK) return rc;
    }
    rc = launch_knn_cosine_f16(qh, (int)nq, dh, (int)ndb, p.ldh, p.tc, seed_thr, cand_val, cand_idx, stream);
  } else rc = launch_knn_cosine_tc(qhi, qlo, (int)nq, dhi, dlo, (int)ndb, p.ld, p.passes, p.tc, cand_val, cand_idx, stream);
  if (rc != BGNN_OK) return rc;
  // error bound of the approximate dot product of two unit rows (see DESIGN.md "kNN exactness")
  const float delta = p.f16 ? (9.7657e-4f + 5.97e-8f * sqrtf((float)d) + 4.0e-6f) : ((p.passes == 3) ? 3.0e-5f : 2.0e-3f);
  rc = launch_knn_merge(cand_val, cand_idx, p.tc.nlists, p.tc.kc, (int)nq, k, 1, qhi, qlo, dhi, dlo, p.ld, p.ld,
                        apply_sigmoid, delta, seed_thr, nullptr, nullptr, 0, (long long*)out_idx, out_val, out_gap, fb_rows,
                        fb_count, eps, out_count, stream);
  if (rc != BGNN_OK) return rc;
  // uncertified rows, exactly: a handful -> one db slice per CTA (knn_exact_rows); more -> the tiled sweep + merge.
  // The count lives on the device, so both are launched and each returns at once when it is not its case.
  const int few = knn_exact_rows_max((int)ndb, p.ld, k);
  rc = launch_knn_exact_rows(qhi, qlo, dhi, dlo, (int)ndb, p.ld, p.ld, apply_sigmoid, k, fb_rows, fb_count,
                             (long long*)out_idx, out_val, out_gap, eps, out_count, xr_ws,
                             knn_exact_rows_workspace_bytes(k), stream);
  if (rc != BGNN_OK) return rc;
  rc = launch_knn_simt(BGNN_PAIR_DOT, qhi, qlo, (int)nq, dhi, dlo, (int)ndb, p.ld, p.ld, nullptr, 0.f, apply_sigmoid,
                       p.simt.kc, p.simt.nsplit, p.simt.per_split, fb_rows, fb_count, few, cand_val, cand_idx, stream);
  if (rc != BGNN_OK) return rc;
  rc = launch_knn_merge(cand_val, cand_idx, p.simt.nsplit, p.simt.kc, (int)nq, k, 0, nullptr, nullptr, nullptr,
                        nullptr, p.ld, p.ld, apply_sigmoid, -1.f, nullptr, fb_rows, fb_count, few, (long long*)out_idx,
                        out_val, out_gap, nullptr, nullptr, eps, out_count, stream);
  if (rc != BGNN_OK) return rc;
  if (out_stats) { copy_int_kernel<<<1, 1, 0, stream>>>(fb_count, out_stats); BGNN_LAUNCH_CHECK(); }
  return BGNN_OK;
}

size_t bgnn_knn_addrelu_workspace_bytes(int64_t nq, int64_t ndb, int h, int k) {
  if (!knn_args_ok(nq, ndb, h, k)) return 0;
  const SimtPlan p = simt_plan((int)nq, (int)ndb, k, false);
  return align_up((size_t)p.nsplit * nq * p.kc * 4, 256) * 2 + 512;
}

int bgnn_knn_addrelu_f32(const float* Uq, int64_t nq, const float* Udb, int64_t ndb, int h, const float* w2, float b2,
                         int k, int apply_sigmoid, int64_t* out_idx, float* out_val, float* out_gap,
                         void* workspace, size_t workspace_bytes, void* stream_) {
  return bgnn_knn_addrelu_eps_f32(Uq, nq, Udb, ndb, h, w2, b2, k, apply_sigmoid, nanf(""), out_idx, out_val, out_gap, nullptr,
                                  workspace, workspace_bytes, stream_);
}

int bgnn_knn_addrelu_eps_f32(const float* Uq, int64_t nq, const float* Udb, int64_t ndb, int h, const float* w2, float b2,
                             int k, int apply_sigmoid, float eps, int64_t* out_idx, float* out_val, float* out_gap,
                             int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!knn_args_ok(nq, ndb, h, k) || !Uq || !Udb || !w2 || !out_idx || !out_val) return BGNN_ERR_INVALID_ARG;
  if (nq == 0) return BGNN_OK;
  const SimtPlan p = simt_plan((int)nq, (int)ndb, k, false);
  if (!workspace || workspace_bytes < bgnn_knn_addrelu_workspace_bytes(nq, ndb, h, k)) return BGNN_ERR_WORKSPACE;
  Workspace w(workspace, workspace_bytes);
  const size_t ce = (size_t)p.nsplit * nq * p.kc;
  float* cand_val = w.take<float>(ce);
  int* cand_idx = w.take<int>(ce);
  if (!w.ok()) return BGNN_ERR_WORKSPACE;
  int rc = launch_knn_simt(BGNN_PAIR_ADDRELU, Uq, nullptr, (int)nq, Udb, nullptr, (int)ndb, h, h, w2, b2, apply_sigmoid,
                           p.kc, p.nsplit, p.per_split, nullptr, nullptr, 0, cand_val, cand_idx, stream);
  if (rc != BGNN_OK) return rc;
  return launch_knn_merge(cand_val, cand_idx, p.nsplit, p.kc, (int)nq, k, 0, nullptr, nullptr, nullptr, nullptr, h, h,
                          apply_sigmoid, -1.f, nullptr, nullptr, nullptr, 0, (long long*)out_idx, out_val, out_gap, nullptr,
                          nullptr, eps, out_count, stream);
}

size_t bgnn_quantile_workspace_bytes(void) { return quantile_workspace_bytes(); }

int bgnn_quantile_f32(const float* v, int64_t n, int64_t rank_lo, float weight, float* out3, void* workspace,
                      size_t workspace_bytes, void* stream) {
  if (!v || !out3 || n <= 0 || rank_lo < 0 || rank_lo >= n || !(weight >= 0.f && weight <= 1.f)) return BGNN_ERR_INVALID_ARG;
  if (!workspace || workspace_bytes < quantile_workspace_bytes()) return BGNN_ERR_WORKSPACE;
  return launch_quantile(v, n, rank_lo, weight, out3, workspace, workspace_bytes, (cudaStream_t)stream);
}

int bgnn_edge_validity_f32(const int64_t* e0, const int64_t* e1, int64_t e, const float* e_sim, const float* thr_conf,
                           const int64_t* pred_a, const int64_t* y_a, const int64_t* pred_b, const int64_t* y_b,
                           const uint8_t* gate_a, const uint8_t* gate_b, const float* x_a, const float* x_b, int d,
                           float thres_feat_sim, uint8_t* keep, int64_t* counts, void* stream) {
  if (e < 0 || d <= 0 || !counts) return BGNN_ERR_INVALID_ARG;
  if (e > 0 && (!e0 || !e1 || !e_sim || !pred_a || !y_a || !pred_b || !y_b || !x_a || !x_b || !keep)) return BGNN_ERR_INVALID_ARG;
  return launch_edge_validity((const long long*)e0, (const long long*)e1, e, e_sim, thr_conf, (const long long*)pred_a,
                              (const long long*)y_a, (const long long*)pred_b, (const long long*)y_b, gate_a, gate_b, x_a, x_b,
                              d, thres_feat_sim, keep, (long long*)counts, (cudaStream_t)stream);
}

size_t bgnn_edges_to_csr_workspace_bytes(int64_t e) { return e < 0 ? 0 : csr_build_workspace_bytes(e); }

int bgnn_edges_to_csr(const int64_t* src, const int64_t* dst, int64_t e, int64_t n, int dedup, int32_t* rowptr,
                      int32_t* col, int64_t* perm, int64_t* e_out, void* workspace, size_t workspace_bytes,
                      void* stream) {
  if (e < 0 || n < 0 || !rowptr || !e_out || (e > 0 && (!src || !dst || !col))) return BGNN_ERR_INVALID_ARG;
  if (e > 0 && (!workspace || workspace_bytes < csr_build_workspace_bytes(e))) return BGNN_ERR_WORKSPACE;
  return launch_edges_to_csr((const long long*)src, (const long long*)dst, e, n, dedup, rowptr, col, (long long*)perm,
                             (long long*)e_out, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t bgnn_graph_prepare_workspace_bytes(int64_t e, int64_t n, int rewrite_self_loops) {
  return (e < 0 || n < 0) ? 0 : graph_prepare_workspace_bytes(e, n, rewrite_self_loops);
}

int bgnn_graph_prepare(const int64_t* src, const int64_t* dst, int64_t e, int64_t n, int rewrite_self_loops,
                       int32_t* rowptr, int32_t* col, int32_t* t_rowptr, int32_t* t_col, int32_t* csr_to_csc,
                       int32_t* order, int32_t* t_order, int64_t* e_out, void* workspace, size_t workspace_bytes,
                       void* stream) {
  if (e < 0 || n < 0 || !rowptr || !e_out || (e > 0 && (!src || !dst))) return BGNN_ERR_INVALID_ARG;
  const int64_t m = e + (rewrite_self_loops ? n : 0);
  if (m > 0 && !col) return BGNN_ERR_INVALID_ARG;
  if (m > 0 && (!workspace || workspace_bytes < graph_prepare_workspace_bytes(e, n, rewrite_self_loops))) return BGNN_ERR_WORKSPACE;
  return launch_graph_prepare((const long long*)src, (const long long*)dst, e, n, rewrite_self_loops, rowptr, col, t_rowptr,
                              t_col, csr_to_csc, order, t_order, (long long*)e_out, workspace, workspace_bytes,
                              (cudaStream_t)stream);
}

int bgnn_spmm_csr_ld_f32(const int32_t* rowptr, const int32_t* col, const float* edge_w, const float* gather_scale,
                         const float* out_scale, const float* X, int64_t ldx, int64_t n_rows, int f, int reduce_mean,
                         float* Y, int64_t ldy, void* stream) {
  if (n_rows < 0 || f < 0 || ldx < f || ldy < f || (n_rows > 0 && f > 0 && (!rowptr || !X || !Y))) return BGNN_ERR_INVALID_ARG;
  return launch_spmm_csr(rowptr, col, edge_w, gather_scale, out_scale, X, ldx, n_rows, f, reduce_mean, Y, ldy,
                         (cudaStream_t)stream);
}

int bgnn_spmm_csr_f32(const int32_t* rowptr, const int32_t* col, const float* edge_w, const float* gather_scale,
                      const float* out_scale, const float* X, int64_t n_rows, int f, int reduce_mean, float* Y,
                      void* stream) {
  return bgnn_spmm_csr_ld_f32(rowptr, col, edge_w, gather_scale, out_scale, X, f, n_rows, f, reduce_mean, Y, f, stream);
}

int bgnn_gatv2_fwd_part_f32(const int32_t* rowptr, const int32_t* col, const int32_t* row_order, const uint8_t* dst_is_src,
                            const float* Hs, const float* Ht, const float* af_t2s, const float* af_s2t, float slope,
                            int64_t n_rows, int64_t row_off, int c, float* out, float* row_max, float* row_sum, float* score,
                            void* stream) {
  if (n_rows < 0 || row_off < 0 || c <= 0) return BGNN_ERR_INVALID_ARG;
  if (n_rows > 0 && (!rowptr || !col || !dst_is_src || !Hs || !Ht || !af_t2s || !af_s2t || !out)) return BGNN_ERR_INVALID_ARG;
  return launch_gatv2_fwd(rowptr, col, row_order, dst_is_src, Hs, Ht, af_t2s, af_s2t, slope, n_rows, row_off, c, out, row_max,
                          row_sum, score, 0, (cudaStream_t)stream);
}

int bgnn_gatv2_fwd_ord_f32(const int32_t* rowptr, const int32_t* col, const int32_t* row_order, const uint8_t* dst_is_src,
                           const float* Hs, const float* Ht, const float* af_t2s, const float* af_s2t, float slope,
                           int64_t n, int c, float* out, float* row_max, float* row_sum, void* stream) {
  return bgnn_gatv2_fwd_part_f32(rowptr, col, row_order, dst_is_src, Hs, Ht, af_t2s, af_s2t, slope, n, 0, c, out, row_max,
                                 row_sum, nullptr, stream);
}

int bgnn_gatv2_fwd_f32(const int32_t* rowptr, const int32_t* col, const uint8_t* dst_is_src, const float* Hs,
                       const float* Ht, const float* af_t2s, const float* af_s2t, float slope, int64_t n, int c,
                       float* out, float* row_max, float* row_sum, void* stream) {
  return bgnn_gatv2_fwd_ord_f32(rowptr, col, nullptr, dst_is_src, Hs, Ht, af_t2s, af_s2t, slope, n, c, out, row_max,
                                row_sum, stream);
}

size_t bgnn_rows_by_degree_workspace_bytes(int64_t n) { return n < 0 ? 0 : rows_by_degree_workspace_bytes(n); }

int bgnn_rows_by_degree(const int32_t* rowptr, int64_t n, int min_degree, int32_t* order, void* workspace,
                        size_t workspace_bytes, void* stream) {
  if (n < 0 || min_degree < 0 || (n > 0 && (!rowptr || !order || !workspace))) return BGNN_ERR_INVALID_ARG;
  if (n > 0 && workspace_bytes < rows_by_degree_workspace_bytes(n)) return BGNN_ERR_WORKSPACE;
  return launch_rows_by_degree(rowptr, n, min_degree, order, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t bgnn_gatv2_bwd_workspace_bytes(int64_t n, int64_t e, int c) {
  return (n < 0 || e < 0 || c <= 0) ? 0 : gatv2_bwd_workspace_bytes(n, e, c);
}

int bgnn_gatv2_bwd_part_f32(const int32_t* rowptr, const int32_t* col, const int32_t* t_rowptr, const int32_t* t_col,
                            const int32_t* csr_to_csc, const int32_t* row_order, const int32_t* t_row_order, int64_t e,
                            const uint8_t* dst_is_src, const float* Hs, const float* Ht, const float* af_t2s,
                            const float* af_s2t, float slope, int64_t n_rows, int64_t row_off, int64_t n_src, int c,
                            const float* out, const float* row_max, const float* row_sum, const float* score,
                            const float* gout, float* gHs, float* gHt, float* g_af_t2s, float* g_af_s2t, void* workspace,
                            size_t workspace_bytes, void* stream) {
  if (n_rows < 0 || n_src < 0 || row_off < 0 || row_off + n_rows > n_src || e < 0 || c <= 0) return BGNN_ERR_INVALID_ARG;
  if (n_src > 0 && (!t_rowptr || !dst_is_src || !Hs || !Ht || !af_t2s || !af_s2t || !g_af_t2s || !g_af_s2t || (!gHs && !gHt)))
    return BGNN_ERR_INVALID_ARG;
  if (n_rows > 0 && (!rowptr || !out || !row_max || !row_sum || !gout)) return BGNN_ERR_INVALID_ARG;
  if (e > 0 && (!col || !t_col || !csr_to_csc)) return BGNN_ERR_INVALID_ARG;
  const size_t need = gatv2_bwd_workspace_bytes(n_src, e, c);
  if (need == 0) return BGNN_ERR_UNSUPPORTED;
  if (n_src > 0 && (!workspace || workspace_bytes < need)) return BGNN_ERR_WORKSPACE;
  return launch_gatv2_bwd(rowptr, col, t_rowptr, t_col, csr_to_csc, row_order, t_row_order, e, dst_is_src, Hs, Ht, af_t2s,
                          af_s2t, slope, n_rows, row_off, n_src, c, out, row_max, row_sum, score, gout, gHs, gHt, g_af_t2s,
                          g_af_s2t, workspace, workspace_bytes, (cudaStream_t)stream);
}

int bgnn_gatv2_bwd_ord_f32(const int32_t* rowptr, const int32_t* col, const int32_t* t_rowptr, const int32_t* t_col,
                           const int32_t* csr_to_csc, const int32_t* row_order, const int32_t* t_row_order, int64_t e,
                           const uint8_t* dst_is_src, const float* Hs, const float* Ht, const float* af_t2s,
                           const float* af_s2t, float slope, int64_t n, int c, const float* out, const float* row_max,
                           const float* row_sum, const float* gout, float* gHs, float* gHt, float* g_af_t2s,
                           float* g_af_s2t, void* workspace, size_t workspace_bytes, void* stream) {
  if (n > 0 && (!gHs || !gHt)) return BGNN_ERR_INVALID_ARG;
  return bgnn_gatv2_bwd_part_f32(rowptr, col, t_rowptr, t_col, csr_to_csc, row_order, t_row_order, e, dst_is_src, Hs, Ht,
                                 af_t2s, af_s2t, slope, n, 0, n, c, out, row_max, row_sum, nullptr, gout, gHs, gHt, g_af_t2s,
                                 g_af_s2t, workspace, workspace_bytes, stream);
}

int bgnn_gatv2_bwd_f32(const int32_t* rowptr, const int32_t* col, const int32_t* t_rowptr, const int32_t* t_col,
                       const int32_t* csr_to_csc, int64_t e, const uint8_t* dst_is_src, const float* Hs,
                       const float* Ht, const float* af_t2s, const float* af_s2t, float slope, int64_t n, int c,
                       const float* out, const float* row_max, const float* row_sum, const float* gout, float* gHs,
                       float* gHt, float* g_af_t2s, float* g_af_s2t, void* workspace, size_t workspace_bytes,
                       void* stream) {
  return bgnn_gatv2_bwd_ord_f32(rowptr, col, t_rowptr, t_col, csr_to_csc, nullptr, nullptr, e, dst_is_src, Hs, Ht, af_t2s,
                                af_s2t, slope, n, c, out, row_max, row_sum, gout, gHs, gHt, g_af_t2s, g_af_s2t, workspace,
                                workspace_bytes, stream);
}

int bgnn_adapted_transform_fwd_f32(const float* P, const uint8_t* is_src, const float* wd, const float* kg,
                                   const float* bias, int64_t n, int c, float* Hs, float* Ht, float* gates, void* stream) {
  if (n < 0 || c <= 0) return BGNN_ERR_INVALID_ARG;
  if (n > 0 && (!P || !is_src || !wd || !kg || !Hs || !Ht || !gates)) return BGNN_ERR_INVALID_ARG;
  return launch_adapted_transform_fwd(P, is_src, wd, kg, bias, n, c, Hs, Ht, gates, (cudaStream_t)stream);
}

size_t bgnn_adapted_transform_bwd_workspace_bytes(int c) { return c <= 0 ? 0 : adapted_transform_bwd_workspace_bytes(c); }

int bgnn_adapted_transform_bwd_f32(const float* gHs, const float* gHt, const float* gates, const uint8_t* is_src,
                                   const float* wd, int64_t n, int c, int ldp, float* gP, float* g_wd_kg, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  if (n < 0 || c <= 0 || ldp < 2 * c + 2) return BGNN_ERR_INVALID_ARG;
  if (n > 0 && (!gHs || !gHt || !gates || !is_src || !wd || !gP || !g_wd_kg || !workspace)) return BGNN_ERR_INVALID_ARG;
  return launch_adapted_transform_bwd(gHs, gHt, gates, is_src, wd, n, c, ldp, 1, gP, g_wd_kg, workspace, workspace_bytes,
                                      (cudaStream_t)stream);
}

int bgnn_gatv2_heads_supported(int heads, int c) { return gatv2_heads_supported(heads, c) ? 1 : 0; }

int bgnn_gatv2_heads_fwd_part_f32(const int32_t* rowptr, const int32_t* col, const uint8_t* dst_is_src, const float* Hs,
                                  const float* Ht, const float* af_t2s, const float* af_s2t, float slope, int64_t n_rows,
                                  int64_t row_off, int heads, int c, float* out, float* row_max, float* row_sum, void* stream) {
  if (n_rows < 0 || row_off < 0 || c <= 0 || heads <= 0) return BGNN_ERR_INVALID_ARG;
  if (n_rows > 0 && (!rowptr || !col || !dst_is_src || !Hs || !Ht || !af_t2s || !af_s2t || !out)) return BGNN_ERR_INVALID_ARG;
  return launch_gatv2_heads_fwd(rowptr, col, dst_is_src, Hs, Ht, af_t2s, af_s2t, slope, n_rows, row_off, heads, c, out, row_max,
                                row_sum, (cudaStream_t)stream);
}

int bgnn_gatv2_heads_fwd_f32(const int32_t* rowptr, const int32_t* col, const uint8_t* dst_is_src, const float* Hs,
                             const float* Ht, const float* af_t2s, const float* af_s2t, float slope, int64_t n, int heads,
                             int c, float* out, float* row_max, float* row_sum, void* stream) {
  return bgnn_gatv2_heads_fwd_part_f32(rowptr, col, dst_is_src, Hs, Ht, af_t2s, af_s2t, slope, n, 0, heads, c, out, row_max,
                                       row_sum, stream);
}

size_t bgnn_gatv2_heads_bwd_workspace_bytes(int64_t n, int64_t e, int heads, int c) {
  return (n < 0 || e < 0 || c <= 0 || heads <= 0) ? 0 : gatv2_heads_bwd_workspace_bytes(n, e, heads, c);
}

int bgnn_gatv2_heads_bwd_part_f32(const int32_t* rowptr, const int32_t* col, const int32_t* t_rowptr, const int32_t* t_col,
                                  const int32_t* csr_to_csc, int64_t e, const uint8_t* dst_is_src, const float* Hs,
                                  const float* Ht, const float* af_t2s, const float* af_s2t, float slope, int64_t n_rows,
                                  int64_t row_off, int64_t n_src, int heads, int c, const float* out, const float* row_max,
                                  const float* row_sum, const float* gout, float* gHs, float* gHt, float* g_af_t2s,
                                  float* g_af_s2t, void* workspace, size_t workspace_bytes, void* stream) {
  if (n_rows < 0 || n_src < 0 || row_off < 0 || row_off + n_rows > n_src || e < 0 || c <= 0 || heads <= 0) return BGNN_ERR_INVALID_ARG;
  if (n_src > 0 && (!t_rowptr || !dst_is_src || !Hs || !Ht || !af_t2s || !af_s2t || !g_af_t2s || !g_af_s2t || (!gHs && !gHt)))
    return BGNN_ERR_INVALID_ARG;
  if (n_rows > 0 && (!rowptr || !out || !row_max || !row_sum || !gout)) return BGNN_ERR_INVALID_ARG;
  if (e > 0 && (!col || !t_col || !csr_to_csc)) return BGNN_ERR_INVALID_ARG;
  if (!gatv2_heads_supported(heads, c)) return BGNN_ERR_UNSUPPORTED;
  if (n_src > 0 && (!workspace || workspace_bytes < gatv2_heads_bwd_workspace_bytes(n_src, e, heads, c))) return BGNN_ERR_WORKSPACE;
  return launch_gatv2_heads_bwd(rowptr, col, t_rowptr, t_col, csr_to_csc, e, dst_is_src, Hs, Ht, af_t2s, af_s2t, slope, n_rows,
                                row_off, n_src, heads, c, out, row_max, row_sum, gout, gHs, gHt, g_af_t2s, g_af_s2t, workspace,
                                workspace_bytes, (cudaStream_t)stream);
}

int bgnn_gatv2_heads_bwd_f32(const int32_t* rowptr, const int32_t* col, const int32_t* t_rowptr, const int32_t* t_col,
                             const int32_t* csr_to_csc, int64_t e, const uint8_t* dst_is_src, const float* Hs,
                             const float* Ht, const float* af_t2s, const float* af_s2t, float slope, int64_t n, int heads,
                             int c, const float* out, const float* row_max, const float* row_sum, const float* gout,
                             float* gHs, float* gHt, float* g_af_t2s, float* g_af_s2t, void* workspace,
                             size_t workspace_bytes, void* stream) {
  if (n > 0 && (!gHs || !gHt)) return BGNN_ERR_INVALID_ARG;
  return bgnn_gatv2_heads_bwd_part_f32(rowptr, col, t_rowptr, t_col, csr_to_csc, e, dst_is_src, Hs, Ht, af_t2s, af_s2t, slope,
                                       n, 0, n, heads, c, out, row_max, row_sum, gout, gHs, gHt, g_af_t2s, g_af_s2t, workspace,
                                       workspace_bytes, stream);
}

int bgnn_adapted_skinny_supported(int c, int d) { return adapted_skinny_supported(c, d) ? 1 : 0; }

int bgnn_adapted_skinny_fwd_f32(const float* x, const uint8_t* is_src, const float* wcat, const float* bias,
                                const float* wd, const float* kg, int64_t n, int d, int c, float* Hs, float* Ht,
                                float* gates, void* stream) {
  if (n < 0 || c <= 0 || d <= 0) return BGNN_ERR_INVALID_ARG;
  if (n > 0 && (!x || !is_src || !wcat || !wd || !kg || !Hs || !Ht || !gates)) return BGNN_ERR_INVALID_ARG;
  return launch_adapted_skinny_fwd(x, is_src, wcat, bias, wd, kg, n, d, c, 1, Hs, Ht, gates, (cudaStream_t)stream);
}

size_t bgnn_adapted_skinny_bwd_workspace_bytes(int c, int d) {
  return (c <= 0 || d <= 0) ? 0 : adapted_skinny_bwd_workspace_bytes(c, d, 1);
}

int bgnn_adapted_skinny_bwd_f32(const float* x, const uint8_t* is_src, const float* wcat, const float* wd,
                                const float* gates, const float* gHs, const float* gHt, int64_t n, int d, int c,
                                float* gx, float* red, void* workspace, size_t workspace_bytes, void* stream) {
  if (n < 0 || c <= 0 || d <= 0) return BGNN_ERR_INVALID_ARG;
  if (n > 0 && (!x || !is_src || !wcat || !wd || !gates || !gHs || !gHt || !gx || !red || !workspace))
    return BGNN_ERR_INVALID_ARG;
  return launch_adapted_skinny_bwd(x, is_src, wcat, wd, gates, gHs, gHt, nullptr, n, d, c, 1, gx, red, workspace,
                                   workspace_bytes, (cudaStream_t)stream);
}

size_t bgnn_domain_colsum_workspace_bytes(int d) { return d <= 0 ? 0 : domain_colsum_workspace_bytes(d); }

int bgnn_domain_colsum_f32(const float* x, const uint8_t* is_src, int64_t n, int d, float* sums, void* workspace,
                           size_t workspace_bytes, void* stream) {
  if (n < 0 || d <= 0 || !sums || !workspace || (n > 0 && (!x || !is_src))) return BGNN_ERR_INVALID_ARG;
  return launch_domain_colsum(x, is_src, n, d, sums, workspace, workspace_bytes, (cudaStream_t)stream);
}

int bgnn_rowpanel_gemm_supported(int k, int ld_a, int no) { return rowpanel_gemm_supported(k, ld_a, no) ? 1 : 0; }

int bgnn_rowpanel_gemm_act_f32(const float* A, int64_t n, int k, int ld_a, const float* b_hi, const float* b_lo,
                               const float* scale, const float* bias, int act, const float* res, int ld_res, int no,
                               float* Y, int ldy, void* stream) {
  if (n < 0 || k <= 0 || no <= 0 || ld_a < k || ldy < no || act < 0 || act > 2 || (res && ld_res < no)) return BGNN_ERR_INVALID_ARG;
  if (n > 0 && (!A || !b_hi || !b_lo || !Y)) return BGNN_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(b_hi) | reinterpret_cast<uintptr_t>(b_lo)) & 15)
    return BGNN_ERR_INVALID_ARG;
  return launch_rowpanel_gemm(A, n, k, ld_a, b_hi, b_lo, scale, bias, act, res, ld_res, no, Y, ldy, (cudaStream_t)stream);
}

int bgnn_rowpanel_gemm_f32(const float* A, int64_t n, int k, int ld_a, const float* b_hi, const float* b_lo,
                           const float* bias, int no, float* Y, int ldy, void* stream) {
  return bgnn_rowpanel_gemm_act_f32(A, n, k, ld_a, b_hi, b_lo, nullptr, bias, 0, nullptr, 0, no, Y, ldy, stream);
}

int bgnn_adapted_wide_supported(int c, int d) { return adapted_wide_supported(c, d) ? 1 : 0; }

int bgnn_adapted_wide_fwd_f32(const float* x, int64_t n, int d, const float* wcat_hi, const float* wcat_lo, int c,
                              const uint8_t* is_src, const float* wd, const float* kg, const float* bias, float* Hs,
                              float* Ht, float* gates, void* stream) {
  if (n < 0 || c <= 0 || d <= 0) return BGNN_ERR_INVALID_ARG;
  if (n > 0 && (!x || !wcat_hi || !wcat_lo || !is_src || !wd || !kg || !Hs || !Ht || !gates)) return BGNN_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wcat_hi) | reinterpret_cast<uintptr_t>(wcat_lo) |
       reinterpret_cast<uintptr_t>(Hs) | reinterpret_cast<uintptr_t>(Ht)) & 15)
    return BGNN_ERR_INVALID_ARG;
  return launch_adapted_wide_fwd(x, n, d, wcat_hi, wcat_lo, c, is_src, wd, kg, bias, Hs, Ht, gates, (cudaStream_t)stream);
}

int bgnn_wgrad_gemm_supported(int d, int ld_x, int no, int ld_g) { return wgrad_gemm_supported(d, ld_x, no, ld_g) ? 1 : 0; }

size_t bgnn_wgrad_gemm_workspace_bytes(int no) { return no <= 0 ? 0 : wgrad_gemm_workspace_bytes(no); }

int bgnn_wgrad_gemm_f32(const float* G, int ld_g, int no, const float* X, int ld_x, int d, int64_t n, float* W, int ldw,
                        float* colsum, void* workspace, size_t workspace_bytes, void* stream) {
  if (n < 0 || no <= 0 || d <= 0 || ld_g < no || ld_x < d || ldw < d || !W || !workspace) return BGNN_ERR_INVALID_ARG;
  if (n > 0 && (!G || !X)) return BGNN_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(X)) & 15) return BGNN_ERR_INVALID_ARG;
  return launch_wgrad_gemm(G, ld_g, no, X, ld_x, d, n, W, ldw, colsum, workspace, workspace_bytes, (cudaStream_t)stream);
}

int bgnn_bn_relu_supported(int c) { return bn_relu_supported(c) ? 1 : 0; }

size_t bgnn_bn_relu_workspace_bytes(int c) { return c <= 0 ? 0 : bn_relu_workspace_bytes(c); }

int bgnn_bn_relu_fwd_f32(const float* x, int64_t n, int c, const float* weight, const float* bias, float eps, float momentum,
                         float* running_mean, float* running_var, int relu, float* y, float* stats, void* workspace,
                         size_t workspace_bytes, void* stream) {
  if (n < 0 || c <= 0 || !stats || !workspace) return BGNN_ERR_INVALID_ARG;
  if (n > 0 && !x) return BGNN_ERR_INVALID_ARG;     // y may be NULL: statistics only
  return launch_bn_relu_fwd(x, n, c, weight, bias, eps, momentum, running_mean, running_var, relu, y, stats, workspace,
                            workspace_bytes, (cudaStream_t)stream);
}

int bgnn_bn_relu_apply_f32(const float* x, int64_t n, int c, const float* stats, int relu, float* y, void* stream) {
  if (n < 0 || c <= 0 || !stats) return BGNN_ERR_INVALID_ARG;
  if (n > 0 && (!x || !y)) return BGNN_ERR_INVALID_ARG;
  return launch_bn_relu_apply(x, n, c, stats, relu, y, (cudaStream_t)stream);
}

int bgnn_bn_relu_bwd_reduce_f32(const float* gy, const float* x, int64_t n, int c, const float* stats, int relu, float* gwb,
                                void* workspace, size_t workspace_bytes, void* stream) {
  if (n < 0 || c <= 0 || !stats || !gwb || !workspace) return BGNN_ERR_INVALID_ARG;
  if (n > 0 && (!gy || !x)) return BGNN_ERR_INVALID_ARG;
  return launch_bn_relu_bwd_reduce(gy, x, n, c, stats, relu, gwb, workspace, workspace_bytes, (cudaStream_t)stream);
}

int bgnn_bn_relu_bwd_apply_f32(const float* gy, const float* x, int64_t n, int c, const float* stats, int relu,
                               const float* coef, float* gx, void* stream) {
  if (n < 0 || c <= 0 || !stats || !coef) return BGNN_ERR_INVALID_ARG;
  if (n > 0 && (!gy || !x || !gx)) return BGNN_ERR_INVALID_ARG;
  return launch_bn_relu_bwd_apply(gy, x, n, c, stats, relu, coef, gx, (cudaStream_t)stream);
}

int bgnn_bn_relu_bwd_f32(const float* gy, const float* x, int64_t n, int c, const float* stats, int relu, float* gx,
                         float* gwb, void* workspace, size_t workspace_bytes, void* stream) {
  if (n < 0 || c <= 0 || !stats || !gwb || !workspace) return BGNN_ERR_INVALID_ARG;
  if (n > 0 && (!gy || !x || !gx)) return BGNN_ERR_INVALID_ARG;
  return launch_bn_relu_bwd(gy, x, n, c, stats, relu, gx, gwb, workspace, workspace_bytes, (cudaStream_t)stream);
}

int bgnn_adapted_skinny_heads_supported(int c, int d, int heads) { return adapted_skinny_heads_supported(c, d, heads) ? 1 : 0; }

int bgnn_adapted_skinny_heads_fwd_f32(const float* x, const uint8_t* is_src, const float* wcat, const float* bias,
                                      const float* wd, const float* kg, int64_t n, int d, int c, int heads, float* Hs,
                                      float* Ht, float* gates, void* stream) {
  if (n < 0 || c <= 0 || d <= 0 || heads <= 0) return BGNN_ERR_INVALID_ARG;
  if (n > 0 && (!x || !is_src || !wcat || !wd || !kg || !Hs || !Ht || !gates)) return BGNN_ERR_INVALID_ARG;
  return launch_adapted_skinny_fwd(x, is_src, wcat, bias, wd, kg, n, d, c, heads, Hs, Ht, gates, (cudaStream_t)stream);
}

size_t bgnn_adapted_skinny_heads_bwd_workspace_bytes(int c, int d, int heads) {
  if (c <= 0 || d <= 0 || heads <= 0) return 0;
  const size_t a = adapted_skinny_bwd_workspace_bytes(c, d, heads), b = adapted_skinny_pre_workspace_bytes(c, heads);
  return a > b ? a : b;
}

int bgnn_adapted_skinny_heads_pre_f32(const uint8_t* is_src, const float* wd, const float* gates, const float* gHs,
                                      const float* gHt, int64_t n, int c, int heads, float* pre, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  if (n < 0 || c <= 0 || heads <= 0 || !pre || !workspace) return BGNN_ERR_INVALID_ARG;
  if (n > 0 && (!is_src || !wd || !gates || !gHs || !gHt)) return BGNN_ERR_INVALID_ARG;
  return launch_adapted_skinny_pre(is_src, wd, gates, gHs, gHt, n, c, heads, pre, workspace, workspace_bytes,
                                   (cudaStream_t)stream);
}

int bgnn_adapted_skinny_heads_bwd_f32(const float* x, const uint8_t* is_src, const float* wcat, const float* wd,
                                      const float* gates, const float* gHs, const float* gHt, const float* gm, int64_t n,
                                      int d, int c, int heads, float* gx, float* red, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  if (n < 0 || c <= 0 || d <= 0 || heads <= 0) return BGNN_ERR_INVALID_ARG;
  if (n > 0 && (!x || !is_src || !wcat || !wd || !gates || !gHs || !gHt || !gx || !red || !workspace))
    return BGNN_ERR_INVALID_ARG;
  return launch_adapted_skinny_bwd(x, is_src, wcat, wd, gates, gHs, gHt, gm, n, d, c, heads, gx, red, workspace,
                                   workspace_bytes, (cudaStream_t)stream);
}

int bgnn_adapted_transform_bwd_gates_f32(const float* gHs, const float* gHt, const float* gates, const uint8_t* is_src,
                                         const float* wd, int64_t n, int c, float* dg, float* g_wd_kg, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  if (n < 0 || c <= 0) return BGNN_ERR_INVALID_ARG;
  if (n > 0 && (!gHs || !gHt || !gates || !is_src || !wd || !dg || !g_wd_kg || !workspace)) return BGNN_ERR_INVALID_ARG;
  return launch_adapted_transform_bwd(gHs, gHt, gates, is_src, wd, n, c, 4, 0, dg, g_wd_kg, workspace, workspace_bytes,
                                      (cudaStream_t)stream);
}

int bgnn_wgrad_gemm_cat_f32(const float* G0, int ld0, int no0, const float* G1, int ld1, int no1, const float* G2, int ld2,
                            int no2, const float* X, int ld_x, int d, int64_t n, float* W, int ldw, float* colsum,
                            void* workspace, size_t workspace_bytes, void* stream) {
  const float* G[3] = {G0, G1, G2};
  const int ld[3] = {ld0, ld1, ld2}, no[3] = {no0, no1, no2};
  const int nblk = G2 ? 3 : (G1 ? 2 : 1);
  if (n < 0 || d <= 0 || ld_x < d || ldw < d || !G0 || !W || !workspace || (G2 && !G1)) return BGNN_ERR_INVALID_ARG;
  if (n > 0 && !X) return BGNN_ERR_INVALID_ARG;
  uintptr_t bits = reinterpret_cast<uintptr_t>(X);
  for (int i = 0; i < nblk; ++i) bits |= reinterpret_cast<uintptr_t>(G[i]);
  if (bits & 15) return BGNN_ERR_INVALID_ARG;
  return launch_wgrad_gemm_cat(G, ld, no, nblk, X, ld_x, d, n, W, ldw, colsum, workspace, workspace_bytes, (cudaStream_t)stream);
}

int bgnn_adapted_skinny_tc_supported(int c, int d, int heads) { return adapted_skinny_tc_supported(c, d, heads) ? 1 : 0; }

int bgnn_adapted_skinny_heads_tc_fwd_f32(const float* x, int64_t n, int d, const float* wcat_hi, const float* wcat_lo, int c,
                                         int heads, const uint8_t* is_src, const float* wd, const float* kg,
                                         const float* bias, float* Hs, float* Ht, float* gates, void* stream) {
  if (n < 0 || c <= 0 || d <= 0 || heads <= 0) return BGNN_ERR_INVALID_ARG;
  if (n > 0 && (!x || !wcat_hi || !wcat_lo || !is_src || !wd || !kg || !Hs || !Ht || !gates)) return BGNN_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wcat_hi) | reinterpret_cast<uintptr_t>(wcat_lo)) & 15)
    return BGNN_ERR_INVALID_ARG;
  return launch_adapted_skinny_tc_fwd(x, n, d, wcat_hi, wcat_lo, c, heads, is_src, wd, kg, bias, Hs, Ht, gates,
                                      (cudaStream_t)stream);
}

int bgnn_tf32_planes_f32(const float* w, int rows, int cols, int64_t stride_r, int64_t stride_c, int rows_p, int cols_p,
                         float* hi, float* lo, void* stream) {
  if (!hi || !lo || (rows > 0 && cols > 0 && !w)) return BGNN_ERR_INVALID_ARG;
  return launch_tf32_planes(w, rows, cols, stride_r, stride_c, rows_p, cols_p, hi, lo, (cudaStream_t)stream);
}

}  // extern "C"
