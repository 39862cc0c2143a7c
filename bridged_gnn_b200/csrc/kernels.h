// Internal launcher declarations (C++).  The public surface is include/bgnn_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bgnn {

// knn_simt.cu
#define BGNN_PAIR_DOT 0
#define BGNN_PAIR_ADDRELU 1
int launch_knn_simt(int mode, const float* Q, const float* Qlo, int nq, const float* DB, const float* DBlo, int ndb,
                    int d, int ld, const float* w, float bias, int apply_sigmoid, int kc, int nsplit, int db_per_split,
                    const int* row_list, const int* row_count, int few_rows, float* cand_val, int* cand_idx,
                    cudaStream_t stream);
// exact sweep for a handful of listed rows (0 < *row_count <= knn_exact_rows_max(), decided on the device)
int knn_exact_rows_max(int ndb, int ld, int k);
size_t knn_exact_rows_workspace_bytes(int k);
int launch_knn_exact_rows(const float* Q, const float* Qlo, const float* DB, const float* DBlo, int ndb, int d, int ld,
                          int apply_sigmoid, int k, const int* row_list, const int* row_count, long long* out_idx,
                          float* out_val, float* out_gap, float eps, int* out_count, void* ws, size_t ws_bytes,
                          cudaStream_t stream);

// knn_select.cu
#define BGNN_MERGE_MAX_CAND 1024
int launch_normalize_split(const float* x, long long n, int d, int ld, int normalize, float* hi, float* lo,
                           cudaStream_t stream);
int launch_knn_merge(const float* cand_val, const int* cand_idx, int nlists, int kc, int nq, int k, int rescore,
                     const float* qhi, const float* qlo, const float* dhi, const float* dlo, int d, int ld,
                     int apply_sigmoid, float delta, const float* seed_thr, const int* row_list, const int* row_count,
                     int few_rows, long long* out_idx, float* out_val, float* out_gap, int* fb_rows, int* fb_count,
                     float eps, int* out_count, cudaStream_t stream);

// knn_cosine_sm100.cu  (tcgen05 / TMEM / TMA)
struct TcPlan {
  int bn;            // db rows per MMA tile (256, or 128 when k needs longer lists)
  int nsplit;        // db splits (grid.y)
  int tiles_per_split;
  int kc;            // candidates kept per (row, list)
  int stages;        // slots of the TMA->MMA operand ring
  int kps;           // f16 sweep: k-blocks (64 features) per ring slot; 1 elsewhere
  int nlists;        // lists per row (nsplit; f16 sweep: ew * nsplit)
  int pair;          // f16 sweep: CTA pairs (cluster of 2, cta_group::2) sharing every db tile
  int ew;            // f16 sweep: epilogue warps per TMEM lane quarter = lists per (row, split): 2 or 4
};
TcPlan tc_plan(int nq, int ndb, int d, int k, int passes);
int launch_knn_cosine_tc(const float* qhi, const float* qlo, int nq, const float* dhi, const float* dlo, int ndb, int d,
                         int passes, const TcPlan& plan, float* cand_val, int* cand_idx, cudaStream_t stream);

// knn_cosine_f16_sm100.cu  (tcgen05 kind::f16, single pass)
constexpr int kSeedKc = 4;    // list length of the threshold-seeding sweep over a db sample
TcPlan tc_plan_f16(int nq, int ndb, int d, int k, int kc_fixed = 0);
int knn_seed_rows(int ndb, int k);   // sampled db rows for threshold seeding (0 = no seeding)
int launch_knn_seed_f16(const void* qh, int nq, const void* dh, int ndb, int ldh, int srows, const TcPlan& seed_plan,
                        void* sample, float* cand_val, int* cand_idx, float* seed, cudaStream_t stream);
int launch_normalize_f16(const float* x, long long n, int d, int ld, int ldh, int normalize, float* xn, void* xh,
                         cudaStream_t stream);
int launch_knn_cosine_f16(const void* qh, int nq, const void* dh, int ndb, int ldh, const TcPlan& plan,
                          const float* thr_init, float* cand_val, int* cand_idx, cudaStream_t stream);

// edge_filter.cu
size_t quantile_workspace_bytes();
int launch_quantile(const float* v, long long n, long long rank_lo, float weight, float* out3, void* ws, size_t ws_bytes,
                    cudaStream_t stream);
int launch_edge_validity(const long long* e0, const long long* e1, long long e, const float* e_sim, const float* thr_conf,
                         const long long* pred_a, const long long* y_a, const long long* pred_b, const long long* y_b,
                         const uint8_t* gate_a, const uint8_t* gate_b, const float* x_a, const float* x_b, int d,
                         float thres_feat_sim, uint8_t* keep, long long* counts, cudaStream_t stream);

// csr_build.cu
size_t csr_build_workspace_bytes(long long e);
int launch_edges_to_csr(const long long* src, const long long* dst, long long e, long long n, int dedup, int* rowptr,
                        int* col, long long* perm, long long* e_out, void* ws, size_t ws_bytes, cudaStream_t stream);

size_t graph_prepare_workspace_bytes(long long e, long long n, int rewrite);
int launch_graph_prepare(const long long* src, const long long* dst, long long e, long long n, int rewrite, int* rowptr,
                         int* col, int* t_rowptr, int* t_col, int* csr_to_csc, int* order, int* t_order, long long* e_out,
                         void* ws, size_t ws_bytes, cudaStream_t stream);

size_t rows_by_degree_workspace_bytes(long long n);
int launch_rows_by_degree(const int* rowptr, long long n, int min_degree, int* order, void* ws, size_t ws_bytes,
                          cudaStream_t stream);

// spmm_csr.cu
int launch_spmm_csr(const int* rowptr, const int* col, const float* edge_w, const float* gather_scale,
                    const float* out_scale, const float* X, long long ldx, long long n_rows, int f, int reduce_mean, float* Y,
                    long long ldy, cudaStream_t stream);

// gatv2_fused.cu
int launch_gatv2_fwd(const int* rowptr, const int* col, const int* order, const uint8_t* dst_is_src, const float* Hs, const float* Ht,
                     const float* af_t2s, const float* af_s2t, float slope, long long n, long long row_off, int c, float* out,
                     float* row_max, float* row_sum, float* score, int score_only, cudaStream_t stream);
size_t gatv2_bwd_workspace_bytes(long long n, long long e, int c);
int launch_gatv2_bwd(const int* rowptr, const int* col, const int* t_rowptr, const int* t_col, const int* csr_to_csc,
                     const int* order, const int* t_order, long long e, const uint8_t* dst_is_src, const float* Hs, const float* Ht, const float* af_t2s,
                     const float* af_s2t, float slope, long long n, long long row_off, long long n_src, int c, const float* out,
                     const float* row_max, const float* row_sum, const float* score, const float* gout, float* gHs, float* gHt,
                     float* g_af_t2s, float* g_af_s2t, void* ws, size_t ws_bytes, cudaStream_t stream);

// adapted_transform.cu
int launch_adapted_transform_fwd(const float* P, const uint8_t* is_src, const float* wd, const float* kg, const float* bias,
                                 long long n, int c, float* Hs, float* Ht, float* gates, cudaStream_t stream);
size_t adapted_transform_bwd_workspace_bytes(int c);
int launch_adapted_transform_bwd(const float* gHs, const float* gHt, const float* gates, const uint8_t* is_src,
                                 const float* wd, long long n, int c, int ldp, int copy, float* gP, float* g_wd_kg,
                                 void* ws, size_t ws_bytes, cudaStream_t stream);

// gatv2_heads.cu: 2-3 narrow aggregations over the same graph in one pass
bool gatv2_heads_supported(int heads, int c);
int launch_gatv2_heads_fwd(const int* rowptr, const int* col, const uint8_t* dst_is_src, const float* Hs, const float* Ht,
                           const float* af_t2s, const float* af_s2t, float slope, long long n, long long row_off, int heads, int c,
                           float* out, float* row_max, float* row_sum, cudaStream_t stream);
size_t gatv2_heads_bwd_workspace_bytes(long long n, long long e, int heads, int c);
int launch_gatv2_heads_bwd(const int* rowptr, const int* col, const int* t_rowptr, const int* t_col, const int* csr_to_csc,
                           long long e, const uint8_t* dst_is_src, const float* Hs, const float* Ht, const float* af_t2s,
                           const float* af_s2t, float slope, long long n, long long row_off, long long n_src, int heads, int c,
                           const float* out, const float* row_max, const float* row_sum, const float* gout, float* gHs, float* gHt,
                           float* g_af_t2s, float* g_af_s2t, void* ws, size_t ws_bytes, cudaStream_t stream);

// adapted_skinny.cu
bool adapted_skinny_supported(int c, int d);
bool adapted_skinny_heads_supported(int c, int d, int heads);
int launch_adapted_skinny_fwd(const float* x, const uint8_t* is_src, const float* wcat, const float* bias, const float* wd,
                              const float* kg, long long n, int d, int c, int heads, float* Hs, float* Ht, float* gates,
                              cudaStream_t stream);
size_t adapted_skinny_bwd_workspace_bytes(int c, int d, int heads);
int launch_adapted_skinny_bwd(const float* x, const uint8_t* is_src, const float* wcat, const float* wd, const float* gates,
                              const float* gHs, const float* gHt, const float* gm, long long n, int d, int c, int heads,
                              float* gx, float* red, void* ws, size_t ws_bytes, cudaStream_t stream);
size_t adapted_skinny_pre_workspace_bytes(int c, int heads);
int launch_adapted_skinny_pre(const uint8_t* is_src, const float* wd, const float* gates, const float* gHs, const float* gHt,
                              long long n, int c, int heads, float* pre, void* ws, size_t ws_bytes, cudaStream_t stream);
bool domain_colsum_supported(int d);
size_t domain_colsum_workspace_bytes(int d);
int launch_domain_colsum(const float* x, const uint8_t* is_src, long long n, int d, float* sums, void* ws, size_t ws_bytes,
                         cudaStream_t stream);

// rowpanel_gemm_sm100.cu
bool rowpanel_gemm_supported(int k, int ld_a, int no);
int launch_rowpanel_gemm(const float* A, long long n, int k, int ld_a, const float* bhi, const float* blo, const float* scale,
                         const float* bias, int act, const float* res, int ld_res, int no, float* Y, int ldy,
                         cudaStream_t stream);
int launch_tf32_planes(const float* w, int rows, int cols, long long stride_r, long long stride_c, int rows_p, int cols_p, float* hi,
                       float* lo, cudaStream_t stream);
bool adapted_wide_supported(int c, int d);
bool adapted_skinny_tc_supported(int c, int d, int heads);
int launch_adapted_skinny_tc_fwd(const float* x, long long n, int d, const float* wcat_hi, const float* wcat_lo, int c, int heads,
                                 const uint8_t* is_src, const float* wd, const float* kg, const float* bias, float* Hs,
                                 float* Ht, float* gates, cudaStream_t stream);
int launch_adapted_wide_fwd(const float* x, long long n, int d, const float* wcat_hi, const float* wcat_lo, int c,
                            const uint8_t* is_src, const float* wd, const float* kg, const float* bias, float* Hs, float* Ht,
                            float* gates, cudaStream_t stream);

// wgrad_gemm_sm100.cu
bool wgrad_gemm_supported(int d, int ld_x, int no, int ld_g);
size_t wgrad_gemm_workspace_bytes(int no);
int launch_wgrad_gemm(const float* G, int ld_g, int no, const float* X, int ld_x, int d, long long n, float* W, int ldw,
                      float* colsum, void* ws, size_t ws_bytes, cudaStream_t stream);
int launch_wgrad_gemm_cat(const float* const* G, const int* ld_g, const int* no, int nblk, const float* X, int ld_x, int d,
                          long long n, float* W, int ldw, float* colsum, void* ws, size_t ws_bytes, cudaStream_t stream);

// bn_relu.cu
bool bn_relu_supported(int c);
size_t bn_relu_workspace_bytes(int c);
int launch_bn_relu_fwd(const float* x, long long n, int c, const float* w, const float* b, float eps, float momentum,
                       float* running_mean, float* running_var, int relu, float* y, float* stats, void* ws, size_t ws_bytes,
                       cudaStream_t stream);
int launch_bn_relu_apply(const float* x, long long n, int c, const float* stats, int relu, float* y, cudaStream_t stream);
int launch_bn_relu_bwd_reduce(const float* gy, const float* x, long long n, int c, const float* stats, int relu, float* gwb,
                              void* ws, size_t ws_bytes, cudaStream_t stream);
int launch_bn_relu_bwd_apply(const float* gy, const float* x, long long n, int c, const float* stats, int relu,
                             const float* coef, float* gx, cudaStream_t stream);
int launch_bn_relu_bwd(const float* gy, const float* x, long long n, int c, const float* stats, int relu, float* gx, float* gwb,
                       void* ws, size_t ws_bytes, cudaStream_t stream);

}  // namespace bgnn
