/* libbgnn_b200 -- C ABI of the B200-native Bridged-GNN hot path.
 *
 * The reference (wendongbi/Bridged-GNN) is pure Python on PyTorch/PyG and has no FFI; each entry
 * point below names the reference call site(s) it replaces (paths relative to the reference
 * checkout).  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller unless stated otherwise; the library
 *    allocates nothing persistent.  Scratch comes from a caller-provided workspace whose size is
 *    returned by the matching *_workspace_bytes() query (pure host arithmetic, no GPU needed).
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream and does
 *    no host synchronisation.
 *  - return value: 0 = OK; negative = BGNN_ERR_* below; positive = cudaError_t of a failed launch.
 *  - node indices are int64 at this boundary (PyG convention), int32 inside CSR (N, E < 2^31).
 *  - sm_100a only: there is no CPU or other-architecture fallback.
 */
#ifndef BGNN_B200_H_
#define BGNN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BGNN_ERR_INVALID_ARG (-1)
#define BGNN_ERR_WORKSPACE (-2)
#define BGNN_ERR_UNSUPPORTED (-3)
#define BGNN_ERR_DRIVER (-4)

/* kNN algorithms for the cosine head */
#define BGNN_KNN_SIMT_F32 0   /* exact fp32 on CUDA cores                                    */
#define BGNN_KNN_TC_3XTF32 1  /* tcgen05 3xTF32 split + exact fp32 re-score + certification */
#define BGNN_KNN_TC_1XTF32 2  /* tcgen05 single TF32 pass + exact re-score + certification   */
#define BGNN_KNN_TC_F16 3     /* tcgen05 single FP16 pass (resident query block) + exact re-score + certification */

int bgnn_version(void);
const char* bgnn_error_string(int code);

/* ---- bridged-graph construction: fused all-pairs similarity + per-row top-k ------------------
 *
 * Replaces main_bridged_graph.py:45-67 and :90-111 (pair_enumeration models/models.py:265-282,
 * similarity models/models.py:124-130 / 945-948 [cosine] and :949-954 [mlp], sim_mat.topk at
 * main_bridged_graph.py:60,104).  The [nq, ndb] similarity matrix is never written to memory.
 *
 * Selection key (parity definition): post-sigmoid fp32 similarity descending, db index ascending.
 * Outputs: out_idx[nq,k] int64 (db row of each neighbour, best first), out_val[nq,k] fp32 similarity,
 * out_gap[nq] fp32 = sim_k - sim_(k+1) (+inf if ndb == k; rows with gap < 1e-6 are near-ties),
 * out_stats[4] int32: [0] rows re-done by the exact fallback (tensor-core algorithms), [1..3] reserved.
 * out_gap / out_stats may be NULL.  k > ndb is BGNN_ERR_INVALID_ARG (torch.topk raises there too).
 */

/* cosine head: sim = sigmoid( <q_i/max(|q_i|,1e-8), db_j/max(|db_j|,1e-8)> ).
 * q [nq,d], db [ndb,d] row-major fp32 are the vectors fed to CosineSimilarity (u = z' + biasatt(z')).
 * normalize=0 if rows are already unit length.  q == db (same pointer, nq == ndb) is the
 * within-domain case; self matches are kept, as in the reference. */
size_t bgnn_knn_cosine_workspace_bytes(int64_t nq, int64_t ndb, int d, int k, int algo);
int bgnn_knn_cosine_f32(const float* q, int64_t nq, const float* db, int64_t ndb, int d, int k, int normalize,
                        int apply_sigmoid, int algo, int64_t* out_idx, float* out_val, float* out_gap,
                        int32_t* out_stats, void* workspace, size_t workspace_bytes, void* stream);

/* The same with the bridge-matching threshold `epsilon` (main_bridged_graph.py:33; the reference accepts the argument
 * and never reads it) fused into the selection epilogue: a neighbour is emitted only if its similarity is > eps
 * (eps = NaN: off, identical to bgnn_knn_cosine_f32).  Dropped places keep their similarity in out_val and get
 * out_idx = -1; rows are best first, so the kept neighbours are a prefix of length out_count[row] (int32 [nq],
 * may be NULL).  out_gap is unaffected. */
int bgnn_knn_cosine_eps_f32(const float* q, int64_t nq, const float* db, int64_t ndb, int d, int k, int normalize,
                            int apply_sigmoid, int algo, float eps, int64_t* out_idx, float* out_val, float* out_gap,
                            int32_t* out_count, int32_t* out_stats, void* workspace, size_t workspace_bytes, void* stream);

/* v2 'mlp' head with eval-mode BatchNorm folded (models/models.py:918-925):
 * sim = sigmoid( sum_h w2[h] * relu(Uq[i,h] + Udb[j,h]) + b2 ),  Uq [nq,h], Udb [ndb,h], w2 [h]. */
size_t bgnn_knn_addrelu_workspace_bytes(int64_t nq, int64_t ndb, int h, int k);
int bgnn_knn_addrelu_f32(const float* Uq, int64_t nq, const float* Udb, int64_t ndb, int h, const float* w2, float b2,
                         int k, int apply_sigmoid, int64_t* out_idx, float* out_val, float* out_gap,
                         void* workspace, size_t workspace_bytes, void* stream);

int bgnn_knn_addrelu_eps_f32(const float* Uq, int64_t nq, const float* Udb, int64_t ndb, int h, const float* w2, float b2,
                             int k, int apply_sigmoid, float eps, int64_t* out_idx, float* out_val, float* out_gap,
                             int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream);

/* ---- edge-validity filters of the build -----------------------------------------------------
 *
 * bgnn_quantile_f32: order statistics of v[n] by radix select (no sort, no size cap).  rank_lo = floor(q (n-1)),
 * weight = q (n-1) - rank_lo; out3 = (v_(rank_lo), v_(rank_lo+1), their linear interpolation as torch.quantile
 * computes it).  Replaces e_sim.quantile(q) at main_bridged_graph.py:134, 236.
 *
 * bgnn_edge_validity_f32: the four removal rules of check_added_edges_cross_domain_validity /
 * check_added_edges_within_domain_validity (main_bridged_graph.py:225-264, 123-161) in one pass over the edges:
 *   1. e_sim[i] < *thr_conf (device scalar; NULL = rule off)
 *   2. pred_a[e0] != y_a[e0] (only where gate_a[e1] != 0 when gate_a is given), or pred_b[e1] != y_b[e1] where
 *      gate_b[e1] != 0 (gate_b NULL = everywhere)
 *   3. pred_a[e0] != pred_b[e1]
 *   4. cosine_similarity(x_a[e0], x_b[e1]) < thres_feat_sim   (x_a, x_b [*, d] raw features, ATen's formula)
 * keep [e] bytes: 1 = edge survives.  counts [5] int64: edges newly removed by rule 1, 2, 3, 4 in that order
 * (the numbers the reference prints) and the number kept.  cross-domain: a = source side, b = target side,
 * gate_a = NULL, gate_b = target train_mask; within-domain: a = b, gate_a = gate_b = train_mask. */
size_t bgnn_quantile_workspace_bytes(void);
int bgnn_quantile_f32(const float* v, int64_t n, int64_t rank_lo, float weight, float* out3, void* workspace,
                      size_t workspace_bytes, void* stream);
int bgnn_edge_validity_f32(const int64_t* e0, const int64_t* e1, int64_t e, const float* e_sim, const float* thr_conf,
                           const int64_t* pred_a, const int64_t* y_a, const int64_t* pred_b, const int64_t* y_b,
                           const uint8_t* gate_a, const uint8_t* gate_b, const float* x_a, const float* x_b, int d,
                           float thres_feat_sim, uint8_t* keep, int64_t* counts, void* stream);

/* ---- graph format: edge list -> destination-major CSR ---------------------------------------
 *
 * Replaces torch_sparse.SparseTensor(row=edge_index[1], col=edge_index[0]) (models/backbones.py:464)
 * and, with dedup=1, torch_geometric.utils.coalesce (main_bridged_graph.py:75,113,193) up to the
 * (dst,src) instead of (src,dst) sort order.  rowptr [n+1] int32, col [e] int32 (source of each
 * edge, rows sorted by destination then source), perm [e] int64 (position in the input edge list;
 * may be NULL), e_out [2] int64: [0] = number of edges kept, [1] = number of edges with a node id outside [0, n)
 * (PyG raises an index error on those; here they are parked behind the last row where no kernel reads them, and the
 * host layer turns a non-zero count into an error). */
size_t bgnn_edges_to_csr_workspace_bytes(int64_t e);
int bgnn_edges_to_csr(const int64_t* src, const int64_t* dst, int64_t e, int64_t n, int dedup, int32_t* rowptr,
                      int32_t* col, int64_t* perm, int64_t* e_out, void* workspace, size_t workspace_bytes,
                      void* stream);

/* One-call graph preparation of the aggregation kernels, straight from the caller's edge_index.
 * Replaces KTGNN.graph_partition's self-loop rewrite (models/KTGNN.py:385-398: remove_self_loops + add_self_loops;
 * the split by destination domain needs no edge copy here, the kernels pick the branch per destination row) and the
 * per-forward SparseTensor build (models/backbones.py:464).  rewrite_self_loops != 0: input edges with src == dst are
 * dropped and one (v, v) per node is added.  Capacity of col / t_col / csr_to_csc: e + (rewrite_self_loops ? n : 0).
 *   rowptr [n+1], col            destination-major CSR (rows sorted by destination, then source)
 *   t_rowptr [n+1], t_col        CSR of the transposed graph (rows = sources, entries = destinations, ascending)
 *   csr_to_csc                   slot of every CSR edge in the transposed CSR       (the three: all NULL or none)
 *   order, t_order [n]           rows by descending in- / out-degree (bgnn_rows_by_degree(.., 0)); either may be NULL
 *   e_out [2] int64              [0] = edges kept, [1] = input edges with a node id outside [0, n) (parked, unread)
 * Identical to bgnn_edges_to_csr on the rewritten edge list and on its transpose. */
size_t bgnn_graph_prepare_workspace_bytes(int64_t e, int64_t n, int rewrite_self_loops);
int bgnn_graph_prepare(const int64_t* src, const int64_t* dst, int64_t e, int64_t n, int rewrite_self_loops,
                       int32_t* rowptr, int32_t* col, int32_t* t_rowptr, int32_t* t_col, int32_t* csr_to_csc,
                       int32_t* order, int32_t* t_order, int64_t* e_out, void* workspace, size_t workspace_bytes,
                       void* stream);

/* ---- message passing ------------------------------------------------------------------------
 *
 * CSR SpMM:  Y[i,:] = out_scale[i] * (1/deg_i if reduce_mean) * sum_e edge_w[e] * gather_scale[col[e]] * X[col[e],:]
 * Replaces torch_sparse.matmul(adj_t, x, reduce=...) under SAGEConv (models/backbones.py:464-468,
 * models/models.py:250-253) and GCNConv's normalised propagate (models/backbones.py:272-274).
 * edge_w [e], gather_scale [n_cols], out_scale [n_rows] may each be NULL (= 1).  The backward pass
 * is the same call on the transposed CSR. */
int bgnn_spmm_csr_f32(const int32_t* rowptr, const int32_t* col, const float* edge_w, const float* gather_scale,
                      const float* out_scale, const float* X, int64_t n_rows, int f, int reduce_mean, float* Y,
                      void* stream);

/* The same with row strides (elements) of X and Y, so that a column panel of a wider matrix is addressed in place:
 * the destination-partitioned multi-GPU path gathers X panel by panel and runs this on panel p while panel p+1 is
 * still on the wire.  128-bit loads need ldx, ldy multiples of 4 and 16-byte aligned bases. */
int bgnn_spmm_csr_ld_f32(const int32_t* rowptr, const int32_t* col, const float* edge_w, const float* gather_scale,
                         const float* out_scale, const float* X, int64_t ldx, int64_t n_rows, int f, int reduce_mean,
                         float* Y, int64_t ldy, void* stream);

/* Fused AdaptedConv aggregation (models/KTGNN.py:292-305, message :317-319; PyG softmax +
 * propagate underneath).  Per destination row i: (H, a) = (Hs, af_t2s) if dst_is_src[i] else
 * (Ht, af_s2t); score_j = a . leaky_relu(H[j] + H[i], slope); out[i] = sum_j softmax_j(score) H[j].
 * Hs = lin_s(x_t2s), Ht = lin_t(x_s2t), both [n,c]; dst_is_src [n] bytes (central_mask);
 * row_max / row_sum [n] are saved for the backward pass (may be NULL for inference). */
int bgnn_gatv2_fwd_f32(const int32_t* rowptr, const int32_t* col, const uint8_t* dst_is_src, const float* Hs,
                       const float* Ht, const float* af_t2s, const float* af_s2t, float slope, int64_t n, int c,
                       float* out, float* row_max, float* row_sum, void* stream);

/* Same with an optional processing order of the rows (row_order [n] int32, a permutation of 0..n-1, or NULL):
 * bgnn_rows_by_degree gives "longest rows first", which removes the tail a hub row of a kNN graph otherwise
 * forms and evens out the rows that share a warp.  Rows shorter than min_degree keep their natural order after
 * the long ones (use that for narrow feature rows, whose row-level accesses should stay coalesced; 0 = full
 * sort).  Forward results do not depend on the order; in the backward the longest rows of an order get a whole warp
 * each (their edges are split over several lane groups), which changes the fp32 summation order of those rows'
 * gradients -- every order gives run-to-run reproducible results. */
int bgnn_gatv2_fwd_ord_f32(const int32_t* rowptr, const int32_t* col, const int32_t* row_order, const uint8_t* dst_is_src,
                           const float* Hs, const float* Ht, const float* af_t2s, const float* af_s2t, float slope,
                           int64_t n, int c, float* out, float* row_max, float* row_sum, void* stream);
/* Training / destination-partitioned form of the forward.  n_rows destination rows with LOCAL ids 0..n_rows-1 (rowptr,
 * row_order, out, row_max, row_sum) whose global node id is row + row_off; Hs, Ht and dst_is_src are indexed by
 * GLOBAL node id (col entries are global ids).  score [e] (CSR edge order) or NULL: the per-edge attention scores
 * a . leaky_relu(H[src] + H[dst]), kept for bgnn_gatv2_bwd_part_f32 so that the backward never recomputes them.
 * Single GPU: row_off = 0. */
int bgnn_gatv2_fwd_part_f32(const int32_t* rowptr, const int32_t* col, const int32_t* row_order, const uint8_t* dst_is_src,
                            const float* Hs, const float* Ht, const float* af_t2s, const float* af_s2t, float slope,
                            int64_t n_rows, int64_t row_off, int c, float* out, float* row_max, float* row_sum, float* score,
                            void* stream);
size_t bgnn_rows_by_degree_workspace_bytes(int64_t n);
int bgnn_rows_by_degree(const int32_t* rowptr, int64_t n, int min_degree, int32_t* order, void* workspace,
                        size_t workspace_bytes, void* stream);

/* Backward of the above.  (rowptr,col) = CSR by destination with e edges, (t_rowptr,t_col) = CSR of the
 * transposed graph (rows = sources, entries = destinations), csr_to_csc [e] = slot of every CSR edge in the
 * transposed CSR.  Writes gHs, gHt [n,c] (fully), g_af_t2s, g_af_s2t [c].  Deterministic, atomic-free. */
size_t bgnn_gatv2_bwd_workspace_bytes(int64_t n, int64_t e, int c);
int bgnn_gatv2_bwd_f32(const int32_t* rowptr, const int32_t* col, const int32_t* t_rowptr, const int32_t* t_col,
                       const int32_t* csr_to_csc, int64_t e, const uint8_t* dst_is_src, const float* Hs,
                       const float* Ht, const float* af_t2s, const float* af_s2t, float slope, int64_t n, int c,
                       const float* out, const float* row_max, const float* row_sum, const float* gout, float* gHs,
                       float* gHt, float* g_af_t2s, float* g_af_s2t, void* workspace, size_t workspace_bytes,
                       void* stream);

/* Same with optional processing orders for the destination-major (row_order) and the transposed (t_row_order) CSR.
 * g_af_* are summed in a fixed order for a given row_order, so results are reproducible run to run. */
int bgnn_gatv2_bwd_ord_f32(const int32_t* rowptr, const int32_t* col, const int32_t* t_rowptr, const int32_t* t_col,
                           const int32_t* csr_to_csc, const int32_t* row_order, const int32_t* t_row_order, int64_t e,
                           const uint8_t* dst_is_src, const float* Hs, const float* Ht, const float* af_t2s,
                           const float* af_s2t, float slope, int64_t n, int c, const float* out, const float* row_max,
                           const float* row_sum, const float* gout, float* gHs, float* gHt, float* g_af_t2s,
                           float* g_af_s2t, void* workspace, size_t workspace_bytes, void* stream);

/* Backward with the forward's scores and the destination-partitioned layout: (rowptr, col) has n_rows local destination
 * rows, (t_rowptr, t_col) n_src source rows (all nodes) whose entries are LOCAL destination ids; gout, out, row_max,
 * row_sum are local, Hs, Ht, dst_is_src, gHs, gHt global ([n_src, c]; the destination-side part of local row r lands at
 * r + row_off).  score [e] from bgnn_gatv2_fwd_part_f32, or NULL (recomputed by one extra sweep).  gHs or gHt may be
 * NULL on a rank that owns no destination row of that domain (nothing would be written but zeros).  The workspace is
 * sized by bgnn_gatv2_bwd_workspace_bytes(n_src, e, c).  Single GPU: row_off = 0, n_src = n_rows. */
int bgnn_gatv2_bwd_part_f32(const int32_t* rowptr, const int32_t* col, const int32_t* t_rowptr, const int32_t* t_col,
                            const int32_t* csr_to_csc, const int32_t* row_order, const int32_t* t_row_order, int64_t e,
                            const uint8_t* dst_is_src, const float* Hs, const float* Ht, const float* af_t2s,
                            const float* af_s2t, float slope, int64_t n_rows, int64_t row_off, int64_t n_src, int c,
                            const float* out, const float* row_max, const float* row_sum, const float* score,
                            const float* gout, float* gHs, float* gHt, float* g_af_t2s, float* g_af_s2t, void* workspace,
                            size_t workspace_bytes, void* stream);

/* Two or three NARROW aggregations (heads) over the same graph in one pass -- KT-GNN's classifier convs
 * clf_base(x), clf_target(clf_transformer(x)), clf_target(x) (models/KTGNN.py:432-434).  Hs, Ht, out, gout, gHs,
 * gHt are [n, heads*c] with head h in columns h*c .. h*c+c-1; af_*, g_af_* are [heads*c]; row_max / row_sum are
 * [n, heads].  Per head the semantics are exactly bgnn_gatv2_fwd_f32 / _bwd_f32 (scores, softmax and sums never mix
 * heads); an edge costs one index load, one gather and one 32-byte record for all heads.  heads in {2, 3}, c <= 4
 * (bgnn_gatv2_heads_supported). */
int bgnn_gatv2_heads_supported(int heads, int c);
int bgnn_gatv2_heads_fwd_f32(const int32_t* rowptr, const int32_t* col, const uint8_t* dst_is_src, const float* Hs,
                             const float* Ht, const float* af_t2s, const float* af_s2t, float slope, int64_t n, int heads,
                             int c, float* out, float* row_max, float* row_sum, void* stream);
size_t bgnn_gatv2_heads_bwd_workspace_bytes(int64_t n, int64_t e, int heads, int c);
int bgnn_gatv2_heads_bwd_f32(const int32_t* rowptr, const int32_t* col, const int32_t* t_rowptr, const int32_t* t_col,
                             const int32_t* csr_to_csc, int64_t e, const uint8_t* dst_is_src, const float* Hs,
                             const float* Ht, const float* af_t2s, const float* af_s2t, float slope, int64_t n, int heads,
                             int c, const float* out, const float* row_max, const float* row_sum, const float* gout,
                             float* gHs, float* gHt, float* g_af_t2s, float* g_af_s2t, void* workspace,
                             size_t workspace_bytes, void* stream);

/* Destination-partitioned forms (see bgnn_gatv2_fwd_part_f32 / bgnn_gatv2_bwd_part_f32 for the index conventions). */
int bgnn_gatv2_heads_fwd_part_f32(const int32_t* rowptr, const int32_t* col, const uint8_t* dst_is_src, const float* Hs,
                                  const float* Ht, const float* af_t2s, const float* af_s2t, float slope, int64_t n_rows,
                                  int64_t row_off, int heads, int c, float* out, float* row_max, float* row_sum, void* stream);
int bgnn_gatv2_heads_bwd_part_f32(const int32_t* rowptr, const int32_t* col, const int32_t* t_rowptr, const int32_t* t_col,
                                  const int32_t* csr_to_csc, int64_t e, const uint8_t* dst_is_src, const float* Hs,
                                  const float* Ht, const float* af_t2s, const float* af_s2t, float slope, int64_t n_rows,
                                  int64_t row_off, int64_t n_src, int heads, int c, const float* out, const float* row_max,
                                  const float* row_sum, const float* gout, float* gHs, float* gHt, float* g_af_t2s,
                                  float* g_af_s2t, void* workspace, size_t workspace_bytes, void* stream);

/* Node-wise epilogue of AdaptedConv's domain-shift transform (models/KTGNN.py:275-284).  The host computes
 * P [n, 2c+2] = x [W_s; W_t; a_g_s2t[:D]; a_g_t2s[:D]]^T, wd [2c] = [W_s Delta; W_t Delta] and
 * kg [2] = [a_g_s2t[D:].Delta, a_g_t2s[D:].Delta]; bias [2c] = (b_s, b_t) or NULL.  This call writes
 *   gates[i] = tanh(P[i, 2c:2c+2] + kg),  Hs[i] = P[i,0:c] + b_s + (1-c_i) gates[i,1] wd[0:c],
 *   Ht[i] = P[i,c:2c] + b_t - c_i gates[i,0] wd[c:2c]          (c_i = is_src[i]),
 * i.e. lin_s(x_t2s) and lin_t(x_s2t) of the reference without materialising the shifted copies of x. */
int bgnn_adapted_transform_fwd_f32(const float* P, const uint8_t* is_src, const float* wd, const float* kg,
                                   const float* bias, int64_t n, int c, float* Hs, float* Ht, float* gates, void* stream);

/* Backward: gP [n, 2c+2] with row stride ldp >= 2c+2 (columns [0, 2c+2) fully written; a stride that is a multiple
 * of 4 lets bgnn_wgrad_gemm_f32 / bgnn_rowpanel_gemm_f32 read gP in place), red [4c+2] = (d wd [2c], d kg [2],
 * d bias [2c]).  Deterministic two-stage reductions. */
size_t bgnn_adapted_transform_bwd_workspace_bytes(int c);
int bgnn_adapted_transform_bwd_f32(const float* gHs, const float* gHt, const float* gates, const uint8_t* is_src,
                                   const float* wd, int64_t n, int c, int ldp, float* gP, float* g_wd_kg,
                                   void* workspace, size_t workspace_bytes, void* stream);
/* The same reductions, writing only the two gate columns of gP: dg [n, 4] = (gP[:, 2c], gP[:, 2c+1], -, -).  For a
 * conv whose input needs no gradient (the first layer) the other 2c columns of gP are gHs and gHt themselves, and
 * bgnn_wgrad_gemm_cat_f32 reads them in place: gP is never materialised. */
int bgnn_adapted_transform_bwd_gates_f32(const float* gHs, const float* gHt, const float* gates, const uint8_t* is_src,
                                         const float* wd, int64_t n, int c, float* dg, float* g_wd_kg, void* workspace,
                                         size_t workspace_bytes, void* stream);

/* The same transform for NARROW outputs (classifier heads, c <= 4; d % 4 == 0, d <= 256), without the dense
 * contraction on the host: reads x [n,d] once per direction.  wcat [2c+2, d] = [W_s; W_t; a_g_s2t[:d]; a_g_t2s[:d]],
 * bias [2c+2] or NULL, wd / kg as above.  bgnn_adapted_skinny_supported tells whether (c, d) is covered. */
int bgnn_adapted_skinny_supported(int c, int d);
int bgnn_adapted_skinny_fwd_f32(const float* x, const uint8_t* is_src, const float* wcat, const float* bias,
                                const float* wd, const float* kg, int64_t n, int d, int c, float* Hs, float* Ht,
                                float* gates, void* stream);
/* Backward: gx [n,d] (fully written), red [(2c+2)*d + (2c+2) + 2c] = (d wcat row-major, column sums of the
 * pre-activation gradient = (d bias [2c], d kg [2]), d wd [2c]).  Deterministic two-stage reductions. */
size_t bgnn_adapted_skinny_bwd_workspace_bytes(int c, int d);
int bgnn_adapted_skinny_bwd_f32(const float* x, const uint8_t* is_src, const float* wcat, const float* wd,
                                const float* gates, const float* gHs, const float* gHt, int64_t n, int d, int c,
                                float* gx, float* red, void* workspace, size_t workspace_bytes, void* stream);

/* sums [2,d]: column sums of x [n,d] over the source-domain rows (is_src != 0) and over the target-domain rows;
 * Delta of models/KTGNN.py:275-276 is sums[0]/Ns - sums[1]/Nt.  d % 4 == 0, d <= 1024.  Deterministic. */
size_t bgnn_domain_colsum_workspace_bytes(int d);
int bgnn_domain_colsum_f32(const float* x, const uint8_t* is_src, int64_t n, int d, float* sums, void* workspace,
                           size_t workspace_bytes, void* stream);

/* Row-panel GEMM on the tcgen05 tensor cores with fp32-grade accuracy (3 x TF32, the streamed operand split on
 * chip):  Y [n, no] (row stride ldy) = A [n, k] (row stride ld_a, ld_a % 4 == 0, 16-byte aligned) . B^T, with B
 * given as two planes b_hi / b_lo [nop, kp] row-major, nop = no rounded up to 16, kp = k rounded up to 32, zero
 * padded, b_hi = B rounded to tf32 (low 13 mantissa bits zero) and b_lo = B - b_hi rounded likewise.  no <= 256.
 * bias [no] (added to every row) or NULL.  Replaces the fp32 SIMT
 * GEMMs torch runs for x @ w_cat.t() and dP @ w_cat in AdaptedConv (models/KTGNN.py:277-284). */
int bgnn_rowpanel_gemm_supported(int k, int ld_a, int no);
int bgnn_rowpanel_gemm_f32(const float* A, int64_t n, int k, int ld_a, const float* b_hi, const float* b_lo,
                           const float* bias, int no, float* Y, int ldy, void* stream);
/* The same with an epilogue on the accumulator tile:  Y = act(A . B^T * scale + bias) + res.  scale, bias [no] or NULL;
 * act 0 none, 1 ReLU, 2 tanh; res [n, no] (row stride ld_res) or NULL.  One launch per dense layer of the embedding
 * producers that feed the build -- eval-mode BatchNorm folded into (scale, bias), the activation and the residual of
 * u = z' + biasatt(z') fused: models/models.py:852-893 (MLP), :70-99 / :124-127 (lin_self, biasatt), :1092-1096
 * (equavilent_trans_layer + Tanh). */
int bgnn_rowpanel_gemm_act_f32(const float* A, int64_t n, int k, int ld_a, const float* b_hi, const float* b_lo,
                               const float* scale, const float* bias, int act, const float* res, int ld_res, int no,
                               float* Y, int ldy, void* stream);
/* The same contraction for a WIDE AdaptedConv (c % 32 == 0, 2c+2 <= 256, d % 4 == 0, d <= 256) with the node-wise
 * epilogue of bgnn_adapted_transform_fwd_f32 fused in: x [n,d] is read once, P is never written.
 * wcat_hi / wcat_lo: planes of [W_s; W_t; a_g_s2t[:d]; a_g_t2s[:d]] as above; bias [2c] or NULL. */
int bgnn_adapted_wide_supported(int c, int d);
int bgnn_adapted_wide_fwd_f32(const float* x, int64_t n, int d, const float* wcat_hi, const float* wcat_lo, int c,
                              const uint8_t* is_src, const float* wd, const float* kg, const float* bias, float* Hs,
                              float* Ht, float* gates, void* stream);

/* Weight-gradient contraction W [no, d] (row stride ldw) = G^T X = sum_i G[i, :]^T X[i, :] over n rows, on the tcgen05
 * tensor cores with fp32-grade accuracy (3 x TF32, both operands split on chip, each read from HBM once; per-CTA
 * partials are added in a fixed order: deterministic).  G [n, no] row stride ld_g, X [n, d] row stride ld_x, both
 * strides multiples of 4 and both bases 16-byte aligned; d <= 128, no <= 256.  colsum [no] (NULL to skip; needs
 * d <= 96) receives the column sums of G from the same pass (a bias gradient).  This is g_w_cat = dP^T x of AdaptedConv
 * (models/KTGNN.py:277-284) and the weight gradient of the Linear layers of clf_transformer (models/KTGNN.py:363). */
int bgnn_wgrad_gemm_supported(int d, int ld_x, int no, int ld_g);
size_t bgnn_wgrad_gemm_workspace_bytes(int no);
int bgnn_wgrad_gemm_f32(const float* G, int ld_g, int no, const float* X, int ld_x, int d, int64_t n, float* W, int ldw,
                        float* colsum, void* workspace, size_t workspace_bytes, void* stream);
/* The same with G given as up to three column blocks side by side, G = [G0 | G1 | G2] (G1 / G2 may be NULL; every
 * block but the last a multiple of 32 columns wide; no = no0 + no1 + no2 sizes the workspace): each block is read
 * from its own matrix through its own TMA descriptor. */
int bgnn_wgrad_gemm_cat_f32(const float* G0, int ld0, int no0, const float* G1, int ld1, int no1, const float* G2, int ld2,
                            int no2, const float* X, int ld_x, int d, int64_t n, float* W, int ldw, float* colsum,
                            void* workspace, size_t workspace_bytes, void* stream);

/* BatchNorm1d (+ ReLU) over x [n, c] in training mode (models/KTGNN.py:363-366, 425-429: nn.BatchNorm1d followed by
 * ReLU), c % 4 == 0, c <= 1024; two passes over x each way, deterministic.
 * fwd: batch statistics (biased variance for the normalisation), y = [relu]((x - mean) * weight * invstd + bias);
 *      weight / bias [c] or NULL (= 1 / 0); running_mean / running_var [c] or NULL are updated in place with
 *      `momentum` (unbiased variance), like torch;  stats [4c] = (mean, invstd, weight * invstd, bias) is what apply
 *      and bwd take.
 * apply: y = [relu]((x - stats.mean) * stats.scale + stats.bias) only -- the eval-mode forward with
 *      stats = (running_mean, -, weight / sqrt(running_var + eps), bias).
 * bwd: gx [n, c] and gwb [2c] = (d weight, d bias) from gy = dL/dy; the ReLU mask is recomputed from x. */
int bgnn_bn_relu_supported(int c);
size_t bgnn_bn_relu_workspace_bytes(int c);
int bgnn_bn_relu_fwd_f32(const float* x, int64_t n, int c, const float* weight, const float* bias, float eps, float momentum,
                         float* running_mean, float* running_var, int relu, float* y, float* stats, void* workspace,
                         size_t workspace_bytes, void* stream);
int bgnn_bn_relu_apply_f32(const float* x, int64_t n, int c, const float* stats, int relu, float* y, void* stream);
int bgnn_bn_relu_bwd_f32(const float* gy, const float* x, int64_t n, int c, const float* stats, int relu, float* gx,
                         float* gwb, void* workspace, size_t workspace_bytes, void* stream);
/* Multi-GPU (rows partitioned over ranks): bgnn_bn_relu_fwd_f32 with y = NULL yields this rank's (mean, invstd) only;
 * the host combines the ranks' statistics, and bgnn_bn_relu_apply_f32 normalises with the global ones.  Backward in two
 * halves around an all-reduce: _reduce gives this rank's gwb [2c] = (sum g xhat | sum g) under the global statistics,
 * _apply takes coef [2c] = (mean g | mean g xhat) over the rows of all ranks. */
int bgnn_bn_relu_bwd_reduce_f32(const float* gy, const float* x, int64_t n, int c, const float* stats, int relu, float* gwb,
                                void* workspace, size_t workspace_bytes, void* stream);
int bgnn_bn_relu_bwd_apply_f32(const float* gy, const float* x, int64_t n, int c, const float* stats, int relu,
                               const float* coef, float* gx, void* stream);

/* The narrow transform for `heads` (1 or 2) convs that read the SAME x (the classifier heads clf_base / clf_target,
 * models/KTGNN.py:432-434): parameters stacked per head (wcat [heads*(2c+2), d], bias [heads*(2c+2)] or NULL,
 * wd [heads*2c], kg [heads*2]), outputs side by side (Hs, Ht [n, heads*c], gates [n, heads*2]) -- x is read once.
 * Backward in two calls: _pre (no pass over x) gives pre [heads*(2c+2)] = per head (d wd [2c], d kg [2]); the host
 * turns that into gm [2, d], the gradient reaching x through the source / target domain means, and _bwd adds
 * gm[0] (source rows) / gm[1] (target rows) into gx in its single pass over x (gm may be NULL).
 * red [heads*((2c+2)*d + (2c+2) + 2c)] = (d wcat, per head column sums (d bias, d kg), d wd). */
int bgnn_adapted_skinny_heads_supported(int c, int d, int heads);
int bgnn_adapted_skinny_heads_fwd_f32(const float* x, const uint8_t* is_src, const float* wcat, const float* bias,
                                      const float* wd, const float* kg, int64_t n, int d, int c, int heads, float* Hs,
                                      float* Ht, float* gates, void* stream);
size_t bgnn_adapted_skinny_heads_bwd_workspace_bytes(int c, int d, int heads);
int bgnn_adapted_skinny_heads_pre_f32(const uint8_t* is_src, const float* wd, const float* gates, const float* gHs,
                                      const float* gHt, int64_t n, int c, int heads, float* pre, void* workspace,
                                      size_t workspace_bytes, void* stream);
int bgnn_adapted_skinny_heads_bwd_f32(const float* x, const uint8_t* is_src, const float* wcat, const float* wd,
                                      const float* gates, const float* gHs, const float* gHt, const float* gm, int64_t n,
                                      int d, int c, int heads, float* gx, float* red, void* workspace,
                                      size_t workspace_bytes, void* stream);

/* Forward of bgnn_adapted_skinny_heads_fwd_f32 on the tensor cores: the heads * (2c+2) <= 20 dot products per row run
 * as a 3 x TF32 row-panel GEMM (x streamed by TMA, split on chip) and the gates / corrections are applied to the
 * accumulator row; same outputs and layouts.  wcat_hi / wcat_lo: tf32 planes of the stacked weight rows, padded to
 * (16, 32) multiples as for bgnn_rowpanel_gemm_f32. */
int bgnn_adapted_skinny_tc_supported(int c, int d, int heads);
int bgnn_adapted_skinny_heads_tc_fwd_f32(const float* x, int64_t n, int d, const float* wcat_hi, const float* wcat_lo, int c,
                                         int heads, const uint8_t* is_src, const float* wd, const float* kg,
                                         const float* bias, float* Hs, float* Ht, float* gates, void* stream);

/* Operand preparation for the 3 x TF32 kernels above: hi [rows_p, cols_p] = w rounded to tf32, lo = (w - hi) rounded
 * to tf32, zero outside [rows, cols].  w is addressed as w[r * stride_r + c * stride_c] (elements), so a transposed
 * view needs no copy. */
int bgnn_tf32_planes_f32(const float* w, int rows, int cols, int64_t stride_r, int64_t stride_c, int rows_p, int cols_p,
                         float* hi, float* lo, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BGNN_B200_H_ */
