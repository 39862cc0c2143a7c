// Exact fp32 all-pairs similarity + per-row top-k on CUDA cores.
//
// Two pair kernels share one tiled sweep:
//   MODE_DOT     score = sum_h Q[i,h] * DB[j,h]                       (cosine head on unit rows;
//                reference: models/models.py:124-130, 945-948)
//   MODE_ADDRELU score = sum_h w[h] * relu(Q[i,h] + DB[j,h]) + bias   (eval-mode fold of the v2 'mlp'
//                head, models/models.py:918-925, 949-954; no GEMM form)
// followed by sigmoid (models.py:129, 955) and a per-row top-k under the parity key
// (fp32 post-sigmoid value desc, index asc), replacing sim_mat.topk at main_bridged_graph.py:60,104.
// The [nq, ndb] similarity matrix is never written to HBM.
//
// The per-pair sum is a single fmaf chain over h = 0..d-1; knn_select.cu re-scores tensor-core
// nominees with the same chain, so both routes produce bit-identical similarities.
//
// It is the shipping path for the add-ReLU head, the exact fallback for rows the tensor-core path
// cannot certify, and the small-problem path for the cosine head.
#include "common.cuh"
#include "kernels.h"

namespace bgnn {

constexpr int ST_TQ = 64;    // query rows per CTA tile
constexpr int ST_TD = 64;    // db rows per tile step
constexpr int ST_KB = 16;    // feature slice
constexpr int ST_LD = 68;    // padded leading dim of the transposed slices (multiple of 4 -> float4 reads)
constexpr int ST_THREADS = 256;

// 64 rows x 16 features of (hi [+ lo]); thread -> (row = tid/4, 4 consecutive features); stored
// transposed [h][row].  `rows` optionally maps tile rows to matrix rows (exact-fallback row list).
__device__ __forceinline__ void load_slice(float (*dst)[ST_LD], const float* __restrict__ hi,
                                           const float* __restrict__ lo, const int* __restrict__ rows, int row0,
                                           int nrows, int d, int ld, bool vec_ok, int h0, int tid) {
  const int r = tid >> 2, hq = (tid & 3) * 4;
  const int gr = row0 + r;
  const bool rok = gr < nrows;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (rok) {
    const long long grow = rows ? (long long)__ldg(rows + gr) : (long long)gr;
    const long long off = grow * ld + h0 + hq;
    if (vec_ok && h0 + hq + 3 < d) {
      float4 t = __ldg(reinterpret_cast<const float4*>(hi + off));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      if (lo) {
        float4 u = __ldg(reinterpret_cast<const float4*>(lo + off));
        v[0] += u.x; v[1] += u.y; v[2] += u.z; v[3] += u.w;
      }
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (h0 + hq + c < d) v[c] = __ldg(hi + off + c) + (lo ? __ldg(lo + off + c) : 0.f);
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) dst[hq + c][r] = v[c];
}

template <int MODE>
__global__ void __launch_bounds__(ST_THREADS)
knn_simt_kernel(const float* __restrict__ Q, const float* __restrict__ Qlo, int nq, const float* __restrict__ DB,
                const float* __restrict__ DBlo, int ndb, int d, int ld, int vec_ok, const float* __restrict__ w,
                float bias, int apply_sigmoid, int kc, int db_per_split, const int* __restrict__ row_list,
                const int* __restrict__ row_count, int few_rows, float* __restrict__ cand_val, int* __restrict__ cand_idx) {
  __shared__ __align__(16) float sq[ST_KB][ST_LD];
  __shared__ __align__(16) float sd[ST_KB][ST_LD];
  __shared__ float sw[ST_KB];
  __shared__ float stile[ST_TQ][ST_TD + 1];
  extern __shared__ __align__(16) unsigned char dyn[];   // lists: val[kc][64], idx[kc][64]
  float* lval = reinterpret_cast<float*>(dyn);
  int* lidx = reinterpret_cast<int*>(dyn + (size_t)kc * ST_TQ * sizeof(float));

  const int tid = threadIdx.x;
  const int nrows = row_list ? min(__ldg(row_count), nq) : nq;
  if (row_list && nrows <= few_rows) return;      // a handful of rows: knn_exact_rows_kernel does them
  const int split = blockIdx.y;
  const int db_begin = split * db_per_split;
  const int db_end = min(ndb, db_begin + db_per_split);
  const int ty = tid >> 4, tx = tid & 15;

  // grid-stride over query tiles: the exact-fallback launch does not know its row count on the host
  for (int q0 = blockIdx.x * ST_TQ; q0 < nrows; q0 += gridDim.x * ST_TQ) {
  ListState st = list_init();
  for (int t0 = db_begin; t0 < db_end; t0 += ST_TD) {
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

    for (int h0 = 0; h0 < d; h0 += ST_KB) {
      __syncthreads();
      load_slice(sq, Q, Qlo, row_list, q0, nrows, d, ld, vec_ok != 0, h0, tid);
      load_slice(sd, DB, DBlo, nullptr, t0, db_end, d, ld, vec_ok != 0, h0, tid);
      if (MODE == BGNN_PAIR_ADDRELU && tid < ST_KB) sw[tid] = (h0 + tid < d) ? __ldg(w + h0 + tid) : 0.f;
      __syncthreads();
#pragma unroll
      for (int h = 0; h < ST_KB; ++h) {
        float4 a4 = *reinterpret_cast<const float4*>(&sq[h][ty * 4]);
        float4 b4 = *reinterpret_cast<const float4*>(&sd[h][tx * 4]);
        float av[4] = {a4.x, a4.y, a4.z, a4.w};
        float bv[4] = {b4.x, b4.y, b4.z, b4.w};
        if (MODE == BGNN_PAIR_DOT) {
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
        } else {
          float wh = sw[h];
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(wh, fmaxf(av[a] + bv[b], 0.f), acc[a][b]);
        }
      }
    }
    // stage the tile's scores so that one thread per query row can scan them in index order
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        float s = acc[a][b] + bias;
        if (apply_sigmoid) s = sigmoid_f32(s);
        stile[ty * 4 + a][tx * 4 + b] = s;
      }
    __syncthreads();
    if (tid < ST_TQ && q0 + tid < nrows) {
      int lim = min(ST_TD, db_end - t0);
      for (int c = 0; c < lim; ++c) {
        float s = stile[tid][c];
        if (s > st.thr) st = list_insert(smem_addr(lval + tid), smem_addr(lidx + tid), ST_TQ * 4, kc, st, s, t0 + c);
      }
    }
  }
  if (tid < ST_TQ && q0 + tid < nrows) {
    long long grow = row_list ? (long long)row_list[q0 + tid] : (long long)(q0 + tid);
    long long base = ((long long)split * nq + grow) * kc;
    for (int s = 0; s < kc; ++s) {
      bool f = s < st.cnt;
      cand_val[base + s] = f ? lval[s * ST_TQ + tid] : -INFINITY;
      cand_idx[base + s] = f ? lidx[s * ST_TQ + tid] : -1;
    }
  }
  __syncthreads();
  }  // query tiles
}

int launch_knn_simt(int mode, const float* Q, const float* Qlo, int nq, const float* DB, const float* DBlo, int ndb,
                    int d, int ld, const float* w, float bias, int apply_sigmoid, int kc, int nsplit, int db_per_split,
                    const int* row_list, const int* row_count, int few_rows, float* cand_val, int* cand_idx,
                    cudaStream_t stream) {
  if (nq <= 0 || nsplit <= 0) return BGNN_OK;
  int qtiles = (nq + ST_TQ - 1) / ST_TQ;
  if (row_list) qtiles = qtiles < 16 ? qtiles : 16;     // fallback rows are few; CTAs stride over the list
  dim3 grid(qtiles, nsplit);
  size_t dyn = (size_t)kc * ST_TQ * (sizeof(float) + sizeof(int));
  if (dyn > 160 * 1024) return BGNN_ERR_UNSUPPORTED;
  auto aligned = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const int vec_ok = (ld % 4 == 0) && aligned(Q) && aligned(Qlo) && aligned(DB) && aligned(DBlo);
  if (mode == BGNN_PAIR_DOT) {
    auto kern = knn_simt_kernel<BGNN_PAIR_DOT>;
    BGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    kern<<<grid, ST_THREADS, dyn, stream>>>(Q, Qlo, nq, DB, DBlo, ndb, d, ld, vec_ok, w, bias, apply_sigmoid, kc,
                                            db_per_split, row_list, row_count, few_rows, cand_val, cand_idx);
  } else {
    auto kern = knn_simt_kernel<BGNN_PAIR_ADDRELU>;
    BGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    kern<<<grid, ST_THREADS, dyn, stream>>>(Q, Qlo, nq, DB, DBlo, ndb, d, ld, vec_ok, w, bias, apply_sigmoid, kc,
                                            db_per_split, row_list, row_count, few_rows, cand_val, cand_idx);
  }
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

// ---- exact sweep for a HANDFUL of rows -----------------------------------------------------------------------
// The tensor-core build leaves 0-2 uncertified rows out of 2.6e5; the tiled sweep above spends 3.9 ms on them
// (64-row tiles with one live row, 48 CTAs).  Here every CTA takes a slice of the db for each listed row: one
// thread per db row (the same fmaf chain over h, so similarities are bit-identical), the slice's scores staged
// in shared memory, k1 rounds of block arg-best under the parity key; a second kernel merges the per-CTA lists.
constexpr int XR_THREADS = 256;
constexpr int XR_CTAS = kNumSMs * 2;
constexpr int XR_MAX_ROWS = 8;

__device__ __forceinline__ bool xr_better(float av, int ai, float bv, int bi) {   // a ranks above b
  return av > bv || (av == bv && ai < bi);
}

// block arg-best over (val[i], idx[i]) for i in [0, m): result in all threads
__device__ __forceinline__ void xr_block_best(const float* val, const int* idx, int m, float* s_v, int* s_i, int* s_p,
                                              float& bv, int& bi, int& bp) {
  bv = -INFINITY; bi = 0x7fffffff; bp = -1;
  for (int i = threadIdx.x; i < m; i += XR_THREADS) {
    const int j = idx[i];
    if (j < 0) continue;
    const float v = val[i];
    if (bp < 0 || xr_better(v, j, bv, bi)) { bv = v; bi = j; bp = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    const int op = __shfl_xor_sync(0xffffffffu, bp, o);
    if (op >= 0 && (bp < 0 || xr_better(ov, oi, bv, bi))) { bv = ov; bi = oi; bp = op; }
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { s_v[w] = bv; s_i[w] = bi; s_p[w] = bp; }
  __syncthreads();
  bv = s_v[0]; bi = s_i[0]; bp = s_p[0];
#pragma unroll
  for (int k = 1; k < XR_THREADS / 32; ++k)
    if (s_p[k] >= 0 && (bp < 0 || xr_better(s_v[k], s_i[k], bv, bi))) { bv = s_v[k]; bi = s_i[k]; bp = s_p[k]; }
  __syncthreads();
}

__global__ void __launch_bounds__(XR_THREADS)
knn_exact_rows_kernel(const float* __restrict__ Q, const float* __restrict__ Qlo, const float* __restrict__ DB,
                      const float* __restrict__ DBlo, int ndb, int d, int ld, int apply_sigmoid, int k1,
                      const int* __restrict__ row_list, const int* __restrict__ row_count, float* __restrict__ part_val,
                      int* __restrict__ part_idx) {
  extern __shared__ __align__(16) float xr_smem[];   // q row [ld], slice scores [slice], slice ids [slice]
  __shared__ float s_v[XR_THREADS / 32];
  __shared__ int s_i[XR_THREADS / 32], s_p[XR_THREADS / 32];
  const int count = __ldg(row_count);
  if (count <= 0 || count > XR_MAX_ROWS) return;
  const int slice = (ndb + gridDim.x - 1) / gridDim.x;
  const int begin = blockIdx.x * slice, end = min(ndb, begin + slice);
  float* sq = xr_smem;
  float* sv = xr_smem + ld;
  int* si = reinterpret_cast<int*>(sv + slice);
  for (int r = 0; r < count; ++r) {
    const long long row = row_list[r];
    for (int c = threadIdx.x; c < ld; c += XR_THREADS) sq[c] = Q[row * ld + c] + (Qlo ? Qlo[row * ld + c] : 0.f);
    __syncthreads();
    for (int j = begin + threadIdx.x; j < end; j += XR_THREADS) {
      const float4* ph = reinterpret_cast<const float4*>(DB + (long long)j * ld);
      const float4* pl = DBlo ? reinterpret_cast<const float4*>(DBlo + (long long)j * ld) : nullptr;
      float acc = 0.f;
      for (int h4 = 0; h4 < d / 4; ++h4) {
        float4 x = __ldg(ph + h4);
        if (pl) { const float4 y = __ldg(pl + h4); x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }
        const float4 q = *reinterpret_cast<const float4*>(sq + 4 * h4);
        acc = fmaf(q.x, x.x, acc);
        acc = fmaf(q.y, x.y, acc);
        acc = fmaf(q.z, x.z, acc);
        acc = fmaf(q.w, x.w, acc);
      }
      sv[j - begin] = apply_sigmoid ? sigmoid_f32(acc) : acc;
      si[j - begin] = j;
    }
    __syncthreads();
    const int m = max(0, end - begin);
    for (int round = 0; round < k1; ++round) {
      float bv; int bi, bp;
      xr_block_best(sv, si, m, s_v, s_i, s_p, bv, bi, bp);
      if (threadIdx.x == 0) {
        const long long o = ((long long)r * gridDim.x + blockIdx.x) * k1 + round;
        part_val[o] = bp >= 0 ? bv : -INFINITY;
        part_idx[o] = bp >= 0 ? bi : -1;
        if (bp >= 0) si[bp] = -1;                 // consumed
      }
      __syncthreads();
    }
  }
}

// one CTA per listed row: top-k (+1 for the gap) of the nparts * k1 per-slice winners
__global__ void __launch_bounds__(XR_THREADS)
knn_exact_rows_merge_kernel(const float* __restrict__ part_val, const int* __restrict__ part_idx, int nparts, int k1, int k,
                            const int* __restrict__ row_list, const int* __restrict__ row_count,
                            long long* __restrict__ out_idx, float* __restrict__ out_val, float* __restrict__ out_gap,
                            float eps, int* __restrict__ out_count) {
  extern __shared__ __align__(16) float xm_smem[];   // values [nparts*k1], ids [nparts*k1]
  __shared__ float s_v[XR_THREADS / 32];
  __shared__ int s_i[XR_THREADS / 32], s_p[XR_THREADS / 32];
  const int count = __ldg(row_count);
  const int r = blockIdx.x;
  if (count > XR_MAX_ROWS || r >= count) return;
  const int m = nparts * k1;
  float* sv = xm_smem;
  int* si = reinterpret_cast<int*>(sv + m);
  for (int i = threadIdx.x; i < m; i += XR_THREADS) {
    sv[i] = part_val[(long long)r * m + i];
    si[i] = part_idx[(long long)r * m + i];
  }
  __syncthreads();
  const long long row = row_list[r];
  float vk = -INFINITY;
  const bool use_eps = eps == eps;                  // epsilon threshold fused into the selection (NaN = off)
  int kept = 0;
  for (int round = 0; round <= k; ++round) {
    float bv; int bi, bp;
    xr_block_best(sv, si, m, s_v, s_i, s_p, bv, bi, bp);
    if (threadIdx.x == 0) {
      if (round < k) {
        const bool emit = bp >= 0 && (!use_eps || bv > eps);
        kept += emit ? 1 : 0;
        out_idx[row * k + round] = emit ? (long long)bi : -1LL;
        out_val[row * k + round] = bp >= 0 ? bv : -INFINITY;
        if (round == k - 1) vk = bp >= 0 ? bv : -INFINITY;
        if (bp >= 0) si[bp] = -1;
      } else {
        if (out_gap) out_gap[row] = bp >= 0 ? (vk - bv) : INFINITY;
        if (out_count) out_count[row] = kept;
      }
    }
    __syncthreads();
  }
}

constexpr int XR_SMEM_MAX = 200 * 1024;

// CTAs (= db slices) of the few-rows sweep: enough that a slice's scores fit in shared memory, few enough that
// the merge kernel can hold all per-slice winners; 0 = this shape is left to the tiled sweep.
static int xr_ctas(int ndb, int ld, int k) {
  const int k1 = k + 1;
  const int max_slice = (XR_SMEM_MAX - ld * 4) / 8;
  int ctas = XR_CTAS;
  while (ctas > 1 && (ndb + ctas - 1) / ctas < XR_THREADS) ctas >>= 1;   // at least one db row per thread
  if ((ndb + ctas - 1) / ctas > max_slice) ctas = (ndb + max_slice - 1) / max_slice;
  if ((size_t)2 * ctas * k1 * sizeof(float) > (size_t)XR_SMEM_MAX) return 0;
  return ctas;
}

int knn_exact_rows_max(int ndb, int ld, int k) { return xr_ctas(ndb, ld, k) > 0 ? XR_MAX_ROWS : 0; }

size_t knn_exact_rows_workspace_bytes(int k) {
  // per-slice lists: at most XR_SMEM_MAX / 8 entries per listed row (the merge kernel's shared-memory bound)
  return align_up((size_t)XR_MAX_ROWS * (XR_SMEM_MAX / 8) * sizeof(float), 256) * 2 + 256;
}

// Handles the listed rows iff 0 < *row_count <= knn_exact_rows_max(..) (decided on the device; no-op otherwise).
int launch_knn_exact_rows(const float* Q, const float* Qlo, const float* DB, const float* DBlo, int ndb, int d, int ld,
                          int apply_sigmoid, int k, const int* row_list, const int* row_count, long long* out_idx,
                          float* out_val, float* out_gap, float eps, int* out_count, void* ws, size_t ws_bytes,
                          cudaStream_t stream) {
  if (ld % 4 != 0 || d % 4 != 0) return BGNN_ERR_INVALID_ARG;
  if (ws_bytes < knn_exact_rows_workspace_bytes(k)) return BGNN_ERR_WORKSPACE;
  const int ctas = xr_ctas(ndb, ld, k);
  if (ctas == 0) return BGNN_OK;                     // the caller passes few_rows = 0 to the tiled sweep
  const int k1 = k + 1;
  Workspace w(ws, ws_bytes);
  float* pv = w.take<float>((size_t)XR_MAX_ROWS * ctas * k1);
  int* pi = w.take<int>((size_t)XR_MAX_ROWS * ctas * k1);
  if (!w.ok()) return BGNN_ERR_WORKSPACE;
  const int slice = (ndb + ctas - 1) / ctas;
  const size_t dyn1 = ((size_t)ld + 2 * (size_t)slice) * sizeof(float);
  const size_t dyn2 = (size_t)2 * ctas * k1 * sizeof(float);
  BGNN_CUDA_TRY(cudaFuncSetAttribute(knn_exact_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn1));
  BGNN_CUDA_TRY(cudaFuncSetAttribute(knn_exact_rows_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn2));
  knn_exact_rows_kernel<<<ctas, XR_THREADS, dyn1, stream>>>(Q, Qlo, DB, DBlo, ndb, d, ld, apply_sigmoid, k1, row_list,
                                                           row_count, pv, pi);
  BGNN_LAUNCH_CHECK();
  knn_exact_rows_merge_kernel<<<XR_MAX_ROWS, XR_THREADS, dyn2, stream>>>(pv, pi, ctas, k1, k, row_list, row_count, out_idx,
                                                                        out_val, out_gap, eps, out_count);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

}  // namespace bgnn
