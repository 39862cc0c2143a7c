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

enum { MODE_DOT = 0, MODE_ADDRELU = 1 };

__device__ __forceinline__ void load_slice(float (*dst)[ST_LD], const float* __restrict__ src, const int* rows,
                                           int row0, int nrows, int d, int h0, int tid) {
  // 64 rows x 16 features; thread -> (row = tid/4, 4 consecutive features); stored transposed [h][row].
  int r = tid >> 2, hq = (tid & 3) * 4;
  int gr = row0 + r;
  bool rok = gr < nrows;
  long long grow = rok ? (rows ? (long long)rows[gr] : (long long)gr) : 0;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (rok) {
    const float* p = src + grow * d + h0 + hq;
    if ((d & 3) == 0 && h0 + hq + 3 < d) {
      float4 t = __ldg(reinterpret_cast<const float4*>(p));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (h0 + hq + c < d) v[c] = __ldg(p + c);
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) dst[hq + c][r] = v[c];
}

template <int MODE>
__global__ void __launch_bounds__(ST_THREADS)
knn_simt_kernel(const float* __restrict__ Q, int nq, const float* __restrict__ DB, int ndb, int d,
                const float* __restrict__ w, float bias, int apply_sigmoid, int kc, int db_per_split,
                const int* __restrict__ row_list, const int* __restrict__ row_count,
                float* __restrict__ cand_val, int* __restrict__ cand_idx) {
  __shared__ __align__(16) float sq[ST_KB][ST_LD];
  __shared__ __align__(16) float sd[ST_KB][ST_LD];
  __shared__ float sw[ST_KB];
  __shared__ float stile[ST_TQ][ST_TD + 1];
  extern __shared__ __align__(16) unsigned char dyn[];   // lists: val[kc][64], idx[kc][64]
  float* lval = reinterpret_cast<float*>(dyn);
  int* lidx = reinterpret_cast<int*>(dyn + (size_t)kc * ST_TQ * sizeof(float));

  const int tid = threadIdx.x;
  const int nrows = row_list ? min(*row_count, nq) : nq;
  const int q0 = blockIdx.x * ST_TQ;
  if (q0 >= nrows) return;
  const int split = blockIdx.y;
  const int db_begin = split * db_per_split;
  const int db_end = min(ndb, db_begin + db_per_split);
  const int ty = tid >> 4, tx = tid & 15;

  ListState st = list_init();
  for (int t0 = db_begin; t0 < db_end; t0 += ST_TD) {
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

    for (int h0 = 0; h0 < d; h0 += ST_KB) {
      __syncthreads();
      load_slice(sq, Q, row_list, q0, nrows, d, h0, tid);
      load_slice(sd, DB, nullptr, t0, db_end, d, h0, tid);
      if (MODE == MODE_ADDRELU && tid < ST_KB) sw[tid] = (h0 + tid < d) ? __ldg(w + h0 + tid) : 0.f;
      __syncthreads();
#pragma unroll
      for (int h = 0; h < ST_KB; ++h) {
        float4 a4 = *reinterpret_cast<const float4*>(&sq[h][ty * 4]);
        float4 b4 = *reinterpret_cast<const float4*>(&sd[h][tx * 4]);
        float av[4] = {a4.x, a4.y, a4.z, a4.w};
        float bv[4] = {b4.x, b4.y, b4.z, b4.w};
        if (MODE == MODE_DOT) {
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
        if (s > st.thr) st = list_insert(lval + tid, lidx + tid, ST_TQ, kc, st, s, t0 + c);
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
}

int launch_knn_simt(int mode, const float* Q, int nq, const float* DB, int ndb, int d, const float* w, float bias,
                    int apply_sigmoid, int kc, int nsplit, int db_per_split, const int* row_list,
                    const int* row_count, float* cand_val, int* cand_idx, cudaStream_t stream) {
  if (nq <= 0) return BGNN_OK;
  dim3 grid((nq + ST_TQ - 1) / ST_TQ, nsplit);
  size_t dyn = (size_t)kc * ST_TQ * (sizeof(float) + sizeof(int));
  if (mode == MODE_DOT) {
    BGNN_CUDA_TRY(cudaFuncSetAttribute(knn_simt_kernel<MODE_DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    knn_simt_kernel<MODE_DOT><<<grid, ST_THREADS, dyn, stream>>>(Q, nq, DB, ndb, d, w, bias, apply_sigmoid, kc,
                                                                   db_per_split, row_list, row_count, cand_val, cand_idx);
  } else {
    BGNN_CUDA_TRY(cudaFuncSetAttribute(knn_simt_kernel<MODE_ADDRELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    knn_simt_kernel<MODE_ADDRELU><<<grid, ST_THREADS, dyn, stream>>>(Q, nq, DB, ndb, d, w, bias, apply_sigmoid, kc,
                                                                       db_per_split, row_list, row_count, cand_val, cand_idx);
  }
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

// ---------------------------------------------------------------------------------------------
// Row-normalisation prologue:  out = x / max(||x||_2, 1e-8)  (ATen cosine_similarity, eps 1e-8),
// optionally split into tf32-exact hi and the fp32 remainder lo (x == hi + lo exactly).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
normalize_split_kernel(const float* __restrict__ x, long long n, int d, int normalize, float* __restrict__ hi,
                       float* __restrict__ lo) {
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  int lane = threadIdx.x & 31;
  const float* p = x + row * d;
  float scale = 1.f;
  if (normalize) {
    float ss = 0.f;
    for (int c = lane; c < d; c += 32) { float v = __ldg(p + c); ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    scale = 1.0f / fmaxf(sqrtf(ss), 1e-8f);
  }
  for (int c = lane; c < d; c += 32) {
    float v = __ldg(p + c) * scale;
    if (lo) {
      float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
      hi[row * d + c] = h;
      lo[row * d + c] = v - h;
    } else {
      hi[row * d + c] = v;
    }
  }
}

int launch_normalize_split(const float* x, long long n, int d, int normalize, float* hi, float* lo, cudaStream_t stream) {
  if (n <= 0) return BGNN_OK;
  long long blocks = (n + 7) / 8;
  normalize_split_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, n, d, normalize, hi, lo);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

// ---------------------------------------------------------------------------------------------
// Merge / finalize: one warp per query row.
//   gathers nlists*kc candidates, optionally re-scores them exactly in fp32 from the split operands
//   (x = hi + lo), applies sigmoid, selects the top-k under (value desc, index asc), writes int64
//   indices + fp32 values sorted best-first and gap = v_k - v_{k+1}; in certified mode it checks that
//   no discarded column could beat the k-th best (approximate threshold + delta) and appends rows
//   that fail to the exact-fallback list.
// ---------------------------------------------------------------------------------------------
constexpr int MG_WARPS = 4;
constexpr int MG_MAXC = 1024;  // candidates per row held in shared memory

__global__ void __launch_bounds__(MG_WARPS * 32)
knn_merge_kernel(const float* __restrict__ cand_val, const int* __restrict__ cand_idx, int nlists, int kc, int nq, int k,
                 int rescore, const float* __restrict__ qhi, const float* __restrict__ qlo,
                 const float* __restrict__ dhi, const float* __restrict__ dlo, int d, int apply_sigmoid,
                 float delta, const int* __restrict__ row_list, const int* __restrict__ row_count,
                 long long* __restrict__ out_idx, float* __restrict__ out_val, float* __restrict__ out_gap,
                 int* __restrict__ fb_rows, int* __restrict__ fb_count) {
  __shared__ float sv[MG_WARPS][MG_MAXC];
  __shared__ int si[MG_WARPS][MG_MAXC];
  extern __shared__ float sqrow[];  // [MG_WARPS][d] when rescoring
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nrows = row_list ? min(*row_count, nq) : nq;
  long long r = (long long)blockIdx.x * MG_WARPS + wid;
  if (r >= nrows) return;
  const long long row = row_list ? (long long)row_list[r] : r;
  const int total = nlists * kc;
  float* v = sv[wid];
  int* ix = si[wid];
  float worst_thr = -INFINITY;  // max over full lists of their approximate threshold (certified mode)
  for (int l = 0; l < nlists; ++l) {
    long long base = ((long long)l * nq + row) * kc;
    float lmin = INFINITY;
    bool full = true;
    for (int s = lane; s < kc; s += 32) {
      float cv = cand_val[base + s];
      int ci = cand_idx[base + s];
      v[l * kc + s] = cv;
      ix[l * kc + s] = ci;
      if (ci < 0) full = false; else lmin = fminf(lmin, cv);
    }
    full = __all_sync(0xffffffffu, full);
    lmin = -warp_max(-lmin);
    if (full) worst_thr = fmaxf(worst_thr, lmin);
  }
  __syncwarp();
  if (rescore) {
    float* qr = sqrow + (size_t)wid * d;
    for (int c = lane; c < d; c += 32) qr[c] = qhi[row * d + c] + (qlo ? qlo[row * d + c] : 0.f);
    __syncwarp();
    for (int c = lane; c < total; c += 32) {
      int j = ix[c];
      if (j >= 0) {
        const float* ph = dhi + (long long)j * d;
        const float* pl = dlo ? dlo + (long long)j * d : nullptr;
        float acc = 0.f;
        for (int h = 0; h < d; ++h) {
          float x = ph[h] + (pl ? pl[h] : 0.f);
          acc = fmaf(qr[h], x, acc);
        }
        v[c] = acc;
      }
    }
    __syncwarp();
  }
  if (apply_sigmoid && rescore) {
    for (int c = lane; c < total; c += 32)
      if (ix[c] >= 0) v[c] = sigmoid_f32(v[c]);
    __syncwarp();
  }
  // k+1 rounds of warp arg-best under (value desc, index asc)
  float prev = 0.f, kth = -INFINITY;
  for (int round = 0; round <= k; ++round) {
    float bv = -INFINITY;
    int bi = 0x7fffffff, bs = -1;
    for (int c = lane; c < total; c += 32) {
      int j = ix[c];
      if (j < 0) continue;
      float cv = v[c];
      if (cv > bv || (cv == bv && j < bi)) { bv = cv; bi = j; bs = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      int os = __shfl_xor_sync(0xffffffffu, bs, o);
      if (os >= 0 && (bs < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; bs = os; }
    }
    if (round < k) {
      if (lane == 0) {
        out_idx[row * k + round] = (bs >= 0) ? (long long)bi : -1LL;
        out_val[row * k + round] = (bs >= 0) ? bv : -INFINITY;
      }
      prev = bv;
      if (round == k - 1) kth = (bs >= 0) ? bv : -INFINITY;
      if (bs >= 0 && (bs & 31) == lane) ix[bs] = -1;   // consume (bs % 32 == owning lane)
      __syncwarp();
      // duplicates of the same db index from overlapping lists cannot occur: lists cover disjoint columns
    } else if (lane == 0 && out_gap) {
      out_gap[row] = (bs >= 0) ? (prev - bv) : INFINITY;
    }
  }
  if (delta >= 0.f && fb_rows && lane == 0) {
    // Any column not kept by a full list scored (approximately) <= that list's threshold.
    float bound = worst_thr + delta;
    if (apply_sigmoid) bound = sigmoid_f32(bound);
    if (worst_thr > -INFINITY && !(bound < kth)) {
      int slot = atomicAdd(fb_count, 1);
      fb_rows[slot] = (int)row;
    }
  }
}

int launch_knn_merge(const float* cand_val, const int* cand_idx, int nlists, int kc, int nq, int k, int rescore,
                     const float* qhi, const float* qlo, const float* dhi, const float* dlo, int d, int apply_sigmoid,
                     float delta, const int* row_list, const int* row_count, long long* out_idx, float* out_val,
                     float* out_gap, int* fb_rows, int* fb_count, cudaStream_t stream) {
  if (nq <= 0) return BGNN_OK;
  if (nlists * kc > MG_MAXC) return BGNN_ERR_UNSUPPORTED;
  size_t dyn = rescore ? (size_t)MG_WARPS * d * sizeof(float) : 0;
  knn_merge_kernel<<<(nq + MG_WARPS - 1) / MG_WARPS, MG_WARPS * 32, dyn, stream>>>(
      cand_val, cand_idx, nlists, kc, nq, k, rescore, qhi, qlo, dhi, dlo, d, apply_sigmoid, delta, row_list,
      row_count, out_idx, out_val, out_gap, fb_rows, fb_count);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

}  // namespace bgnn
