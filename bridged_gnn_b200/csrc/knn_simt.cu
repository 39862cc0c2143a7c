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
                const int* __restrict__ row_count, float* __restrict__ cand_val, int* __restrict__ cand_idx) {
  __shared__ __align__(16) float sq[ST_KB][ST_LD];
  __shared__ __align__(16) float sd[ST_KB][ST_LD];
  __shared__ float sw[ST_KB];
  __shared__ float stile[ST_TQ][ST_TD + 1];
  extern __shared__ __align__(16) unsigned char dyn[];   // lists: val[kc][64], idx[kc][64]
  float* lval = reinterpret_cast<float*>(dyn);
  int* lidx = reinterpret_cast<int*>(dyn + (size_t)kc * ST_TQ * sizeof(float));

  const int tid = threadIdx.x;
  const int nrows = row_list ? min(__ldg(row_count), nq) : nq;
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
                    const int* row_list, const int* row_count, float* cand_val, int* cand_idx, cudaStream_t stream) {
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
                                            db_per_split, row_list, row_count, cand_val, cand_idx);
  } else {
    auto kern = knn_simt_kernel<BGNN_PAIR_ADDRELU>;
    BGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    kern<<<grid, ST_THREADS, dyn, stream>>>(Q, Qlo, nq, DB, DBlo, ndb, d, ld, vec_ok, w, bias, apply_sigmoid, kc,
                                            db_per_split, row_list, row_count, cand_val, cand_idx);
  }
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

}  // namespace bgnn
