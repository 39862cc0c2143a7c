// Prologue and finalisation kernels of the kNN build.
//
//  normalize_split_kernel  x -> x / max(||x||_2, 1e-8)  (ATen cosine_similarity's eps; reference call
//                          models/models.py:128, 948), written with row stride `ld` (zero padded) and
//                          optionally split into a tf32-exact `hi` plane and the fp32 remainder `lo`
//                          (hi + lo == x exactly) for the tensor-core sweep.
//  knn_merge_kernel        one warp per query row: gathers the per-split candidate lists, optionally
//                          re-scores the nominees exactly (the same fmaf chain as knn_simt.cu), applies
//                          sigmoid, selects the top-k under the parity key (value desc, index asc)
//                          -- the replacement of sim_mat.topk at main_bridged_graph.py:60,104 -- and,
//                          for tensor-core nominees, certifies the row against the approximation bound;
//                          rows that cannot be certified are appended to the exact-fallback list.
#include "common.cuh"
#include "kernels.h"

namespace bgnn {

__global__ void __launch_bounds__(256)
normalize_split_kernel(const float* __restrict__ x, long long n, int d, int ld, int normalize,
                       float* __restrict__ hi, float* __restrict__ lo) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const float* p = x + row * d;
  float denom = 1.f;
  if (normalize) {
    float ss = 0.f;
    for (int c = lane; c < d; c += 32) { float v = __ldg(p + c); ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    denom = fmaxf(sqrtf(ss), 1e-8f);
  }
  for (int c = lane; c < ld; c += 32) {
    float v = 0.f;
    if (c < d) { v = __ldg(p + c); if (normalize) v = v / denom; }
    if (lo) {
      float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);   // 10 explicit mantissa bits = tf32
      hi[row * ld + c] = h;
      lo[row * ld + c] = v - h;
    } else {
      hi[row * ld + c] = v;
    }
  }
}

int launch_normalize_split(const float* x, long long n, int d, int ld, int normalize, float* hi, float* lo,
                           cudaStream_t stream) {
  if (n <= 0) return BGNN_OK;
  long long blocks = (n + 7) / 8;
  normalize_split_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, n, d, ld, normalize, hi, lo);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

constexpr int MG_WARPS = 4;
constexpr int MG_MAXC = BGNN_MERGE_MAX_CAND;  // candidates per row held in shared memory

__global__ void __launch_bounds__(MG_WARPS * 32)
knn_merge_kernel(const float* __restrict__ cand_val, const int* __restrict__ cand_idx, int nlists, int kc, int nq, int k,
                 int rescore, const float* __restrict__ qhi, const float* __restrict__ qlo,
                 const float* __restrict__ dhi, const float* __restrict__ dlo, int d, int ld, int apply_sigmoid,
                 float delta, const float* __restrict__ seed_thr, const int* __restrict__ row_list,
                 const int* __restrict__ row_count, int few_rows, long long* __restrict__ out_idx, float* __restrict__ out_val, float* __restrict__ out_gap,
                 int* __restrict__ fb_rows, int* __restrict__ fb_count, float eps, int* __restrict__ out_count) {
  // dynamic shared memory, sized for THIS call's lists (a fixed [MG_MAXC] per warp held the kernel at 5 CTAs per SM):
  // per warp `cap` values, `cap` ids, `cap` 16-bit slots of the nominees that are re-scored; then [MG_WARPS][ld] query
  // rows when rescoring
  extern __shared__ __align__(16) float mg_smem[];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cap = (nlists * kc + 31) / 32 * 32;
  float* sv_all = mg_smem;
  int* si_all = reinterpret_cast<int*>(sv_all + MG_WARPS * cap);
  unsigned short* skl_all = reinterpret_cast<unsigned short*>(si_all + MG_WARPS * cap);
  float* sqrow = reinterpret_cast<float*>(skl_all + MG_WARPS * cap);   // 16-byte aligned: cap is a multiple of 32
  const int nrows = row_list ? min(__ldg(row_count), nq) : nq;
  if (row_list && nrows <= few_rows) return;        // a handful of rows went through knn_exact_rows_kernel
  const long long r = (long long)blockIdx.x * MG_WARPS + wid;
  if (r >= nrows) return;
  const long long row = row_list ? (long long)row_list[r] : r;
  const int total = nlists * kc;
  float* v = sv_all + wid * cap;
  int* ix = si_all + wid * cap;
  // Largest approximate score any list may have discarded: a full list dropped only columns scoring
  // <= its minimum; a list that is not full kept every column of its split.
  // A seeded sweep additionally dropped every column scoring <= the row's seed threshold.
  float worst_thr = seed_thr ? __ldg(seed_thr + row) : -INFINITY;
  for (int l = 0; l < nlists; ++l) {
    const long long base = ((long long)l * nq + row) * kc;
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
    // Nominees that cannot reach the first k + 1 places are not re-scored.  With T = the (k+1)-th largest APPROXIMATE
    // score of the row's nominees, a nominee below T - 2 delta scores, exactly, less than T - delta, which each of the
    // k + 1 nominees at or above T reaches: it is strictly behind k + 1 others whatever the exact values are.  `lo`
    // (bisection, count(v >= lo) >= k + 1 throughout) stands in for T; about half of the nominees go (the lists keep
    // k + 4 each, the union needs k + 1), and the survivors are dealt round-robin to the lanes.
    unsigned short* kl = skl_all + wid * cap;
    if (delta >= 0.f) {
      float vmin = INFINITY, vmax = -INFINITY;
      int nvalid = 0;
      for (int c = lane; c < total; c += 32)
        if (ix[c] >= 0) { vmin = fminf(vmin, v[c]); vmax = fmaxf(vmax, v[c]); ++nvalid; }
      vmin = -warp_max(-vmin);
      vmax = warp_max(vmax);
      nvalid = __reduce_add_sync(0xffffffffu, nvalid);
      if (nvalid > k + 1) {
        float lo = vmin, hi = vmax;
        for (int it = 0; it < 10; ++it) {
          const float mid = 0.5f * (lo + hi);
          int cnt = 0;
          for (int c = lane; c < total; c += 32) cnt += (ix[c] >= 0 && v[c] >= mid) ? 1 : 0;
          cnt = __reduce_add_sync(0xffffffffu, cnt);
          if (cnt >= k + 1) lo = mid; else hi = mid;
        }
        // 1e-6 more: two raw scores that close may round to the same sigmoid, where the index would decide
        const float cut = lo - 2.f * delta - 1e-6f;
        for (int c = lane; c < total; c += 32)
          if (ix[c] >= 0 && v[c] < cut) ix[c] = -1;
      }
    }
    int nkept = 0;
    for (int c0 = 0; c0 < total; c0 += 32) {
      const int c = c0 + lane;
      const bool keep = c < total && ix[c] >= 0;
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) kl[nkept + __popc(m & ((1u << lane) - 1u))] = (unsigned short)c;
      nkept += __popc(m);
    }
    float* qr = sqrow + (size_t)wid * ld;
    for (int c = lane; c < ld; c += 32) qr[c] = qhi[row * ld + c] + (qlo ? qlo[row * ld + c] : 0.f);
    __syncwarp();
    for (int r = lane; r < nkept; r += 32) {
      const int c = kl[r];
      const int j = ix[c];
      const float4* ph = reinterpret_cast<const float4*>(dhi + (long long)j * ld);
      const float4* pl = dlo ? reinterpret_cast<const float4*>(dlo + (long long)j * ld) : nullptr;
      float acc = 0.f;
      for (int h4 = 0; h4 < d / 4; ++h4) {
        float4 x = __ldg(ph + h4);
        if (pl) { float4 y = __ldg(pl + h4); x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }
        const float4 q = *reinterpret_cast<const float4*>(qr + 4 * h4);
        acc = fmaf(q.x, x.x, acc);
        acc = fmaf(q.y, x.y, acc);
        acc = fmaf(q.z, x.z, acc);
        acc = fmaf(q.w, x.w, acc);
      }
      v[c] = apply_sigmoid ? sigmoid_f32(acc) : acc;
    }
    __syncwarp();
  }
  // k+1 rounds of warp arg-best under (value desc, index asc); lists cover disjoint columns, so an
  // index appears at most once.
  float vk = -INFINITY, vk1 = -INFINITY;
  bool have_k1 = false;
  // epsilon threshold of the bridge matching (main_bridged_graph.py:33 `epsilon`), fused into the selection: a
  // neighbour is emitted only if its similarity exceeds eps (NaN = off); values are best first, so the kept
  // neighbours of a row are a prefix whose length goes to out_count
  const bool use_eps = eps == eps;
  int kept = 0;
  for (int round = 0; round <= k; ++round) {
    float bv = -INFINITY;
    int bi = 0x7fffffff, bs = -1;
    for (int c = lane; c < total; c += 32) {
      const int j = ix[c];
      if (j < 0) continue;
      const float cv = v[c];
      if (bs < 0 || cv > bv || (cv == bv && j < bi)) { bv = cv; bi = j; bs = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int os = __shfl_xor_sync(0xffffffffu, bs, o);
      if (os >= 0 && (bs < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; bs = os; }
    }
    if (round < k) {
      const bool emit = bs >= 0 && (!use_eps || bv > eps);
      kept += emit ? 1 : 0;
      if (lane == 0) {
        out_idx[row * k + round] = emit ? (long long)bi : -1LL;
        out_val[row * k + round] = (bs >= 0) ? bv : -INFINITY;
      }
      if (round == k - 1) vk = (bs >= 0) ? bv : -INFINITY;
      if (bs >= 0 && (bs & 31) == lane) ix[bs] = -1;   // consumed (slot c is scanned by lane c % 32)
      __syncwarp();
    } else {
      have_k1 = bs >= 0;
      vk1 = bv;
    }
  }
  if (lane == 0) {
    if (out_count) out_count[row] = kept;
    if (out_gap) out_gap[row] = have_k1 ? (vk - vk1) : INFINITY;
    if (delta >= 0.f && fb_rows && worst_thr > -INFINITY) {
      // every discarded column scores, exactly, at most f(worst_thr + delta); the row is final only if
      // that is strictly below the (k+1)-th kept value, i.e. nothing discarded can enter or tie the
      // first k+1 places.  2.4e-7 covers the last-ulp wobble of expf in sigmoid.
      float bound = worst_thr + delta;
      if (apply_sigmoid) bound = sigmoid_f32(bound) + 2.4e-7f;
      if (!have_k1 || !(bound < vk1)) {
        const int slot = atomicAdd(fb_count, 1);
        fb_rows[slot] = (int)row;
      }
    }
  }
}

int launch_knn_merge(const float* cand_val, const int* cand_idx, int nlists, int kc, int nq, int k, int rescore,
                     const float* qhi, const float* qlo, const float* dhi, const float* dlo, int d, int ld,
                     int apply_sigmoid, float delta, const float* seed_thr, const int* row_list, const int* row_count,
                     int few_rows, long long* out_idx, float* out_val, float* out_gap, int* fb_rows, int* fb_count,
                     float eps, int* out_count, cudaStream_t stream) {
  if (nq <= 0) return BGNN_OK;
  if (nlists * kc > MG_MAXC) return BGNN_ERR_UNSUPPORTED;
  if (rescore && (ld % 4 != 0 || d % 4 != 0)) return BGNN_ERR_INVALID_ARG;
  const size_t cap = (size_t)(nlists * kc + 31) / 32 * 32;
  size_t dyn = MG_WARPS * cap * (sizeof(float) + sizeof(int) + sizeof(unsigned short)) +
               (rescore ? (size_t)MG_WARPS * ld * sizeof(float) : 0);
  if (dyn > 48 * 1024) BGNN_CUDA_TRY(cudaFuncSetAttribute(knn_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  knn_merge_kernel<<<(nq + MG_WARPS - 1) / MG_WARPS, MG_WARPS * 32, dyn, stream>>>(
      cand_val, cand_idx, nlists, kc, nq, k, rescore, qhi, qlo, dhi, dlo, d, ld, apply_sigmoid, delta, seed_thr,
      row_list, row_count, few_rows, out_idx, out_val, out_gap, fb_rows, fb_count, eps, out_count);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

}  // namespace bgnn
