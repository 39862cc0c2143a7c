// Several NARROW AdaptedConv aggregations over the SAME graph in one pass ("heads"), fp32.
//
// KT-GNN ends in three classifier convs (clf_base(x), clf_target(clf_transformer(x)), clf_target(x):
// models/KTGNN.py:432-434) whose outputs have num_classes columns.  Run one by one through gatv2_fused.cu they
// are latency-bound: 8-byte feature rows, and per edge one index load, one sector gather and one scattered
// half-sector record each.  Here HD <= 3 heads of C <= 4 features travel together: H is [n, HD*C] (head h owns
// columns h*C .. h*C+C-1), attention scores, softmax and aggregation stay per head
//     score_h = a_h . leaky_relu(H_h[src] + H_h[dst]),  alpha_h = softmax over the destination's edges,
//     out_h[dst] = sum alpha_h H_h[src]
// but an edge costs ONE index load, ONE gather of HD*C floats and ONE 32-byte record (a full sector:
// (alpha, d score) per head + the leaky-relu branch bits), and every per-row load is shared by the heads.
// ES = 8 lanes share a row and split its EDGES (lane s takes edges s, s+8, ...), so consecutive lanes read
// consecutive index / record entries and a hub row is not one lane's serial loop; the lanes' partial results
// are combined with a fixed butterfly (deterministic).  Semantics per head are exactly those of
// gatv2_fused.cu (PyG softmax incl. +1e-16; (H, a) chosen by the destination's domain).
#include "kernels.h"
#include "rowvec.cuh"

namespace bgnn {

constexpr int MH_ES = 8;                 // lanes per row
constexpr int MH_THREADS = 256;
constexpr int MH_REC_WORDS = 8;          // 32-byte record: (ea, ds) x HD, then the branch-bit word

__device__ __forceinline__ unsigned mh_row_mask(int lane) { return 0xffu << ((lane / MH_ES) * MH_ES); }
__device__ __forceinline__ float mh_lrelu(float t, float slope) { return t > 0.f ? t : t * slope; }

template <int F>
__device__ __forceinline__ void mh_load(const float* __restrict__ p, float (&v)[F]) {
  if (F % 4 == 0) {
#pragma unroll
    for (int k = 0; k < F / 4; ++k) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p) + k);
      v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
    }
  } else if (F % 2 == 0) {
#pragma unroll
    for (int k = 0; k < F / 2; ++k) {
      const float2 t = __ldg(reinterpret_cast<const float2*>(p) + k);
      v[2 * k] = t.x; v[2 * k + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < F; ++k) v[k] = __ldg(p + k);
  }
}

template <int F>
__device__ __forceinline__ void mh_store(float* __restrict__ p, const float (&v)[F]) {
  if (F % 4 == 0) {
#pragma unroll
    for (int k = 0; k < F / 4; ++k) reinterpret_cast<float4*>(p)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  } else if (F % 2 == 0) {
#pragma unroll
    for (int k = 0; k < F / 2; ++k) reinterpret_cast<float2*>(p)[k] = make_float2(v[2 * k], v[2 * k + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < F; ++k) p[k] = v[k];
  }
}

// ------------------------------------------------------------------------------------------ forward
template <int HD, int C>
__global__ void __launch_bounds__(MH_THREADS)
gatv2_heads_fwd_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const uint8_t* __restrict__ dst_is_src,
                       const float* __restrict__ Hs, const float* __restrict__ Ht, const float* __restrict__ af_t2s,
                       const float* __restrict__ af_s2t, float slope, long long n, long long row_off, float* __restrict__ out,
                       float* __restrict__ row_max, float* __restrict__ row_sum) {
  constexpr int F = HD * C;
  const int lane = threadIdx.x & 31, sub = lane % MH_ES;
  const unsigned rmask = mh_row_mask(lane);
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / MH_ES;
  if (row >= n) return;                       // the 8 lanes of a row leave together
  const long long grow = row + row_off;       // global node id (destination-partitioned layout; 0 offset otherwise)
  const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
  const bool is_src = dst_is_src[grow] != 0;
  const float* __restrict__ H = is_src ? Hs : Ht;
  float hi[F], av[F], acc[F], m[HD], l[HD];
  mh_load<F>(H + grow * F, hi);
  mh_load<F>((is_src ? af_t2s : af_s2t), av);
#pragma unroll
  for (int f = 0; f < F; ++f) acc[f] = 0.f;
#pragma unroll
  for (int h = 0; h < HD; ++h) { m[h] = -INFINITY; l[h] = 0.f; }
  int jn = (beg + sub < end) ? __ldg(col + beg + sub) : -1;
  for (int e = beg + sub; e < end; e += MH_ES) {
    const int j = jn;
    float hj[F];
    mh_load<F>(H + (long long)j * F, hj);
    jn = (e + MH_ES < end) ? __ldg(col + e + MH_ES) : -1;
#pragma unroll
    for (int h = 0; h < HD; ++h) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) s = fmaf(av[h * C + c], mh_lrelu(hj[h * C + c] + hi[h * C + c], slope), s);
      const float mn = fmaxf(m[h], s);
      const float sc = expf(m[h] - mn);        // exp(-inf) = 0 on the first edge
      const float p = expf(s - mn);
      l[h] = fmaf(l[h], sc, p);
#pragma unroll
      for (int c = 0; c < C; ++c) acc[h * C + c] = fmaf(acc[h * C + c], sc, p * hj[h * C + c]);
      m[h] = mn;
    }
  }
  // combine the 8 edge shares of the row: online-softmax merge, fixed butterfly
#pragma unroll
  for (int o = 1; o < MH_ES; o <<= 1) {
#pragma unroll
    for (int h = 0; h < HD; ++h) {
      const float om = __shfl_xor_sync(rmask, m[h], o);
      const float ol = __shfl_xor_sync(rmask, l[h], o);
      const float mn = fmaxf(m[h], om);
      const float sa = (m[h] == -INFINITY) ? 0.f : expf(m[h] - mn);
      const float sb = (om == -INFINITY) ? 0.f : expf(om - mn);
      l[h] = l[h] * sa + ol * sb;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float oa = __shfl_xor_sync(rmask, acc[h * C + c], o);
        acc[h * C + c] = acc[h * C + c] * sa + oa * sb;
      }
      m[h] = mn;
    }
  }
  if (sub == 0) {
#pragma unroll
    for (int h = 0; h < HD; ++h) {
      const float inv = 1.0f / (l[h] + 1e-16f);          // PyG softmax denominator
#pragma unroll
      for (int c = 0; c < C; ++c) acc[h * C + c] *= inv;
      if (row_max) row_max[row * HD + h] = m[h];
      if (row_sum) row_sum[row * HD + h] = l[h];
    }
    mh_store<F>(out + row * F, acc);
  }
}

// ------------------------------------------------------------------------------------------ backward, pass A
// per destination row: d H[dst], d a_f partials, and one 32-byte record per edge at its transposed-CSR slot:
//   words 2h, 2h+1 = (alpha_h with the destination's domain in the sign bit, d score_h), word 2*HD = branch bits.
template <int HD, int C>
__global__ void __launch_bounds__(128, 6)
gatv2_heads_bwd_dst_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ csr_to_csc,
                           const uint8_t* __restrict__ dst_is_src, const float* __restrict__ Hs,
                           const float* __restrict__ Ht, const float* __restrict__ af_t2s,
                           const float* __restrict__ af_s2t, float slope, long long n, long long row_off,
                           const float* __restrict__ out,
                           const float* __restrict__ row_max, const float* __restrict__ row_sum,
                           const float* __restrict__ gout, float* __restrict__ gHs, float* __restrict__ gHt,
                           unsigned* __restrict__ erec, float* __restrict__ ga_part) {
  constexpr int F = HD * C;
  constexpr int RPW = 32 / MH_ES;
  __shared__ float s_ga[2 * F][128];           // thread-private running d a_f sums per destination domain
  const int lane = threadIdx.x & 31, sub = lane % MH_ES;
  const unsigned rmask = mh_row_mask(lane);
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
#pragma unroll
  for (int k = 0; k < 2 * F; ++k) s_ga[k][threadIdx.x] = 0.f;
  const long long nsets = (n + RPW - 1) / RPW;
  for (long long set = wid; set < nsets; set += nwarps) {
    const long long row = set * RPW + lane / MH_ES;
    if (row >= n) continue;
    const long long grow = row + row_off;
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    const bool is_src = dst_is_src[grow] != 0;
    float gi[F];
#pragma unroll
    for (int f = 0; f < F; ++f) gi[f] = 0.f;
    if (beg < end) {
      const float* __restrict__ H = is_src ? Hs : Ht;
      float hi[F], av[F], go[F], oi[F], ga[F], D[HD], m[HD], inv[HD], dsum[HD];
      mh_load<F>(H + grow * F, hi);
      mh_load<F>((is_src ? af_t2s : af_s2t), av);
      mh_load<F>(gout + row * F, go);
      mh_load<F>(out + row * F, oi);
#pragma unroll
      for (int h = 0; h < HD; ++h) {
        float d = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) d = fmaf(go[h * C + c], oi[h * C + c], d);
        D[h] = d;
        m[h] = __ldg(row_max + row * HD + h);
        inv[h] = 1.0f / (__ldg(row_sum + row * HD + h) + 1e-16f);
        dsum[h] = 0.f;
      }
#pragma unroll
      for (int f = 0; f < F; ++f) ga[f] = 0.f;
      int jn = (beg + sub < end) ? __ldg(col + beg + sub) : -1;
      int pn = (beg + sub < end) ? __ldg(csr_to_csc + beg + sub) : 0;
      for (int e = beg + sub; e < end; e += MH_ES) {
        const int j = jn, pos = pn;
        float hj[F];
        mh_load<F>(H + (long long)j * F, hj);
        const bool more = e + MH_ES < end;
        jn = more ? __ldg(col + e + MH_ES) : -1;
        pn = more ? __ldg(csr_to_csc + e + MH_ES) : 0;
        unsigned rec[MH_REC_WORDS];
#pragma unroll
        for (int w = 0; w < MH_REC_WORDS; ++w) rec[w] = 0u;
        unsigned bits = 0u;
#pragma unroll
        for (int h = 0; h < HD; ++h) {
          float sp = 0.f, dp = 0.f, lr[C];
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float x = hj[h * C + c];
            const float t = x + hi[h * C + c];
            const bool p = t > 0.f;
            lr[c] = p ? t : t * slope;
            sp = fmaf(av[h * C + c], lr[c], sp);
            dp = fmaf(go[h * C + c], x, dp);
            bits |= p ? (1u << (h * C + c)) : 0u;
          }
          const float alpha = expf(sp - m[h]) * inv[h];
          const float ds = alpha * (dp - D[h]);
          dsum[h] += ds;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            gi[h * C + c] += ((bits >> (h * C + c)) & 1u) ? ds : 0.f;
            ga[h * C + c] = fmaf(ds, lr[c], ga[h * C + c]);
          }
          rec[2 * h] = __float_as_uint(is_src ? -alpha : alpha);     // -0.0f keeps the sign for alpha == 0
          rec[2 * h + 1] = __float_as_uint(ds);
        }
        rec[2 * HD] = bits;
        uint4* dst = reinterpret_cast<uint4*>(erec + (long long)pos * MH_REC_WORDS);
        dst[0] = make_uint4(rec[0], rec[1], rec[2], rec[3]);
        dst[1] = make_uint4(rec[4], rec[5], rec[6], rec[7]);
      }
      // combine the 8 edge shares of the row (fixed butterfly)
#pragma unroll
      for (int o = 1; o < MH_ES; o <<= 1) {
#pragma unroll
        for (int h = 0; h < HD; ++h) dsum[h] += __shfl_xor_sync(rmask, dsum[h], o);
#pragma unroll
        for (int f = 0; f < F; ++f) {
          gi[f] += __shfl_xor_sync(rmask, gi[f], o);
          ga[f] += __shfl_xor_sync(rmask, ga[f], o);
        }
      }
      const float oms = 1.f - slope;
#pragma unroll
      for (int h = 0; h < HD; ++h)
#pragma unroll
        for (int c = 0; c < C; ++c) {
          // d H[dst] = a (.) (slope * sum_j ds_j + (1 - slope) * sum_{t_j > 0} ds_j)
          gi[h * C + c] = av[h * C + c] * fmaf(oms, gi[h * C + c], slope * dsum[h]);
          if (sub == 0) s_ga[(is_src ? 0 : F) + h * C + c][threadIdx.x] += ga[h * C + c];
        }
    }
    float* gdst = is_src ? gHs : gHt;
    if (sub == 0 && gdst) mh_store<F>(gdst + grow * F, gi);
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < F; ++k) {
    float vs = s_ga[k][threadIdx.x], vt = s_ga[F + k][threadIdx.x];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      vs += __shfl_xor_sync(0xffffffffu, vs, o);
      vt += __shfl_xor_sync(0xffffffffu, vt, o);
    }
    if (lane == 0) {
      ga_part[wid * 2 * F + k] = vs;
      ga_part[wid * 2 * F + F + k] = vt;
    }
  }
}

__global__ void __launch_bounds__(256)
mh_reduce_partials_kernel(const float* __restrict__ part, long long nparts, int width, float* __restrict__ o0,
                          float* __restrict__ o1, int f) {
  __shared__ float red[256];
  const int t = blockIdx.x;
  float acc = 0.f;
  for (long long p = threadIdx.x; p < nparts; p += 256) acc += part[p * width + t];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) { if (t < f) o0[t] = red[0]; else o1[t - f] = red[0]; }
}

// ------------------------------------------------------------------------------------------ backward, pass B
// per source row over its outgoing edges (transposed CSR): streams the records, gathers gout[dst]:
//   dH[j] += alpha_ij gout_i + ds_ij a (.) lrelu'(H_j + H_i)   into gHs / gHt by the destination's domain.
template <int HD, int C>
__global__ void __launch_bounds__(MH_THREADS)
gatv2_heads_bwd_src_kernel(const int* __restrict__ t_rowptr, const int* __restrict__ t_col,
                           const uint8_t* __restrict__ dst_is_src, const float* __restrict__ af_t2s,
                           const float* __restrict__ af_s2t, float slope, long long n, long long own_lo, long long own_n,
                           const unsigned* __restrict__ erec,
                           const float* __restrict__ gout, float* __restrict__ gHs, float* __restrict__ gHt) {
  constexpr int F = HD * C;
  const int lane = threadIdx.x & 31, sub = lane % MH_ES;
  const unsigned rmask = mh_row_mask(lane);
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / MH_ES;
  if (row >= n) return;
  float as_[F], at_[F], gs[F], gt[F];
  mh_load<F>(af_t2s, as_);
  mh_load<F>(af_s2t, at_);
#pragma unroll
  for (int f = 0; f < F; ++f) { gs[f] = 0.f; gt[f] = 0.f; }
  const int beg = __ldg(t_rowptr + row), end = __ldg(t_rowptr + row + 1);
  int in_ = (beg + sub < end) ? __ldg(t_col + beg + sub) : -1;
  for (int e = beg + sub; e < end; e += MH_ES) {
    const int i = in_;
    float go[F];
    mh_load<F>(gout + (long long)i * F, go);
    in_ = (e + MH_ES < end) ? __ldg(t_col + e + MH_ES) : -1;
    const uint4* src = reinterpret_cast<const uint4*>(erec + (long long)e * MH_REC_WORDS);
    const uint4 r0 = __ldg(src), r1 = __ldg(src + 1);
    const unsigned rec[MH_REC_WORDS] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    const unsigned bits = rec[2 * HD];
#pragma unroll
    for (int h = 0; h < HD; ++h) {
      const float al = __uint_as_float(rec[2 * h]), ds = __uint_as_float(rec[2 * h + 1]);
      const bool dsrc = signbit(al);
      const float alpha = fabsf(al);
      const float dss = dsrc ? ds : 0.f, dst_ = dsrc ? 0.f : ds, als = dsrc ? alpha : 0.f, alt = dsrc ? 0.f : alpha;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int f = h * C + c;
        const float dl = ((bits >> f) & 1u) ? 1.f : slope;
        gs[f] = fmaf(dss * as_[f], dl, fmaf(als, go[f], gs[f]));
        gt[f] = fmaf(dst_ * at_[f], dl, fmaf(alt, go[f], gt[f]));
      }
    }
  }
#pragma unroll
  for (int o = 1; o < MH_ES; o <<= 1) {
#pragma unroll
    for (int f = 0; f < F; ++f) {
      gs[f] += __shfl_xor_sync(rmask, gs[f], o);
      gt[f] += __shfl_xor_sync(rmask, gt[f], o);
    }
  }
  if (sub == 0) {
    // destination-side part of this row from pass A lives in the array of the row's own domain
    const bool me_src = dst_is_src[row] != 0;
    const float* ownp = me_src ? gHs : gHt;
    if (ownp && row >= own_lo && row < own_lo + own_n) {     // rows this rank owns as destinations
      float own[F];
      mh_load<F>(ownp + row * F, own);
#pragma unroll
      for (int f = 0; f < F; ++f) {
        gs[f] += me_src ? own[f] : 0.f;
        gt[f] += me_src ? 0.f : own[f];
      }
    }
    if (gHs) mh_store<F>(gHs + row * F, gs);
    if (gHt) mh_store<F>(gHt + row * F, gt);
  }
}

// ------------------------------------------------------------------------------------------ host
bool gatv2_heads_supported(int heads, int c) { return (heads == 2 || heads == 3) && c >= 1 && c <= 4; }

constexpr int kMhDstThreads = 128;
constexpr int kMhDstMaxCtasPerSm = 16;
// Pass A takes one row SET (4 rows) per warp, CTAs scheduled by the hardware: with a static round-robin over resident
// warps, the warp that drew a hub row (1800 edges = 220 serial steps of its 8 lanes) kept its other 70 sets waiting
// behind it and set the duration of the launch (profiles/r02n_gat_ncu.txt: 1.0 ms at 18 % DRAM utilisation).
static long long mh_dst_warps(long long n) {
  const long long nsets = (n + (32 / MH_ES) - 1) / (32 / MH_ES);
  const long long ctas = (nsets + (kMhDstThreads / 32) - 1) / (kMhDstThreads / 32);
  return (ctas < 1 ? 1 : ctas) * (kMhDstThreads / 32);
}

#define MH_DISPATCH(CALL)                                                                  \
  do {                                                                                     \
    if (heads == 2) {                                                                      \
      switch (c) { case 1: CALL(2, 1); break; case 2: CALL(2, 2); break; case 3: CALL(2, 3); break; default: CALL(2, 4); break; } \
    } else {                                                                               \
      switch (c) { case 1: CALL(3, 1); break; case 2: CALL(3, 2); break; case 3: CALL(3, 3); break; default: CALL(3, 4); break; } \
    }                                                                                      \
  } while (0)

int launch_gatv2_heads_fwd(const int* rowptr, const int* col, const uint8_t* dst_is_src, const float* Hs, const float* Ht,
                           const float* af_t2s, const float* af_s2t, float slope, long long n, long long row_off, int heads, int c,
                           float* out, float* row_max, float* row_sum, cudaStream_t stream) {
  if (n <= 0) return BGNN_OK;
  if (!gatv2_heads_supported(heads, c)) return BGNN_ERR_UNSUPPORTED;
  const long long blocks = (n * MH_ES + MH_THREADS - 1) / MH_THREADS;
#define CALL(H_, C_)                                                                                             \
  gatv2_heads_fwd_kernel<H_, C_><<<(unsigned)blocks, MH_THREADS, 0, stream>>>(rowptr, col, dst_is_src, Hs, Ht, af_t2s, \
                                                                              af_s2t, slope, n, row_off, out, row_max, row_sum)
  MH_DISPATCH(CALL);
#undef CALL
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

size_t gatv2_heads_bwd_workspace_bytes(long long n, long long e, int heads, int c) {
  if (!gatv2_heads_supported(heads, c)) return 0;
  return align_up((size_t)mh_dst_warps(n) * 2 * heads * c * sizeof(float), 256) +
         align_up((size_t)e * MH_REC_WORDS * sizeof(unsigned), 256) + 1024;
}

template <int HD, int C>
static int mh_launch_dst(long long n, long long row_off, int& nwarps, cudaStream_t stream, const int* rowptr, const int* col,
                         const int* csr_to_csc, const uint8_t* dst_is_src, const float* Hs, const float* Ht,
                         const float* af_t2s, const float* af_s2t, float slope, const float* out, const float* row_max,
                         const float* row_sum, const float* gout, float* gHs, float* gHt, unsigned* erec, float* part) {
  if (n <= 0) { nwarps = 0; return BGNN_OK; }        // a rank that owns no destination row
  auto kern = gatv2_heads_bwd_dst_kernel<HD, C>;
  static int occ = 0;
  if (occ == 0) {
    int o = 0;
    BGNN_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kMhDstThreads, 0));
    occ = o < 1 ? 1 : (o > kMhDstMaxCtasPerSm ? kMhDstMaxCtasPerSm : o);
  }
  constexpr int WPC = kMhDstThreads / 32;
  const long long nsets = (n + (32 / MH_ES) - 1) / (32 / MH_ES);
  long long ctas = (nsets + WPC - 1) / WPC;
  (void)occ;
  nwarps = (int)(ctas * WPC);
  kern<<<(unsigned)ctas, kMhDstThreads, 0, stream>>>(rowptr, col, csr_to_csc, dst_is_src, Hs, Ht, af_t2s, af_s2t, slope, n,
                                                     row_off, out, row_max, row_sum, gout, gHs, gHt, erec, part);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

int launch_gatv2_heads_bwd(const int* rowptr, const int* col, const int* t_rowptr, const int* t_col, const int* csr_to_csc,
                           long long e, const uint8_t* dst_is_src, const float* Hs, const float* Ht, const float* af_t2s,
                           const float* af_s2t, float slope, long long n, long long row_off, long long n_src, int heads, int c,
                           const float* out,
                           const float* row_max, const float* row_sum, const float* gout, float* gHs, float* gHt,
                           float* g_af_t2s, float* g_af_s2t, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (n_src <= 0) return BGNN_OK;
  if (!gatv2_heads_supported(heads, c)) return BGNN_ERR_UNSUPPORTED;
  const int f = heads * c;
  Workspace w(ws, ws_bytes);
  float* part = w.take<float>(mh_dst_warps(n_src) * 2 * f);
  unsigned* erec = w.take<unsigned>((size_t)e * MH_REC_WORDS);
  if (!w.ok()) return BGNN_ERR_WORKSPACE;
  int nparts = 0, rc = BGNN_OK;
#define CALL(H_, C_)                                                                                                \
  rc = mh_launch_dst<H_, C_>(n, row_off, nparts, stream, rowptr, col, csr_to_csc, dst_is_src, Hs, Ht, af_t2s, af_s2t, slope, out, \
                             row_max, row_sum, gout, gHs, gHt, erec, part)
  MH_DISPATCH(CALL);
#undef CALL
  if (rc != BGNN_OK) return rc;
  mh_reduce_partials_kernel<<<2 * f, 256, 0, stream>>>(part, nparts, 2 * f, g_af_t2s, g_af_s2t, f);
  BGNN_LAUNCH_CHECK();
  const long long blocks = (n_src * MH_ES + MH_THREADS - 1) / MH_THREADS;
#define CALL(H_, C_)                                                                                                \
  gatv2_heads_bwd_src_kernel<H_, C_><<<(unsigned)blocks, MH_THREADS, 0, stream>>>(t_rowptr, t_col, dst_is_src, af_t2s, \
                                                                                  af_s2t, slope, n_src, row_off, n, erec, gout, gHs, gHt)
  MH_DISPATCH(CALL);
#undef CALL
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

}  // namespace bgnn
