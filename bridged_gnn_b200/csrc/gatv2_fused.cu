// Fused KT-GNN AdaptedConv aggregation (GATv2-style attention over destination rows), fp32, HBM-bound.
//
// Reference op sequence being replaced (models/KTGNN.py:292-305 + message :317-319, and PyG
// utils.softmax / MessagePassing.propagate underneath):
//     a1 = a_f_t2s(leaky_relu(Hs[src]+Hs[dst]))   for edges whose dst is a source-domain node
//     a2 = a_f_s2t(leaky_relu(Ht[src]+Ht[dst]))   for edges whose dst is a target-domain node
//     alpha = softmax over incoming edges of dst ( exp(s-max) / (sum + 1e-16) )
//     out[dst] = sum alpha * H[src]
// The two edge sets have disjoint destinations, so per destination row i one picks (H, a) by the
// domain of i and does a single online-softmax pass over the row's incoming edges: one gather of
// H[src] per edge, no [E,C] intermediates, no atomics.
//
// Backward (K3b) recomputes the scores from the saved per-row (max, sum):
//   pass A (CSR by dst):  D_i = gout_i . out_i ;  dH[dst] part, d a_f partials (per-CTA, reduced after)
//   pass B (CSC by src):  dH[src] part = sum alpha*gout_i + ds * a (.) lrelu'(H_j+H_i)
// Both passes are atomic-free and deterministic.
#include <stdlib.h>

#include "kernels.h"
#include "rowvec.cuh"

namespace bgnn {

template <int G>
__device__ __forceinline__ unsigned group_mask(int lane) {
  return G == 32 ? 0xffffffffu : (((1u << G) - 1u) << ((lane / G) * G));
}

template <int G>
__device__ __forceinline__ float gsum(float v, unsigned mask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

__device__ __forceinline__ float lrelu(float t, float slope) { return t > 0.f ? t : t * slope; }

// ------------------------------------------------------------------------------------------ forward
template <int VEC, int G, int CH>
__global__ void __launch_bounds__(256)
gatv2_fwd_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ order,
                 const uint8_t* __restrict__ dst_is_src,
                 const float* __restrict__ Hs, const float* __restrict__ Ht, const float* __restrict__ af_t2s,
                 const float* __restrict__ af_s2t, float slope, long long n, int c, float* __restrict__ out,
                 float* __restrict__ row_max, float* __restrict__ row_sum) {
  const int lane = threadIdx.x & 31;
  const int lane_g = threadIdx.x % G;
  const unsigned mask = group_mask<G>(lane);
  // `order` (optional): rows by descending degree -- hub rows start first instead of forming the tail of the
  // launch, and the rows that share a warp have similar lengths
  const long long slot = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (slot >= n) return;  // a group leaves together; shuffles below use the group's own mask
  const long long row = order ? (long long)__ldg(order + slot) : slot;
  const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
  if (beg == end) {
    // no incoming edge (rows owned by another rank in the destination-partitioned multi-GPU layout):
    // the aggregate is 0; skip every feature load
    Chunk<VEC> z;
#pragma unroll
    for (int i = 0; i < VEC; ++i) z.v[i] = 0.f;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      const int c0 = (ch * G + lane_g) * VEC;
      st_chunk<VEC>(out + row * c + c0, z, c0 < c);
    }
    if (lane_g == 0) {
      if (row_max) row_max[row] = -INFINITY;
      if (row_sum) row_sum[row] = 0.f;
    }
    return;
  }
  const bool is_src = dst_is_src[row] != 0;
  const float* __restrict__ H = is_src ? Hs : Ht;
  const float* __restrict__ a = is_src ? af_t2s : af_s2t;
  Chunk<VEC> hi[CH], av[CH], acc[CH];
  bool cok[CH];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) {
    int c0 = (ch * G + lane_g) * VEC;
    cok[ch] = c0 < c;
    hi[ch] = ld_chunk<VEC>(H + row * c + c0, cok[ch]);
    av[ch] = ld_chunk<VEC>(a + c0, cok[ch]);
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[ch].v[i] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  constexpr int U = 4;
  for (int e = beg; e < end; e += U) {
    int j[U];
#pragma unroll
    for (int u = 0; u < U; ++u) j[u] = (e + u < end) ? __ldg(col + e + u) : -1;
    Chunk<VEC> hj[U][CH];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int ch = 0; ch < CH; ++ch)
        hj[u][ch] = ld_chunk<VEC>(H + (long long)(j[u] < 0 ? 0 : j[u]) * c + (ch * G + lane_g) * VEC, cok[ch] && j[u] >= 0);
    float s[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float p = 0.f;
#pragma unroll
      for (int ch = 0; ch < CH; ++ch)
#pragma unroll
        for (int i = 0; i < VEC; ++i) p = fmaf(av[ch].v[i], lrelu(hj[u][ch].v[i] + hi[ch].v[i], slope), p);
      s[u] = p;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      s[u] = gsum<G>(s[u], mask);
      if (j[u] < 0) s[u] = -INFINITY;
    }
    float mb = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3]));
    float mn = fmaxf(m, mb);          // finite: the batch has at least one valid edge
    float sc = expf(m - mn);          // exp(-inf) = 0 on the first batch
    l *= sc;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch)
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[ch].v[i] *= sc;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float p = expf(s[u] - mn);      // 0 for padded slots
      l += p;
#pragma unroll
      for (int ch = 0; ch < CH; ++ch)
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[ch].v[i] = fmaf(p, hj[u][ch].v[i], acc[ch].v[i]);
    }
    m = mn;
  }
  const float inv = 1.0f / (l + 1e-16f);   // PyG softmax denominator
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[ch].v[i] *= inv;
    st_chunk<VEC>(out + row * c + (ch * G + lane_g) * VEC, acc[ch], cok[ch]);
  }
  if (lane_g == 0) {
    if (row_max) row_max[row] = m;
    if (row_sum) row_sum[row] = l;
  }
}

int launch_gatv2_fwd(const int* rowptr, const int* col, const int* order, const uint8_t* dst_is_src, const float* Hs, const float* Ht,
                     const float* af_t2s, const float* af_s2t, float slope, long long n, int c, float* out,
                     float* row_max, float* row_sum, cudaStream_t stream) {
  if (n <= 0) return BGNN_OK;
  int vec, g, ch;
  if (!pick_row_config(c, vec, g, ch)) return BGNN_ERR_UNSUPPORTED;
  long long blocks = (n * g + 255) / 256;
#define CALL(V, G_, C_)                                                                                        \
  gatv2_fwd_kernel<V, G_, C_><<<(unsigned)blocks, 256, 0, stream>>>(rowptr, col, order, dst_is_src, Hs, Ht, af_t2s,   \
                                                                      af_s2t, slope, n, c, out, row_max, row_sum)
  BGNN_ROW_DISPATCH(vec, g, ch, CALL);
#undef CALL
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

// ------------------------------------------------------------------------------------------ backward
// Pass A: per destination row (CSR).  Recomputes the scores from the saved (max, sum), and writes
//   * the destination-side gradient of row i into gHs (source-domain row) or gHt (target-domain row),
//   * per-WARP partial sums of d a_f into ga_part[warp][2][c] (reduced by reduce_partials_kernel),
//   * one record per edge, stored at the edge's slot in the TRANSPOSED CSR so that pass B streams them:
//       ea   = alpha_ij with the destination's domain in the sign bit (negative: source-domain destination)
//       eds  = d score_ij = alpha_ij (gout_i . H_j - gout_i . out_i)
//       emask[CW] = bit c set iff H_j[c] + H_i[c] > 0 (the leaky-relu branch)
// so that pass B needs neither H[dst] nor the softmax statistics again.
//
// The pass is bound by instruction issue and latency, not by memory (profiles/r01e, r01h), so its row mapping
// differs from the forward kernel's: a group of G lanes owns a row and every lane owns EPL = VEC*CH CONTIGUOUS
// features (8, or 16 for c > 256).  Fewer lanes per row means fewer shuffle steps and less per-edge bookkeeping
// per feature, and a lane's leaky-relu bits are exactly byte(s) lane_g*EPL/8.. of the edge's mask, which it
// stores itself -- no cross-lane combination of mask words.  U edges are in flight per group; warps are
// persistent over a static round-robin of row sets (no CTA barrier: a long row delays only its own warp, and
// the d a_f summation order stays fixed).
template <int VEC, int G, int CH, int U, int MINB, int ES>
__global__ void __launch_bounds__(128, MINB)
gatv2_bwd_dst_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ csr_to_csc,
                     const int* __restrict__ order, const uint8_t* __restrict__ dst_is_src, const float* __restrict__ Hs, const float* __restrict__ Ht,
                     const float* __restrict__ af_t2s, const float* __restrict__ af_s2t, float slope, long long n, int c,
                     int cw, const float* __restrict__ out, const float* __restrict__ row_max,
                     const float* __restrict__ row_sum, const float* __restrict__ gout, float* __restrict__ gHs,
                     float* __restrict__ gHt, unsigned* __restrict__ erec, unsigned* __restrict__ emask,
                     float* __restrict__ ga_part) {
  // ES > 1 (narrow rows, G == 1): ES lanes share a row and split its EDGES (lane s takes edges s, s + ES, ...):
  // consecutive lanes read consecutive col entries, and a hub row is no longer one lane's serial loop.
  static_assert(ES == 1 || G == 1, "edge splitting is for rows a single lane can hold");
  constexpr int RL = G * ES;                       // lanes per row
  constexpr int RPW = 32 / RL;                     // rows per warp
  constexpr int EPL = VEC * CH;                    // contiguous features per lane
  static_assert(G == 1 || EPL == 8 || EPL == 16, "multi-lane rows own whole mask bytes");
  const int lane = threadIdx.x & 31;
  const int lane_g = lane % G;
  const int sub = (lane / G) % ES;                 // which share of the row's edges
  const unsigned gmask = group_mask<G>(lane);
  const unsigned rmask = group_mask<RL>(lane);
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int col0 = lane_g * EPL;
  bool cok[CH];
  // running d a_f sums of this lane, per destination domain: thread-private columns of shared memory
  // (kept out of the register file, which bounds this kernel's occupancy)
  __shared__ float s_ga[2 * EPL][128];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) cok[ch] = col0 + ch * VEC < c;
#pragma unroll
  for (int k = 0; k < 2 * EPL; ++k) s_ga[k][threadIdx.x] = 0.f;
  // where this lane's mask bits go: byte offset inside the record (c <= 64) or inside the mask row
  unsigned char* const mask_base = (cw <= 2) ? reinterpret_cast<unsigned char*>(erec) + 8 : reinterpret_cast<unsigned char*>(emask);
  const long long mask_stride = (cw <= 2) ? 16 : (long long)cw * 4;
  const int mask_off = col0 >> 3;
  const bool mask_ok = col0 < c;
  const long long nsets = (n + RPW - 1) / RPW;
  for (long long set = wid; set < nsets; set += nwarps) {
    const long long slot = set * RPW + lane / RL;
    if (slot >= n) continue;                       // whole row groups leave together
    const long long row = order ? (long long)__ldg(order + slot) : slot;
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    const bool is_src = dst_is_src[row] != 0;
    Chunk<VEC> gi[CH];
#pragma unroll
    for (int ch = 0; ch < CH; ++ch)
#pragma unroll
      for (int i = 0; i < VEC; ++i) gi[ch].v[i] = 0.f;
    if (beg < end) {                               // no incoming edge: the destination-side gradient is 0
      const float* __restrict__ H = is_src ? Hs : Ht;
      const float* __restrict__ a = is_src ? af_t2s : af_s2t;
      Chunk<VEC> hi[CH], av[CH], go[CH], ga[CH];
      float dpart = 0.f;
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) {
        hi[ch] = ld_chunk<VEC>(H + row * c + col0 + ch * VEC, cok[ch]);
        av[ch] = ld_chunk<VEC>(a + col0 + ch * VEC, cok[ch]);
        go[ch] = ld_chunk<VEC>(gout + row * c + col0 + ch * VEC, cok[ch]);
        const Chunk<VEC> oi = ld_chunk<VEC>(out + row * c + col0 + ch * VEC, cok[ch]);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          dpart = fmaf(go[ch].v[i], oi.v[i], dpart);
          ga[ch].v[i] = 0.f;
        }
      }
      const float Di = gsum<G>(dpart, gmask);
      const float m = row_max[row];
      const float inv = 1.0f / (row_sum[row] + 1e-16f);
      float dsum = 0.f;                            // sum of d score over the row's edges
      int nj[U], np[U];                            // (source, transposed slot) of the NEXT batch: its gathers start at once
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int ee = beg + u * ES + sub;
        const bool ok = ee < end;
        nj[u] = ok ? __ldg(col + ee) : -1;
        np[u] = ok ? __ldg(csr_to_csc + ee) : 0;
      }
      for (int e = beg; e < end; e += U * ES) {
        int j[U], pos[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { j[u] = nj[u]; pos[u] = np[u]; }
        Chunk<VEC> hj[U][CH];                      // H[src]; overwritten by lrelu(H[src] + H[dst]) below
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int ch = 0; ch < CH; ++ch)
            hj[u][ch] = ld_chunk<VEC>(H + (long long)(j[u] < 0 ? 0 : j[u]) * c + col0 + ch * VEC, cok[ch] && j[u] >= 0);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int ee = e + (U + u) * ES + sub;
          const bool ok = ee < end;
          nj[u] = ok ? __ldg(col + ee) : -1;
          np[u] = ok ? __ldg(csr_to_csc + ee) : 0;
        }
        float sp[U], dp[U];
        unsigned bits[U];                          // leaky-relu branch bits of this lane's features (bit = feature)
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float s = 0.f, d = 0.f;
          unsigned b = 0u;
#pragma unroll
          for (int ch = 0; ch < CH; ++ch)
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
              const float x = hj[u][ch].v[i];
              const float t = x + hi[ch].v[i];
              const bool p = t > 0.f;
              const float l = p ? t : t * slope;
              s = fmaf(av[ch].v[i], l, s);
              d = fmaf(go[ch].v[i], x, d);
              b |= p ? (1u << (ch * VEC + i)) : 0u;
              hj[u][ch].v[i] = l;
            }
          sp[u] = s;
          dp[u] = d;
          bits[u] = b;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          sp[u] = gsum<G>(sp[u], gmask);
          dp[u] = gsum<G>(dp[u], gmask);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float alpha = (j[u] >= 0) ? expf(sp[u] - m) * inv : 0.f;
          const float ds = alpha * (dp[u] - Di);
          dsum += ds;
          // d H[dst] = a (.) sum_j ds_j lrelu'(t_j) = a (.) (slope * sum_j ds_j + (1 - slope) * sum_{t_j > 0} ds_j):
          // only the second sum is per feature
#pragma unroll
          for (int ch = 0; ch < CH; ++ch)
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
              gi[ch].v[i] += ((bits[u] >> (ch * VEC + i)) & 1u) ? ds : 0.f;
              ga[ch].v[i] = fmaf(ds, hj[u][ch].v[i], ga[ch].v[i]);
            }
          if (j[u] >= 0) {
            const float ea = is_src ? -alpha : alpha;      // -0.0f keeps the sign for alpha == 0
            if (G == 1 && cw <= 2) {
              // a single lane owns the row: the whole 16-byte record in one store
              *reinterpret_cast<uint4*>(erec + (long long)pos[u] * 4) =
                  make_uint4(__float_as_uint(ea), __float_as_uint(ds), bits[u], 0u);
              continue;
            }
            if (lane_g == 0)
              *reinterpret_cast<float2*>(erec + (long long)pos[u] * (cw <= 2 ? 4 : 2)) = make_float2(ea, ds);
            if (mask_ok) {
              unsigned char* mp = mask_base + (long long)pos[u] * mask_stride + mask_off;
              if (EPL <= 8) *mp = (unsigned char)bits[u];
              else *reinterpret_cast<unsigned short*>(mp) = (unsigned short)bits[u];
            }
          }
        }
      }
      if (ES > 1) {                                // combine the edge shares of the row (fixed butterfly)
#pragma unroll
        for (int o = G; o < RL; o <<= 1) {
          dsum += __shfl_xor_sync(rmask, dsum, o);
#pragma unroll
          for (int ch = 0; ch < CH; ++ch)
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
              gi[ch].v[i] += __shfl_xor_sync(rmask, gi[ch].v[i], o);
              ga[ch].v[i] += __shfl_xor_sync(rmask, ga[ch].v[i], o);
            }
        }
      }
      const float oms = 1.f - slope, sds = slope * dsum;
#pragma unroll
      for (int ch = 0; ch < CH; ++ch)
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          gi[ch].v[i] = av[ch].v[i] * fmaf(oms, gi[ch].v[i], sds);
          if (sub == 0) s_ga[(is_src ? 0 : EPL) + ch * VEC + i][threadIdx.x] += ga[ch].v[i];
        }
    }
    if (sub == 0) {
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) st_chunk<VEC>((is_src ? gHs : gHt) + row * c + col0 + ch * VEC, gi[ch], cok[ch]);
    }
  }
  __syncwarp();
  // sum this warp's groups (fixed butterfly), then the first group writes the warp's partial
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
    float vs = s_ga[k][threadIdx.x], vt = s_ga[EPL + k][threadIdx.x];
#pragma unroll
    for (int o = G; o < 32; o <<= 1) {
      vs += __shfl_xor_sync(0xffffffffu, vs, o);
      vt += __shfl_xor_sync(0xffffffffu, vt, o);
    }
    const int cc = col0 + k;
    if (lane < G && cc < c) {
      ga_part[wid * 2 * c + cc] = vs;
      ga_part[wid * 2 * c + c + cc] = vt;
    }
  }
}

// Row mapping of pass A: 128-bit loads when c % 4 == 0 (64-bit for other even widths, scalar otherwise), 8
// contiguous features per lane (16 for c > 256; a single lane for c <= 8), the smallest power-of-two group
// that covers the row.
static bool pick_dst_config(int c, int& vec, int& g, int& ch) {
  if (c <= 0 || c > 512) return false;
  vec = (c % 4 == 0) ? 4 : (c % 2 == 0) ? 2 : 1;
  if (c <= 8) {
    g = 1;
    ch = 1;
    while (ch * vec < c) ch <<= 1;                // 1, 2, 4, 8 chunks
    return true;
  }
  const int epl = c > 256 ? 16 : 8;
  if (epl / vec > 8) return false;                // odd / 2-mod-4 widths above 256 are not instantiated
  g = 2;
  while (g < 32 && g * epl < c) g <<= 1;
  ch = epl / vec;
  return true;
}
constexpr int kBwdDstThreads = 128;
constexpr int kBwdDstMaxCtasPerSm = 16;
static long long bwd_dst_max_warps() { return (long long)kNumSMs * kBwdDstMaxCtasPerSm * (kBwdDstThreads / 32); }

// Column-wise reduction of the per-CTA partials: one CTA per column, strided partial sums then a
// fixed-shape tree -> deterministic.
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ part, long long nparts, int width, float* __restrict__ o0,
                       float* __restrict__ o1, int c) {
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
  if (threadIdx.x == 0) { if (t < c) o0[t] = red[0]; else o1[t - c] = red[0]; }
}

// Pass B: per source row j over its outgoing edges (transposed CSR).  Streams the edge records of
// pass A, gathers only gout[dst], and adds the source-side gradient
//   dH[j] += alpha_ij gout_i + ds_ij a (.) lrelu'(H_j + H_i)
// into gHs (edges into source-domain destinations) / gHt (target-domain destinations).
template <int VEC, int G, int CH>
__global__ void __launch_bounds__(256, (VEC * CH <= 4) ? 4 : 1)
gatv2_bwd_src_kernel(const int* __restrict__ t_rowptr, const int* __restrict__ t_col, const int* __restrict__ t_order,
                     const uint8_t* __restrict__ dst_is_src, const float* __restrict__ af_t2s,
                     const float* __restrict__ af_s2t, float slope, long long n, int c, int cw,
                     const unsigned* __restrict__ erec, const unsigned* __restrict__ emask,
                     const float* __restrict__ gout, float* __restrict__ gHs, float* __restrict__ gHt) {
  const int lane_g = threadIdx.x % G;
  const long long slot = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (slot >= n) return;
  const long long row = t_order ? (long long)__ldg(t_order + slot) : slot;
  // per feature: the attention vector of either destination domain and its slope-scaled copy, so that the
  // leaky-relu branch of an edge is one select (a or slope * a) per domain
  Chunk<VEC> as_[CH], asl[CH], at_[CH], atl[CH], gs[CH], gt[CH];
  bool cok[CH];
  int wsel[CH], bsel[CH];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) {
    int c0 = (ch * G + lane_g) * VEC;
    cok[ch] = c0 < c;
    wsel[ch] = min(c0 >> 5, cw - 1);
    bsel[ch] = c0 & 31;
    as_[ch] = ld_chunk<VEC>(af_t2s + c0, cok[ch]);
    at_[ch] = ld_chunk<VEC>(af_s2t + c0, cok[ch]);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      asl[ch].v[i] = as_[ch].v[i] * slope;
      atl[ch].v[i] = at_[ch].v[i] * slope;
      gs[ch].v[i] = 0.f;
      gt[ch].v[i] = 0.f;
    }
  }
  const int beg = __ldg(t_rowptr + row), end = __ldg(t_rowptr + row + 1);
  constexpr int U = 4;
  int nxt[U];                                        // destination ids of the NEXT batch: their gathers start at once
#pragma unroll
  for (int u = 0; u < U; ++u) nxt[u] = (beg + u < end) ? __ldg(t_col + beg + u) : -1;
  for (int e = beg; e < end; e += U) {
    int i_dst[U];
    Chunk<VEC> go[U][CH];
#pragma unroll
    for (int u = 0; u < U; ++u) i_dst[u] = nxt[u];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int ch = 0; ch < CH; ++ch)
        go[u][ch] = ld_chunk<VEC>(gout + (long long)(i_dst[u] < 0 ? 0 : i_dst[u]) * c + (ch * G + lane_g) * VEC,
                                  cok[ch] && i_dst[u] >= 0);
#pragma unroll
    for (int u = 0; u < U; ++u) nxt[u] = (e + U + u < end) ? __ldg(t_col + e + U + u) : -1;
    float dss[U], dst_[U], als[U], alt[U];           // (d score, alpha) routed to the destination's domain
    unsigned mk[U][CH];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = e + u < end;
      float al, ds;
      if (cw <= 2) {
        const uint4 r4 = ok ? __ldg(reinterpret_cast<const uint4*>(erec) + e + u) : make_uint4(0u, 0u, 0u, 0u);
        al = __uint_as_float(r4.x);
        ds = __uint_as_float(r4.y);
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) mk[u][ch] = wsel[ch] ? r4.w : r4.z;
      } else {
        const float2 r2 = ok ? __ldg(reinterpret_cast<const float2*>(erec) + e + u) : make_float2(0.f, 0.f);
        al = r2.x;
        ds = r2.y;
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) mk[u][ch] = ok ? __ldg(emask + (long long)(e + u) * cw + wsel[ch]) : 0u;
      }
      const bool dsrc = signbit(al);
      const float alpha = fabsf(al);
      dss[u] = dsrc ? ds : 0.f;
      dst_[u] = dsrc ? 0.f : ds;
      als[u] = dsrc ? alpha : 0.f;
      alt[u] = dsrc ? 0.f : alpha;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int ch = 0; ch < CH; ++ch)
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const bool pos = (mk[u][ch] >> (bsel[ch] + i)) & 1u;
          const float g0 = go[u][ch].v[i];
          gs[ch].v[i] = fmaf(dss[u], pos ? as_[ch].v[i] : asl[ch].v[i], fmaf(als[u], g0, gs[ch].v[i]));
          gt[ch].v[i] = fmaf(dst_[u], pos ? at_[ch].v[i] : atl[ch].v[i], fmaf(alt[u], g0, gt[ch].v[i]));
        }
  }
  const bool me_src = dst_is_src[row] != 0;
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) {
    int c0 = (ch * G + lane_g) * VEC;
    if (!cok[ch]) continue;
    // destination-side part of this row from pass A lives in the array of the row's own domain
    Chunk<VEC> own = ld_chunk<VEC>((me_src ? gHs : gHt) + row * c + c0, true);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      gs[ch].v[i] += me_src ? own.v[i] : 0.f;
      gt[ch].v[i] += me_src ? 0.f : own.v[i];
    }
    st_chunk<VEC>(gHs + row * c + c0, gs[ch], true);
    st_chunk<VEC>(gHt + row * c + c0, gt[ch], true);
  }
}

static long long bwd_blocks(long long n, int g) { return (n * g + 255) / 256; }

size_t gatv2_bwd_workspace_bytes(long long n, long long e, int c) {
  int vec, g, ch;
  if (!pick_row_config(c, vec, g, ch) || !pick_dst_config(c, vec, g, ch)) return 0;
  const int cw = (c + 31) / 32;
  // records: 16 B per edge when the mask fits two words, else (alpha, dscore) pairs + a mask array
  const size_t rec = cw <= 2 ? (size_t)e * 16 : (size_t)e * 8 + align_up((size_t)e * cw * sizeof(unsigned), 256);
  return align_up((size_t)bwd_dst_max_warps() * 2 * c * sizeof(float), 256) + align_up(rec, 256) + 1024;
}

struct BwdDstArgs {
  const int *rowptr, *col, *csr_to_csc, *order;
  const uint8_t* dst_is_src;
  const float *Hs, *Ht, *af_t2s, *af_s2t;
  float slope;
  long long n;
  int c, cw;
  const float *out, *row_max, *row_sum, *gout;
  float *gHs, *gHt;
  unsigned *erec, *emask;
  float* part;
};

template <int VEC, int G, int CH, int U, int MINB, int ES = 1>
static int launch_bwd_dst_cfg(const BwdDstArgs& a, int& nwarps_out, cudaStream_t stream) {
  auto kern = gatv2_bwd_dst_kernel<VEC, G, CH, U, MINB, ES>;
  static int occ = 0;                               // per instantiation
  if (occ == 0) {
    int o = 0;
    BGNN_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kBwdDstThreads, 0));
    occ = o < 1 ? 1 : (o > kBwdDstMaxCtasPerSm ? kBwdDstMaxCtasPerSm : o);
  }
  constexpr int WPC = kBwdDstThreads / 32;
  const long long nsets = (a.n + (32 / (G * ES)) - 1) / (32 / (G * ES));
  long long ctas = (nsets + WPC - 1) / WPC;
  if (ctas > (long long)kNumSMs * occ) ctas = (long long)kNumSMs * occ;
  nwarps_out = (int)(ctas * WPC);
  kern<<<(unsigned)ctas, kBwdDstThreads, 0, stream>>>(a.rowptr, a.col, a.csr_to_csc, a.order, a.dst_is_src, a.Hs, a.Ht, a.af_t2s,
                                                      a.af_s2t, a.slope, a.n, a.c, a.cw, a.out, a.row_max, a.row_sum,
                                                      a.gout, a.gHs, a.gHt, a.erec, a.emask, a.part);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

template <int VEC, int G, int CH>
static int launch_bwd_dst(const BwdDstArgs& a, int& nwarps_out, cudaStream_t stream) {
  if constexpr (VEC == 4 && G == 8 && CH == 2) {    // tuning experiments on the c = 64 mapping
    static const int variant = getenv("BGNN_GAT_VARIANT") ? atoi(getenv("BGNN_GAT_VARIANT")) : 0;
    if (variant == 1) return launch_bwd_dst_cfg<VEC, G, CH, 4, 4>(a, nwarps_out, stream);
    if (variant == 2) return launch_bwd_dst_cfg<VEC, G, CH, 2, 4>(a, nwarps_out, stream);
    if (variant == 3) return launch_bwd_dst_cfg<VEC, G, CH, 2, 6>(a, nwarps_out, stream);
  }
  constexpr int EPL = VEC * CH;
  if constexpr (G == 1) {                           // narrow rows: 8 lanes split the edges of a row
    constexpr int MINB1 = (EPL <= 2) ? 8 : (EPL <= 4 ? 6 : 5);
    return launch_bwd_dst_cfg<VEC, G, CH, 2, MINB1, 8>(a, nwarps_out, stream);
  } else {
    constexpr int U = (EPL >= 8) ? 2 : 4;
    constexpr int MINB = (EPL > 8) ? 3 : 5;
    return launch_bwd_dst_cfg<VEC, G, CH, U, MINB>(a, nwarps_out, stream);
  }
}

template <int VEC>
static int dispatch_bwd_dst(int g, int ch, const BwdDstArgs& a, int& nwarps_out, cudaStream_t stream) {
  constexpr int CHW = 8 / VEC;                     // chunks per lane once the row needs more than one lane
  switch (g) {
    case 1:
      if (ch == 1) return launch_bwd_dst<VEC, 1, 1>(a, nwarps_out, stream);
      if (ch == 2) return launch_bwd_dst<VEC, 1, 2>(a, nwarps_out, stream);
      if (ch == 4) return launch_bwd_dst<VEC, 1, (VEC <= 2 ? 4 : 2)>(a, nwarps_out, stream);
      return launch_bwd_dst<VEC, 1, (VEC == 1 ? 8 : CHW)>(a, nwarps_out, stream);
    case 2: return launch_bwd_dst<VEC, 2, CHW>(a, nwarps_out, stream);
    case 4: return launch_bwd_dst<VEC, 4, CHW>(a, nwarps_out, stream);
    case 8: return launch_bwd_dst<VEC, 8, CHW>(a, nwarps_out, stream);
    case 16: return launch_bwd_dst<VEC, 16, CHW>(a, nwarps_out, stream);
    default:
      if (ch == CHW) return launch_bwd_dst<VEC, 32, CHW>(a, nwarps_out, stream);
      if (VEC >= 2 && ch == 2 * CHW) return launch_bwd_dst<VEC, 32, (VEC >= 2 ? 2 * CHW : CHW)>(a, nwarps_out, stream);
      return BGNN_ERR_UNSUPPORTED;
  }
}

int launch_gatv2_bwd(const int* rowptr, const int* col, const int* t_rowptr, const int* t_col, const int* csr_to_csc,
                     const int* order, const int* t_order, long long e, const uint8_t* dst_is_src, const float* Hs, const float* Ht, const float* af_t2s,
                     const float* af_s2t, float slope, long long n, int c, const float* out, const float* row_max,
                     const float* row_sum, const float* gout, float* gHs, float* gHt, float* g_af_t2s,
                     float* g_af_s2t, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (n <= 0) return BGNN_OK;
  int vec, g, ch, dvec, dg, dch;
  if (!pick_row_config(c, vec, g, ch) || !pick_dst_config(c, dvec, dg, dch)) return BGNN_ERR_UNSUPPORTED;
  long long blocks = bwd_blocks(n, g);
  const int cw = (c + 31) / 32;
  Workspace w(ws, ws_bytes);
  float* part = w.take<float>(bwd_dst_max_warps() * 2 * c);
  unsigned* erec = w.take<unsigned>(cw <= 2 ? e * 4 : e * 2);
  unsigned* emask = cw <= 2 ? nullptr : w.take<unsigned>(e * cw);
  if (!w.ok()) return BGNN_ERR_WORKSPACE;
  const BwdDstArgs a{rowptr, col, csr_to_csc, order, dst_is_src, Hs, Ht, af_t2s, af_s2t, slope, n, c, cw,
                     out, row_max, row_sum, gout, gHs, gHt, erec, emask, part};
  int nparts = 0;
  const int rc = dvec == 4 ? dispatch_bwd_dst<4>(dg, dch, a, nparts, stream)
               : dvec == 2 ? dispatch_bwd_dst<2>(dg, dch, a, nparts, stream)
                           : dispatch_bwd_dst<1>(dg, dch, a, nparts, stream);
  if (rc != BGNN_OK) return rc;
  reduce_partials_kernel<<<2 * c, 256, 0, stream>>>(part, nparts, 2 * c, g_af_t2s, g_af_s2t, c);
  BGNN_LAUNCH_CHECK();
#define CALL(V, G_, C_)                                                                                            \
  gatv2_bwd_src_kernel<V, G_, C_><<<(unsigned)blocks, 256, 0, stream>>>(t_rowptr, t_col, t_order, dst_is_src, af_t2s,      \
      af_s2t, slope, n, c, cw, erec, emask, gout, gHs, gHt)
  BGNN_ROW_DISPATCH(vec, g, ch, CALL);
#undef CALL
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

}  // namespace bgnn
