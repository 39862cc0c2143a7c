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
// H[src] per edge, no [E,C] intermediates, no atomics.  In training the forward also streams out the
// per-edge score (4 B per edge, CSR order) so that the backward never recomputes a score.
//
// Backward (K3b), two atomic-free, deterministic passes.  With t = H_j + H_i, p = [t > 0],
// lrelu(t) = slope t + (1 - slope) p t and ds_ij = alpha_ij (gout_i . H_j - gout_i . out_i):
//   pass A (CSR by dst, gathers H[src]):   ds, the leaky-relu branch bits p (packed from sign bits),
//        T_i[f] = sum_j p ds, S_i = sum_j ds;   dH[i] += a (.) (slope S_i + (1-slope) T_i)     [destination side]
//        d a += H_i (.) (slope S_i + (1-slope) T_i);   one 16-byte record (alpha, ds, p bits) per edge
//   pass B (CSC by src, gathers gout[dst], streams the records):
//        G_j = sum_i alpha gout_i,  T_j[f] = sum_i p ds,  S_j = sum_i ds   (per destination domain)
//        dH[j] += G_j + a (.) (slope S_j + (1-slope) T_j);   d a += H_j (.) (slope S_j + (1-slope) T_j)
// i.e. per edge and feature pass A costs one FMA (the dot product), one subtraction + one funnel shift (branch bit)
// and one predicated add; everything that involves `a` or the leaky-relu values happens once per ROW.
//
// Destination-partitioned (multi-GPU) layout: a rank processes n_rows destination rows with LOCAL row ids
// (rowptr, out, gout, row statistics) whose global node id is row + row_off; H, dst_is_src and the gradients
// w.r.t. H are indexed by global id (col entries are global); the transposed CSR has n_src rows (all sources)
// whose entries are local destination ids.  Single GPU: row_off = 0, n_src = n_rows.
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
// SCORE_ONLY: only the per-edge scores are produced (for callers of the backward that did not keep them).
template <int VEC, int G, int CH, bool SCORE_ONLY>
__global__ void __launch_bounds__(256)
gatv2_fwd_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ order,
                 const uint8_t* __restrict__ dst_is_src,
                 const float* __restrict__ Hs, const float* __restrict__ Ht, const float* __restrict__ af_t2s,
                 const float* __restrict__ af_s2t, float slope, long long n, long long row_off, int c, float* __restrict__ out,
                 float* __restrict__ row_max, float* __restrict__ row_sum, float* __restrict__ score) {
  const int lane = threadIdx.x & 31;
  const int lane_g = threadIdx.x % G;
  const unsigned mask = group_mask<G>(lane);
  // `order` (optional): rows by descending degree -- hub rows start first instead of forming the tail of the
  // launch, and the rows that share a warp have similar lengths
  const long long slot = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (slot >= n) return;  // a group leaves together; shuffles below use the group's own mask
  const long long row = order ? (long long)__ldg(order + slot) : slot;
  const long long grow = row + row_off;              // global node id of the destination
  const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
  if (beg == end) {
    // no incoming edge: the aggregate is 0; skip every feature load
    if (SCORE_ONLY) return;
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
  const bool is_src = dst_is_src[grow] != 0;
  const float* __restrict__ H = is_src ? Hs : Ht;
  const float* __restrict__ a = is_src ? af_t2s : af_s2t;
  Chunk<VEC> hi[CH], av[CH], acc[CH];
  bool cok[CH];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) {
    int c0 = (ch * G + lane_g) * VEC;
    cok[ch] = c0 < c;
    hi[ch] = ld_chunk<VEC>(H + grow * c + c0, cok[ch]);
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
    if (score) {
      // the scores of the batch go out as one contiguous run: lane u of the group stores edge e + u
      if (G >= U) {
        const float sv = lane_g == 0 ? s[0] : (lane_g == 1 ? s[1] : (lane_g == 2 ? s[2] : s[3]));
        if (lane_g < U && e + lane_g < end) score[e + lane_g] = sv;
      } else if (lane_g == 0) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (e + u < end) score[e + u] = s[u];
      }
    }
    if (SCORE_ONLY) continue;
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
  if (SCORE_ONLY) return;
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
                     const float* af_t2s, const float* af_s2t, float slope, long long n, long long row_off, int c, float* out,
                     float* row_max, float* row_sum, float* score, int score_only, cudaStream_t stream) {
  if (n <= 0) return BGNN_OK;
  int vec, g, ch;
  if (!pick_row_config(c, vec, g, ch)) return BGNN_ERR_UNSUPPORTED;
  long long blocks = (n * g + 255) / 256;
#define CALL(V, G_, C_)                                                                                                  \
  do {                                                                                                                   \
    if (score_only)                                                                                                      \
      gatv2_fwd_kernel<V, G_, C_, true><<<(unsigned)blocks, 256, 0, stream>>>(rowptr, col, order, dst_is_src, Hs, Ht,   \
          af_t2s, af_s2t, slope, n, row_off, c, out, row_max, row_sum, score);                                           \
    else                                                                                                                 \
      gatv2_fwd_kernel<V, G_, C_, false><<<(unsigned)blocks, 256, 0, stream>>>(rowptr, col, order, dst_is_src, Hs, Ht,  \
          af_t2s, af_s2t, slope, n, row_off, c, out, row_max, row_sum, score);                                           \
  } while (0)
  BGNN_ROW_DISPATCH(vec, g, ch, CALL);
#undef CALL
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

// ------------------------------------------------------------------------------------------ backward
// Pass A: per destination row (CSR), scores read back from the forward.  Writes
//   * the destination-side gradient of row i into gHs (source-domain row) or gHt (target-domain row) at row + row_off,
//   * per-WARP partial sums of the destination-side part of d a_f into ga_part[warp][2][c],
//   * one record per edge, stored at the edge's slot in the TRANSPOSED CSR so that pass B streams them:
//       ea   = alpha_ij with the destination's domain in the sign bit (negative: source-domain destination)
//       eds  = d score_ij = alpha_ij (gout_i . H_j - gout_i . out_i)
//       emask[CW] = bit f set iff H_j[f] + H_i[f] > 0 (the leaky-relu branch)
// so that pass B needs neither H[dst] nor the softmax statistics again.
//
// Row mapping: a group of G lanes owns a row and every lane owns EPL = VEC*CH CONTIGUOUS features (8, or 16 for
// c > 256).  Few lanes per row means few shuffle steps per edge, and a lane's branch bits are exactly byte(s)
// lane_g*EPL/8.. of the edge's mask, which it stores itself.  The branch bit of a feature is the SIGN bit of
// (-H_i) - H_j, which is set exactly when H_j + H_i > 0 (+0 when the sum is 0, like the reference's t > 0), so
// the bits of a lane are packed by one funnel shift per feature.  U edges are in flight per group; warps are
// persistent over a static round-robin of row sets (no CTA barrier: a long row delays only its own warp, and the
// d a_f summation order stays fixed).
struct BwdDstParams {
  const int *rowptr, *col, *csr_to_csc, *order;
  const uint8_t* dst_is_src;
  const float *Hs, *Ht, *af_t2s, *af_s2t;
  float slope;
  long long row_off;
  int c, cw;
  const float *out, *row_max, *row_sum, *score, *gout;
  float *gHs, *gHt;
  unsigned *erec, *emask;
};

// The slots [slot_off, slot_off + n) of the row order, processed by the warps wid = 0 .. nwarps-1 of a launch (or of a
// part of one); s_ga: the CTA's [2 * EPL][128] scratch columns.
template <int VEC, int G, int CH, int U, int ES>
__device__ __forceinline__ void bwd_dst_rows(const BwdDstParams& P, long long n, long long slot_off, long long wid, long long nwarps,
                                             float (*s_ga)[128], float* __restrict__ ga_part) {
  const int* __restrict__ rowptr = P.rowptr;
  const int* __restrict__ col = P.col;
  const int* __restrict__ csr_to_csc = P.csr_to_csc;
  const int* __restrict__ order = P.order;
  const uint8_t* __restrict__ dst_is_src = P.dst_is_src;
  const float* __restrict__ Hs = P.Hs;
  const float* __restrict__ Ht = P.Ht;
  const float* __restrict__ af_t2s = P.af_t2s;
  const float* __restrict__ af_s2t = P.af_s2t;
  const float slope = P.slope;
  const long long row_off = P.row_off;
  const int c = P.c, cw = P.cw;
  const float* __restrict__ out = P.out;
  const float* __restrict__ row_max = P.row_max;
  const float* __restrict__ row_sum = P.row_sum;
  const float* __restrict__ score = P.score;
  const float* __restrict__ gout = P.gout;
  float* __restrict__ gHs = P.gHs;
  float* __restrict__ gHt = P.gHt;
  unsigned* __restrict__ erec = P.erec;
  unsigned* __restrict__ emask = P.emask;
  // ES > 1: ES groups of G lanes share a row and split its EDGES (group s takes edges s, s + ES, ...).  Used for
  // narrow rows (G == 1: consecutive lanes read consecutive col entries) and for the HUB rows of wide ones (the
  // longest rows of the degree order get a whole warp each, so that a 2000-edge row is not one group's serial loop).
  constexpr int RL = G * ES;                       // lanes per row
  static_assert(RL <= 32, "a row is owned by at most one warp");
  constexpr int RPW = 32 / RL;                     // rows per warp
  constexpr int EPL = VEC * CH;                    // contiguous features per lane
  static_assert(G == 1 || EPL == 8 || EPL == 16, "multi-lane rows own whole mask bytes");
  const int lane = threadIdx.x & 31;
  const int lane_g = lane % G;
  const int sub = (lane / G) % ES;                 // which share of the row's edges
  const unsigned gmask = group_mask<G>(lane);
  const unsigned rmask = group_mask<RL>(lane);
  const int col0 = lane_g * EPL;
  const float oms = 1.f - slope;
  bool cok[CH];
  // s_ga: running d a_f sums of this lane, per destination domain: thread-private columns of shared memory
  // (kept out of the register file; touched once per ROW)
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
    const long long row = order ? (long long)__ldg(order + slot_off + slot) : slot_off + slot;
    const long long grow = row + row_off;
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    const bool is_src = dst_is_src[grow] != 0;
    float* __restrict__ gdst = is_src ? gHs : gHt;
    Chunk<VEC> gi[CH];                             // sum over the row's edges of [t > 0] ds, per feature
#pragma unroll
    for (int ch = 0; ch < CH; ++ch)
#pragma unroll
      for (int i = 0; i < VEC; ++i) gi[ch].v[i] = 0.f;
    if (beg < end) {                               // no incoming edge: the destination-side gradient is 0
      const float* __restrict__ H = is_src ? Hs : Ht;
      Chunk<VEC> nhi[CH], go[CH];
      float dpart = 0.f;
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) {
        nhi[ch] = ld_chunk<VEC>(H + grow * c + col0 + ch * VEC, cok[ch]);
        go[ch] = ld_chunk<VEC>(gout + row * c + col0 + ch * VEC, cok[ch]);
        const Chunk<VEC> oi = ld_chunk<VEC>(out + row * c + col0 + ch * VEC, cok[ch]);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          dpart = fmaf(go[ch].v[i], oi.v[i], dpart);
          nhi[ch].v[i] = 0.f - nhi[ch].v[i];      // never -0: (-H_i) - H_j is then +0 whenever the sum is 0
        }
      }
      const float Di = gsum<G>(dpart, gmask);
      const float m = row_max[row];
      const float inv = 1.0f / (row_sum[row] + 1e-16f);
      float dsum = 0.f;                            // sum of d score over the row's edges
      int nj[U], np[U];                            // (source, transposed slot, score) of the NEXT batch: its gathers start at once
      float nsc[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int ee = beg + u * ES + sub;
        const bool ok = ee < end;
        nj[u] = ok ? __ldg(col + ee) : -1;
        np[u] = ok ? __ldg(csr_to_csc + ee) : 0;
        nsc[u] = ok ? __ldg(score + ee) : -INFINITY;
      }
      for (int e = beg; e < end; e += U * ES) {
        int j[U], pos[U];
        float sc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { j[u] = nj[u]; pos[u] = np[u]; sc[u] = nsc[u]; }
        Chunk<VEC> hj[U][CH];                      // H[src]; overwritten by (-H[dst]) - H[src] below
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
          nsc[u] = ok ? __ldg(score + ee) : -INFINITY;
        }
        float dp[U];
        unsigned bits[U];                          // leaky-relu branch bits of this lane's features (bit = feature)
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float d = 0.f;
          unsigned b = 0u;
#pragma unroll
          for (int k = EPL - 1; k >= 0; --k) {     // descending: the funnel shift leaves feature k in bit k
            const int ch = k / VEC, i = k % VEC;
            const float x = hj[u][ch].v[i];
            d = fmaf(go[ch].v[i], x, d);
            const float tn = nhi[ch].v[i] - x;     // sign bit set  <=>  H_j + H_i > 0
            b = __funnelshift_l(__float_as_uint(tn), b, 1);
            hj[u][ch].v[i] = tn;
          }
          dp[u] = d;
          bits[u] = b;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) dp[u] = gsum<G>(dp[u], gmask);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float alpha = (j[u] >= 0) ? expf(sc[u] - m) * inv : 0.f;
          const float ds = alpha * (dp[u] - Di);
          dsum += ds;
#pragma unroll
          for (int ch = 0; ch < CH; ++ch)
#pragma unroll
            for (int i = 0; i < VEC; ++i) gi[ch].v[i] += (hj[u][ch].v[i] < 0.f) ? ds : 0.f;
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
            for (int i = 0; i < VEC; ++i) gi[ch].v[i] += __shfl_xor_sync(rmask, gi[ch].v[i], o);
        }
      }
      // sum_j ds_j lrelu'(t_j) = slope sum_j ds_j + (1 - slope) sum_{t_j > 0} ds_j, per feature: times a -> d H[dst],
      // times H[dst] -> destination-side part of d a
      const float sds = slope * dsum;
      const float* __restrict__ a = is_src ? af_t2s : af_s2t;
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) {
        const Chunk<VEC> av = ld_chunk<VEC>(a + col0 + ch * VEC, cok[ch]);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float w = fmaf(oms, gi[ch].v[i], sds);
          gi[ch].v[i] = av.v[i] * w;
          if (sub == 0) s_ga[(is_src ? 0 : EPL) + ch * VEC + i][threadIdx.x] -= nhi[ch].v[i] * w;   // nhi = -H[dst]
        }
      }
    }
    if (sub == 0 && gdst) {
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) st_chunk<VEC>(gdst + grow * c + col0 + ch * VEC, gi[ch], cok[ch]);
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

template <int VEC, int G, int CH, int U, int MINB, int ES>
__global__ void __launch_bounds__(128, MINB)
gatv2_bwd_dst_kernel(const BwdDstParams P, long long n, long long slot_off, float* __restrict__ ga_part) {
  __shared__ float s_ga[2 * VEC * CH][128];
  bwd_dst_rows<VEC, G, CH, U, ES>(P, n, slot_off, (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5),
                                  (long long)gridDim.x * (blockDim.x >> 5), s_ga, ga_part);
}

// One launch for a degree-ordered graph: the first hub_ctas CTAs take the n_hub longest rows with a WARP per row (32 / G
// groups split the edges), the others the remaining rows with a group per row.  The hub CTAs start first and are done
// long before the launch ends, so the longest row costs nothing on top (as a launch of its own it took 0.18 ms).
template <int VEC, int G, int CH, int U, int MINB>
__global__ void __launch_bounds__(128, MINB)
gatv2_bwd_dst_hub_kernel(const BwdDstParams P, long long n, long long n_hub, int hub_ctas, float* __restrict__ ga_part) {
  __shared__ float s_ga[2 * VEC * CH][128];
  const long long wpc = blockDim.x >> 5;
  if ((int)blockIdx.x < hub_ctas)
    bwd_dst_rows<VEC, G, CH, U, 32 / G>(P, n_hub, 0, (long long)blockIdx.x * wpc + (threadIdx.x >> 5), (long long)hub_ctas * wpc,
                                        s_ga, ga_part);
  else
    bwd_dst_rows<VEC, G, CH, U, 1>(P, n - n_hub, n_hub, (long long)(blockIdx.x - hub_ctas) * wpc + (threadIdx.x >> 5),
                                   (long long)(gridDim.x - hub_ctas) * wpc, s_ga, ga_part + (size_t)hub_ctas * wpc * 2 * P.c);
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
constexpr int kBwdSrcThreads = 256;
static long long bwd_dst_max_warps() { return (long long)kNumSMs * kBwdDstMaxCtasPerSm * (kBwdDstThreads / 32); }

// Column-wise reduction of the per-warp partials: one CTA per column, strided partial sums then a
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

// Pass B: per source row j over its outgoing edges (transposed CSR); one row per group of G lanes, CTAs scheduled
// by the hardware in degree order (hub rows first, no tail).  Streams the edge records of pass A and gathers only
// gout[dst].  Per destination domain it accumulates
//   G = sum alpha gout_i,  T[f] = sum_{branch bit f} ds,  S = sum ds
// and finishes the row with   dH[j] = own + G + a (.) (slope S + (1-slope) T),   d a += H_j (.) (slope S + (1-slope) T)
// (own = the destination-side part pass A left in the array of the row's own domain, for rows this rank owns).
// The d a contributions of the CTA's rows are summed in a fixed order (warp butterfly, then over the warps in shared
// memory) into ONE partial per CTA: deterministic.
template <int VEC, int G, int CH, int MINB>
__global__ void __launch_bounds__(kBwdSrcThreads, MINB)
gatv2_bwd_src_kernel(const int* __restrict__ t_rowptr, const int* __restrict__ t_col, const int* __restrict__ t_order,
                     const uint8_t* __restrict__ dst_is_src, const float* __restrict__ Hs, const float* __restrict__ Ht,
                     const float* __restrict__ af_t2s, const float* __restrict__ af_s2t, float slope, long long n_src,
                     long long own_lo, long long own_n, int c, int cw,
                     const unsigned* __restrict__ erec, const unsigned* __restrict__ emask,
                     const float* __restrict__ gout, float* __restrict__ gHs, float* __restrict__ gHt,
                     float* __restrict__ ga_part) {
  extern __shared__ float s_da[];                    // [warps][2c]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lane_g = threadIdx.x % G;
  const long long slot = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const bool live = slot < n_src;                    // whole groups are live or not
  const float oms = 1.f - slope;
  Chunk<VEC> das[CH], dat[CH];
  bool cok[CH];
  int wsel[CH], bsel[CH];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) {
    int c0 = (ch * G + lane_g) * VEC;
    cok[ch] = c0 < c;
    wsel[ch] = min(c0 >> 5, cw - 1);
    bsel[ch] = c0 & 31;
#pragma unroll
    for (int i = 0; i < VEC; ++i) { das[ch].v[i] = 0.f; dat[ch].v[i] = 0.f; }
  }
  if (live) {
    const long long row = t_order ? (long long)__ldg(t_order + slot) : slot;
    const int beg = __ldg(t_rowptr + row), end = __ldg(t_rowptr + row + 1);
    const bool owned = row >= own_lo && row < own_lo + own_n;
    const bool me_src = dst_is_src[row] != 0;
    Chunk<VEC> gs[CH], gt[CH], ts[CH], tt[CH];
#pragma unroll
    for (int ch = 0; ch < CH; ++ch)
#pragma unroll
      for (int i = 0; i < VEC; ++i) { gs[ch].v[i] = 0.f; gt[ch].v[i] = 0.f; ts[ch].v[i] = 0.f; tt[ch].v[i] = 0.f; }
    float ss = 0.f, st = 0.f;
    constexpr int U = 4;
    int nxt[U];                                      // destination ids of the NEXT batch: their gathers start at once
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
        ss += dss[u];
        st += dst_[u];
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int ch = 0; ch < CH; ++ch)
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const bool pos = (mk[u][ch] >> (bsel[ch] + i)) & 1u;
            const float g0 = go[u][ch].v[i];
            gs[ch].v[i] = fmaf(als[u], g0, gs[ch].v[i]);
            gt[ch].v[i] = fmaf(alt[u], g0, gt[ch].v[i]);
            ts[ch].v[i] += pos ? dss[u] : 0.f;
            tt[ch].v[i] += pos ? dst_[u] : 0.f;
          }
    }
    const float sls = slope * ss, slt = slope * st;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      const int c0 = (ch * G + lane_g) * VEC;
      if (!cok[ch]) continue;
      if (beg < end) {
        const Chunk<VEC> hs = ld_chunk<VEC>(Hs + row * c + c0, true), ht = ld_chunk<VEC>(Ht + row * c + c0, true);
        const Chunk<VEC> as_ = ld_chunk<VEC>(af_t2s + c0, true), at_ = ld_chunk<VEC>(af_s2t + c0, true);   // L1-resident
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float us = fmaf(oms, ts[ch].v[i], sls), ut = fmaf(oms, tt[ch].v[i], slt);
          gs[ch].v[i] = fmaf(as_.v[i], us, gs[ch].v[i]);
          gt[ch].v[i] = fmaf(at_.v[i], ut, gt[ch].v[i]);
          das[ch].v[i] = hs.v[i] * us;
          dat[ch].v[i] = ht.v[i] * ut;
        }
      }
      if (owned) {
        // destination-side part of this row from pass A lives in the array of the row's own domain
        const float* ownp = me_src ? gHs : gHt;
        if (ownp) {
          const Chunk<VEC> own = ld_chunk<VEC>(ownp + row * c + c0, true);
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            gs[ch].v[i] += me_src ? own.v[i] : 0.f;
            gt[ch].v[i] += me_src ? 0.f : own.v[i];
          }
        }
      }
      if (gHs) st_chunk<VEC>(gHs + row * c + c0, gs[ch], true);
      if (gHt) st_chunk<VEC>(gHt + row * c + c0, gt[ch], true);
    }
  }
  // source-side part of d a_f: the warp's groups (fixed butterfly), then the CTA's warps in order -> one partial per CTA
#pragma unroll
  for (int ch = 0; ch < CH; ++ch)
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float vs = das[ch].v[i], vt = dat[ch].v[i];
#pragma unroll
      for (int o = G; o < 32; o <<= 1) {
        vs += __shfl_xor_sync(0xffffffffu, vs, o);
        vt += __shfl_xor_sync(0xffffffffu, vt, o);
      }
      const int cc = (ch * G + lane_g) * VEC + i;
      if (lane < G && cc < c) {
        s_da[warp * 2 * c + cc] = vs;
        s_da[warp * 2 * c + c + cc] = vt;
      }
    }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * c; j += kBwdSrcThreads) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kBwdSrcThreads / 32; ++w) acc += s_da[w * 2 * c + j];
    ga_part[(long long)blockIdx.x * 2 * c + j] = acc;
  }
}

static long long bwd_src_blocks(long long n_src, int c) {
  int vec, g, ch;
  if (!pick_row_config(c, vec, g, ch)) return 0;
  const long long b = (n_src * g + kBwdSrcThreads - 1) / kBwdSrcThreads;
  return b < 1 ? 1 : b;
}

size_t gatv2_bwd_workspace_bytes(long long n, long long e, int c) {
  int vec, g, ch;
  if (!pick_row_config(c, vec, g, ch) || !pick_dst_config(c, vec, g, ch)) return 0;
  const int cw = (c + 31) / 32;
  // records: 16 B per edge when the mask fits two words, else (alpha, dscore) pairs + a mask array;
  // + e floats of scores for callers that did not keep the forward's; d a_f partials: one per warp of the two pass-A
  // launches (hub rows, the rest) and one per CTA of pass B
  const size_t rec = cw <= 2 ? (size_t)e * 16 : (size_t)e * 8 + align_up((size_t)e * cw * sizeof(unsigned), 256);
  return align_up((size_t)(2 * bwd_dst_max_warps() + bwd_src_blocks(n, c)) * 2 * c * sizeof(float), 256) + align_up(rec, 256) +
         align_up((size_t)e * sizeof(float), 256) + 1024;
}

struct BwdDstArgs {
  BwdDstParams p;
  long long n, slot_off;
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
  kern<<<(unsigned)ctas, kBwdDstThreads, 0, stream>>>(a.p, a.n, a.slot_off, a.part);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

// hub rows + the rest in one launch (see gatv2_bwd_dst_hub_kernel)
template <int VEC, int G, int CH, int U, int MINB>
static int launch_bwd_dst_merged(const BwdDstArgs& a, long long n_hub, int& nwarps_out, cudaStream_t stream) {
  auto kern = gatv2_bwd_dst_hub_kernel<VEC, G, CH, U, MINB>;
  static int occ = 0;
  if (occ == 0) {
    int o = 0;
    BGNN_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kBwdDstThreads, 0));
    occ = o < 1 ? 1 : (o > kBwdDstMaxCtasPerSm ? kBwdDstMaxCtasPerSm : o);
  }
  constexpr int WPC = kBwdDstThreads / 32;
  const long long resident = (long long)kNumSMs * occ;
  // the hub rows hold a few per cent of the edges: a few per cent of the resident CTAs walk them (a warp per row,
  // longest first), so that the slots they free early are few
  long long hub_ctas = (n_hub + WPC - 1) / WPC;
  if (hub_ctas > resident / 24) hub_ctas = resident / 24 > 0 ? resident / 24 : 1;
  const long long nsets = (a.n - n_hub + (32 / G) - 1) / (32 / G);
  long long main_ctas = (nsets + WPC - 1) / WPC;
  if (main_ctas > resident - hub_ctas) main_ctas = resident - hub_ctas;
  if (main_ctas < 1) main_ctas = 1;
  nwarps_out = (int)((hub_ctas + main_ctas) * WPC);
  kern<<<(unsigned)(hub_ctas + main_ctas), kBwdDstThreads, 0, stream>>>(a.p, a.n, n_hub, (int)hub_ctas, a.part);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

template <int VEC, int G, int CH>
static int launch_bwd_dst(const BwdDstArgs& a, int& nwarps_out, cudaStream_t stream) {
  if constexpr (VEC == 4 && G == 8 && CH == 2) {    // tuning experiments on the c = 64 mapping
    static const int variant = getenv("BGNN_GAT_VARIANT") ? atoi(getenv("BGNN_GAT_VARIANT")) : 0;
    if (variant == 1) return launch_bwd_dst_cfg<VEC, G, CH, 2, 5>(a, nwarps_out, stream);
    if (variant == 2) return launch_bwd_dst_cfg<VEC, G, CH, 2, 8>(a, nwarps_out, stream);
    if (variant == 3) return launch_bwd_dst_cfg<VEC, G, CH, 4, 5>(a, nwarps_out, stream);
    if (variant == 4) return launch_bwd_dst_cfg<VEC, G, CH, 3, 5>(a, nwarps_out, stream);
    if (variant == 5) return launch_bwd_dst_cfg<VEC, G, CH, 4, 4>(a, nwarps_out, stream);
  }
  constexpr int EPL = VEC * CH;
  if constexpr (G == 1) {                           // narrow rows: 8 lanes split the edges of a row
    constexpr int MINB1 = (EPL <= 2) ? 8 : (EPL <= 4 ? 6 : 5);
    return launch_bwd_dst_cfg<VEC, G, CH, 2, MINB1, 8>(a, nwarps_out, stream);
  } else {
    constexpr int U = 2;
    constexpr int MINB = (EPL > 8) ? 3 : 6;
    return launch_bwd_dst_cfg<VEC, G, CH, U, MINB>(a, nwarps_out, stream);
  }
}

// Degree-ordered graph, rows of 2..16 lanes: the longest rows get a WARP each (ES = 32 / G groups split the row's edges)
// inside the same launch.  BGNN_GAT_HUBS=separate keeps the two-launch form for A/B measurements.
template <int VEC, int G, int CH>
static int launch_bwd_dst_hubs(const BwdDstArgs& a, long long n_hub, int& nwarps_out, cudaStream_t stream) {
  if constexpr (G > 1 && G < 32) {
    constexpr int EPL = VEC * CH;
    static const bool separate = getenv("BGNN_GAT_HUBS") && getenv("BGNN_GAT_HUBS")[0] == 's';
    if (!separate) return launch_bwd_dst_merged<VEC, G, CH, 2, (EPL > 8) ? 3 : 6>(a, n_hub, nwarps_out, stream);
    BwdDstArgs h = a;
    h.n = n_hub;
    int nh = 0, nr = 0;
    // 4 edges in flight per group: the single longest row bounds this launch
    int rc = launch_bwd_dst_cfg<VEC, G, CH, (EPL > 8) ? 2 : 4, (EPL > 8) ? 3 : 4, 32 / G>(h, nh, stream);
    if (rc != BGNN_OK) return rc;
    BwdDstArgs r = a;
    r.n = a.n - n_hub;
    r.slot_off = n_hub;
    r.part = a.part + (size_t)nh * 2 * a.p.c;
    if (r.n > 0) {
      rc = launch_bwd_dst<VEC, G, CH>(r, nr, stream);
      if (rc != BGNN_OK) return rc;
    }
    nwarps_out = nh + nr;
    return BGNN_OK;
  } else {
    nwarps_out = 0;
    return BGNN_ERR_UNSUPPORTED;
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

template <int VEC>
static int dispatch_bwd_dst_hubs(int g, const BwdDstArgs& a, long long n_hub, int& nwarps_out, cudaStream_t stream) {
  constexpr int CHW = 8 / VEC;
  switch (g) {
    case 2: return launch_bwd_dst_hubs<VEC, 2, CHW>(a, n_hub, nwarps_out, stream);
    case 4: return launch_bwd_dst_hubs<VEC, 4, CHW>(a, n_hub, nwarps_out, stream);
    case 8: return launch_bwd_dst_hubs<VEC, 8, CHW>(a, n_hub, nwarps_out, stream);
    case 16: return launch_bwd_dst_hubs<VEC, 16, CHW>(a, n_hub, nwarps_out, stream);
    default: nwarps_out = 0; return BGNN_ERR_UNSUPPORTED;
  }
}

// rows at the head of the degree order that get a warp each in pass A (only with a row order; G in 2..16)
constexpr long long kHubRows = 4096;

template <int VEC, int G, int CH, int MINB>
static int launch_bwd_src_min(const int* t_rowptr, const int* t_col, const int* t_order, const uint8_t* dst_is_src, const float* Hs,
                              const float* Ht, const float* af_t2s, const float* af_s2t, float slope, long long n_src,
                              long long own_lo, long long own_n, int c, int cw, const unsigned* erec, const unsigned* emask,
                              const float* gout, float* gHs, float* gHt, float* part, int& nparts_out, cudaStream_t stream) {
  auto kern = gatv2_bwd_src_kernel<VEC, G, CH, MINB>;
  const long long ctas = (n_src * G + kBwdSrcThreads - 1) / kBwdSrcThreads;
  const size_t dyn = (size_t)(kBwdSrcThreads / 32) * 2 * c * sizeof(float);
  nparts_out = (int)ctas;
  kern<<<(unsigned)ctas, kBwdSrcThreads, dyn, stream>>>(t_rowptr, t_col, t_order, dst_is_src, Hs, Ht, af_t2s, af_s2t, slope, n_src,
                                                        own_lo, own_n, c, cw, erec, emask, gout, gHs, gHt, part);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

template <int VEC, int G, int CH>
static int launch_bwd_src_cfg(const int* t_rowptr, const int* t_col, const int* t_order, const uint8_t* dst_is_src, const float* Hs,
                              const float* Ht, const float* af_t2s, const float* af_s2t, float slope, long long n_src,
                              long long own_lo, long long own_n, int c, int cw, const unsigned* erec, const unsigned* emask,
                              const float* gout, float* gHs, float* gHt, float* part, int& nwarps_out, cudaStream_t stream) {
#define BGNN_SRC_GO(M)                                                                                                          \
  launch_bwd_src_min<VEC, G, CH, M>(t_rowptr, t_col, t_order, dst_is_src, Hs, Ht, af_t2s, af_s2t, slope, n_src, own_lo, own_n, c, \
                                    cw, erec, emask, gout, gHs, gHt, part, nwarps_out, stream)
  if constexpr (VEC == 4 && G == 16 && CH == 1) {   // tuning experiments on the c = 64 mapping
    static const int variant = getenv("BGNN_GAT_BVARIANT") ? atoi(getenv("BGNN_GAT_BVARIANT")) : 0;
    if (variant == 1) return BGNN_SRC_GO(2);
    if (variant == 2) return BGNN_SRC_GO(3);
  }
  if constexpr (VEC * CH <= 4) return BGNN_SRC_GO(4);
  else return BGNN_SRC_GO(1);
#undef BGNN_SRC_GO
}

int launch_gatv2_bwd(const int* rowptr, const int* col, const int* t_rowptr, const int* t_col, const int* csr_to_csc,
                     const int* order, const int* t_order, long long e, const uint8_t* dst_is_src, const float* Hs, const float* Ht, const float* af_t2s,
                     const float* af_s2t, float slope, long long n, long long row_off, long long n_src, int c, const float* out,
                     const float* row_max, const float* row_sum, const float* score, const float* gout, float* gHs, float* gHt,
                     float* g_af_t2s, float* g_af_s2t, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (n_src <= 0) return BGNN_OK;
  int vec, g, ch, dvec, dg, dch;
  if (!pick_row_config(c, vec, g, ch) || !pick_dst_config(c, dvec, dg, dch)) return BGNN_ERR_UNSUPPORTED;
  const int cw = (c + 31) / 32;
  Workspace w(ws, ws_bytes);
  float* part = w.take<float>((2 * bwd_dst_max_warps() + bwd_src_blocks(n_src, c)) * 2 * c);
  unsigned* erec = w.take<unsigned>(cw <= 2 ? e * 4 : e * 2);
  unsigned* emask = cw <= 2 ? nullptr : w.take<unsigned>(e * cw);
  float* score_ws = w.take<float>(e);
  if (!w.ok()) return BGNN_ERR_WORKSPACE;
  if (!score && n > 0) {
    // the caller did not keep the forward's scores: one score-only sweep of the forward kernel
    const int rc = launch_gatv2_fwd(rowptr, col, order, dst_is_src, Hs, Ht, af_t2s, af_s2t, slope, n, row_off, c, nullptr,
                                    nullptr, nullptr, score_ws, 1, stream);
    if (rc != BGNN_OK) return rc;
    score = score_ws;
  }
  int npa = 0, npb = 0;
  if (n > 0) {
    BwdDstArgs a;
    a.p = BwdDstParams{rowptr, col, csr_to_csc, order, dst_is_src, Hs, Ht, af_t2s, af_s2t, slope, row_off, c, cw,
                       out, row_max, row_sum, score, gout, gHs, gHt, erec, emask};
    a.n = n;
    a.slot_off = 0;
    a.part = part;
    // the head of the degree order: one warp per row, so that a hub row is not a single group's serial loop
    const long long n_hub = (order && dg >= 2 && dg <= 16) ? (n < kHubRows ? n : kHubRows) : 0;
    int rc;
    if (n_hub > 0 && n > n_hub)
      rc = dvec == 4 ? dispatch_bwd_dst_hubs<4>(dg, a, n_hub, npa, stream)
         : dvec == 2 ? dispatch_bwd_dst_hubs<2>(dg, a, n_hub, npa, stream)
                     : dispatch_bwd_dst_hubs<1>(dg, a, n_hub, npa, stream);
    else
      rc = dvec == 4 ? dispatch_bwd_dst<4>(dg, dch, a, npa, stream)
         : dvec == 2 ? dispatch_bwd_dst<2>(dg, dch, a, npa, stream)
                     : dispatch_bwd_dst<1>(dg, dch, a, npa, stream);
    if (rc != BGNN_OK) return rc;
  }
  float* part_b = part + (size_t)npa * 2 * c;
  int rcb = BGNN_OK;
#define CALL(V, G_, C_)                                                                                                   \
  rcb = launch_bwd_src_cfg<V, G_, C_>(t_rowptr, t_col, t_order, dst_is_src, Hs, Ht, af_t2s, af_s2t, slope, n_src, row_off, n, c, \
                                      cw, erec, emask, gout, gHs, gHt, part_b, npb, stream)
  BGNN_ROW_DISPATCH(vec, g, ch, CALL);
#undef CALL
  if (rcb != BGNN_OK) return rcb;
  reduce_partials_kernel<<<2 * c, 256, 0, stream>>>(part, (long long)npa + npb, 2 * c, g_af_t2s, g_af_s2t, c);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

}  // namespace bgnn
