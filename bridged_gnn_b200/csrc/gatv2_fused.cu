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
gatv2_fwd_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const uint8_t* __restrict__ dst_is_src,
                 const float* __restrict__ Hs, const float* __restrict__ Ht, const float* __restrict__ af_t2s,
                 const float* __restrict__ af_s2t, float slope, long long n, int c, float* __restrict__ out,
                 float* __restrict__ row_max, float* __restrict__ row_sum) {
  const int lane = threadIdx.x & 31;
  const int lane_g = threadIdx.x % G;
  const unsigned mask = group_mask<G>(lane);
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (row >= n) return;  // a group leaves together; shuffles below use the group's own mask
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

int launch_gatv2_fwd(const int* rowptr, const int* col, const uint8_t* dst_is_src, const float* Hs, const float* Ht,
                     const float* af_t2s, const float* af_s2t, float slope, long long n, int c, float* out,
                     float* row_max, float* row_sum, cudaStream_t stream) {
  if (n <= 0) return BGNN_OK;
  int vec, g, ch;
  if (!pick_row_config(c, vec, g, ch)) return BGNN_ERR_UNSUPPORTED;
  long long blocks = (n * g + 255) / 256;
#define CALL(V, G_, C_)                                                                                        \
  gatv2_fwd_kernel<V, G_, C_><<<(unsigned)blocks, 256, 0, stream>>>(rowptr, col, dst_is_src, Hs, Ht, af_t2s,   \
                                                                      af_s2t, slope, n, c, out, row_max, row_sum)
  BGNN_ROW_DISPATCH(vec, g, ch, CALL);
#undef CALL
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

// ------------------------------------------------------------------------------------------ backward
// Pass A: per destination row (CSR).  Recomputes the scores from the saved (max, sum), and writes
//   * the destination-side gradient of row i into gHs (source-domain row) or gHt (target-domain row),
//   * per-CTA partial sums of d a_f into ga_part[cta][2][c],
//   * one record per edge, stored at the edge's slot in the TRANSPOSED CSR so that pass B streams them:
//       ea   = alpha_ij with the destination's domain in the sign bit (negative: source-domain destination)
//       eds  = d score_ij = alpha_ij (gout_i . H_j - gout_i . out_i)
//       emask[CW] = bit c set iff H_j[c] + H_i[c] > 0 (the leaky-relu branch)
// so that pass B needs neither H[dst] nor the softmax statistics again.
template <int VEC, int G, int CH>
__global__ void __launch_bounds__(256)
gatv2_bwd_dst_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ csr_to_csc,
                     const uint8_t* __restrict__ dst_is_src, const float* __restrict__ Hs, const float* __restrict__ Ht,
                     const float* __restrict__ af_t2s, const float* __restrict__ af_s2t, float slope, long long n, int c,
                     int cw, const float* __restrict__ out, const float* __restrict__ row_max,
                     const float* __restrict__ row_sum, const float* __restrict__ gout, float* __restrict__ gHs,
                     float* __restrict__ gHt, unsigned* __restrict__ erec, unsigned* __restrict__ emask,
                     float* __restrict__ ga_part) {
  extern __shared__ float s_ga[];  // [groups][2][c]: every group owns a slice -> no atomics, fixed order
  constexpr int GROUPS = 256 / G;
  constexpr int LPW = 32 / VEC;                    // lanes that share one 32-column mask word
  constexpr int WG = (G < LPW) ? G : LPW;          // lanes of this group that share a word
  for (int t = threadIdx.x; t < GROUPS * 2 * c; t += blockDim.x) s_ga[t] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int lane_g = threadIdx.x % G;
  const unsigned mask = group_mask<G>(lane);
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int beg0 = row < n ? __ldg(rowptr + row) : 0, end0 = row < n ? __ldg(rowptr + row + 1) : 0;
  if (row < n && beg0 == end0) {
    // no incoming edge: the destination-side gradient is 0 (see the forward kernel)
    Chunk<VEC> z;
#pragma unroll
    for (int i = 0; i < VEC; ++i) z.v[i] = 0.f;
    const bool is_src = dst_is_src[row] != 0;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      const int c0 = (ch * G + lane_g) * VEC;
      st_chunk<VEC>((is_src ? gHs : gHt) + row * c + c0, z, c0 < c);
    }
  } else if (row < n) {
    const bool is_src = dst_is_src[row] != 0;
    const float* __restrict__ H = is_src ? Hs : Ht;
    const float* __restrict__ a = is_src ? af_t2s : af_s2t;
    Chunk<VEC> hi[CH], av[CH], go[CH], gi[CH], ga[CH];
    bool cok[CH];
    float dpart = 0.f;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      int c0 = (ch * G + lane_g) * VEC;
      cok[ch] = c0 < c;
      hi[ch] = ld_chunk<VEC>(H + row * c + c0, cok[ch]);
      av[ch] = ld_chunk<VEC>(a + c0, cok[ch]);
      go[ch] = ld_chunk<VEC>(gout + row * c + c0, cok[ch]);
      Chunk<VEC> oi = ld_chunk<VEC>(out + row * c + c0, cok[ch]);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        dpart = fmaf(go[ch].v[i], oi.v[i], dpart);
        gi[ch].v[i] = 0.f;
        ga[ch].v[i] = 0.f;
      }
    }
    const float Di = gsum<G>(dpart, mask);
    const float m = row_max[row];
    const float inv = 1.0f / (row_sum[row] + 1e-16f);
    const int beg = beg0, end = end0;
    constexpr int U = 2;
    for (int e = beg; e < end; e += U) {
      int j[U], pos[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = e + u < end;
        j[u] = ok ? __ldg(col + e + u) : -1;
        pos[u] = ok ? __ldg(csr_to_csc + e + u) : 0;
      }
      Chunk<VEC> hj[U][CH];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int ch = 0; ch < CH; ++ch)
          hj[u][ch] = ld_chunk<VEC>(H + (long long)(j[u] < 0 ? 0 : j[u]) * c + (ch * G + lane_g) * VEC, cok[ch] && j[u] >= 0);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float sp = 0.f, dp = 0.f;
        unsigned bits[CH];
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
          bits[ch] = 0u;
          const int b0 = ((ch * G + lane_g) * VEC) & 31;
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const float t = hj[u][ch].v[i] + hi[ch].v[i];
            sp = fmaf(av[ch].v[i], lrelu(t, slope), sp);
            dp = fmaf(go[ch].v[i], hj[u][ch].v[i], dp);
            bits[ch] |= (t > 0.f ? 1u : 0u) << (b0 + i);
          }
        }
        sp = gsum<G>(sp, mask);
        dp = gsum<G>(dp, mask);
        const float alpha = (j[u] >= 0) ? expf(sp - m) * inv : 0.f;
        const float ds = alpha * (dp - Di);
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
#pragma unroll
          for (int o = WG / 2; o > 0; o >>= 1) bits[ch] |= __shfl_xor_sync(mask, bits[ch], o);
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const float t = hj[u][ch].v[i] + hi[ch].v[i];
            gi[ch].v[i] = fmaf(ds * av[ch].v[i], t > 0.f ? 1.f : slope, gi[ch].v[i]);
            ga[ch].v[i] = fmaf(ds, lrelu(t, slope), ga[ch].v[i]);
          }
        }
        const float ea = is_src ? -alpha : alpha;      // -0.0f keeps the sign for alpha == 0
        if (cw <= 2) {
          // c <= 64: the whole record is one 16-byte store by the group's first lane
          unsigned w0 = bits[0], w1 = 0u;
          if (CH == 1) { if (G * VEC > 32) w1 = __shfl_sync(mask, bits[0], (lane / G) * G + (G > LPW ? LPW : 0)); }
          else w1 = bits[CH > 1 ? 1 : 0];
          if (j[u] >= 0 && lane_g == 0)
            *reinterpret_cast<uint4*>(erec + (long long)pos[u] * 4) =
                make_uint4(__float_as_uint(ea), __float_as_uint(ds), w0, w1);
        } else {
          if (j[u] >= 0 && lane_g == 0)
            *reinterpret_cast<float2*>(erec + (long long)pos[u] * 2) = make_float2(ea, ds);
#pragma unroll
          for (int ch = 0; ch < CH; ++ch) {
            const int w = ((ch * G + lane_g) * VEC) >> 5;
            if (j[u] >= 0 && (lane_g % WG) == 0 && w < cw) emask[(long long)pos[u] * cw + w] = bits[ch];
          }
        }
      }
    }
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      int c0 = (ch * G + lane_g) * VEC;
      st_chunk<VEC>((is_src ? gHs : gHt) + row * c + c0, gi[ch], cok[ch]);
      if (cok[ch]) {
        float* slot = s_ga + (size_t)(threadIdx.x / G) * 2 * c + (is_src ? 0 : c) + c0;
#pragma unroll
        for (int i = 0; i < VEC; ++i)
          if (c0 + i < c) slot[i] = ga[ch].v[i];
      }
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * c; t += blockDim.x) {
    float acc = 0.f;
    for (int gidx = 0; gidx < GROUPS; ++gidx) acc += s_ga[(size_t)gidx * 2 * c + t];
    ga_part[(long long)blockIdx.x * 2 * c + t] = acc;
  }
}

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
__global__ void __launch_bounds__(256)
gatv2_bwd_src_kernel(const int* __restrict__ t_rowptr, const int* __restrict__ t_col,
                     const uint8_t* __restrict__ dst_is_src, const float* __restrict__ af_t2s,
                     const float* __restrict__ af_s2t, float slope, long long n, int c, int cw,
                     const unsigned* __restrict__ erec, const unsigned* __restrict__ emask,
                     const float* __restrict__ gout, float* __restrict__ gHs, float* __restrict__ gHt) {
  const int lane_g = threadIdx.x % G;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (row >= n) return;
  Chunk<VEC> as_[CH], at_[CH], gs[CH], gt[CH];
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
    for (int i = 0; i < VEC; ++i) { gs[ch].v[i] = 0.f; gt[ch].v[i] = 0.f; }
  }
  const int beg = __ldg(t_rowptr + row), end = __ldg(t_rowptr + row + 1);
  constexpr int U = 4;
  for (int e = beg; e < end; e += U) {
    int i_dst[U];
    float al[U], ds[U];
    unsigned mk[U][CH];
    Chunk<VEC> go[U][CH];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = e + u < end;
      i_dst[u] = ok ? __ldg(t_col + e + u) : -1;
      if (cw <= 2) {
        const uint4 r4 = ok ? __ldg(reinterpret_cast<const uint4*>(erec) + e + u) : make_uint4(0u, 0u, 0u, 0u);
        al[u] = __uint_as_float(r4.x);
        ds[u] = __uint_as_float(r4.y);
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) mk[u][ch] = wsel[ch] ? r4.w : r4.z;
      } else {
        const float2 r2 = ok ? __ldg(reinterpret_cast<const float2*>(erec) + e + u) : make_float2(0.f, 0.f);
        al[u] = r2.x;
        ds[u] = r2.y;
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) mk[u][ch] = ok ? __ldg(emask + (long long)(e + u) * cw + wsel[ch]) : 0u;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int ch = 0; ch < CH; ++ch)
        go[u][ch] = ld_chunk<VEC>(gout + (long long)(i_dst[u] < 0 ? 0 : i_dst[u]) * c + (ch * G + lane_g) * VEC,
                                  cok[ch] && i_dst[u] >= 0);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool dsrc = signbit(al[u]);
      const float alpha = fabsf(al[u]);
#pragma unroll
      for (int ch = 0; ch < CH; ++ch)
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float aa = dsrc ? as_[ch].v[i] : at_[ch].v[i];
          const float dl = ((mk[u][ch] >> (bsel[ch] + i)) & 1u) ? 1.f : slope;
          const float g = fmaf(ds[u] * aa, dl, alpha * go[u][ch].v[i]);
          gs[ch].v[i] += dsrc ? g : 0.f;
          gt[ch].v[i] += dsrc ? 0.f : g;
        }
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
  if (!pick_row_config(c, vec, g, ch)) return 0;
  const int cw = (c + 31) / 32;
  // records: 16 B per edge when the mask fits two words, else (alpha, dscore) pairs + a mask array
  const size_t rec = cw <= 2 ? (size_t)e * 16 : (size_t)e * 8 + align_up((size_t)e * cw * sizeof(unsigned), 256);
  return align_up((size_t)bwd_blocks(n, g) * 2 * c * sizeof(float), 256) + align_up(rec, 256) + 1024;
}

int launch_gatv2_bwd(const int* rowptr, const int* col, const int* t_rowptr, const int* t_col, const int* csr_to_csc,
                     long long e, const uint8_t* dst_is_src, const float* Hs, const float* Ht, const float* af_t2s,
                     const float* af_s2t, float slope, long long n, int c, const float* out, const float* row_max,
                     const float* row_sum, const float* gout, float* gHs, float* gHt, float* g_af_t2s,
                     float* g_af_s2t, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (n <= 0) return BGNN_OK;
  int vec, g, ch;
  if (!pick_row_config(c, vec, g, ch)) return BGNN_ERR_UNSUPPORTED;
  long long blocks = bwd_blocks(n, g);
  const int cw = (c + 31) / 32;
  Workspace w(ws, ws_bytes);
  float* part = w.take<float>(blocks * 2 * c);
  unsigned* erec = w.take<unsigned>(cw <= 2 ? e * 4 : e * 2);
  unsigned* emask = cw <= 2 ? nullptr : w.take<unsigned>(e * cw);
  if (!w.ok()) return BGNN_ERR_WORKSPACE;
  size_t dyn = (size_t)(256 / g) * 2 * c * sizeof(float);
#define CALL(V, G_, C_)                                                                                            \
  gatv2_bwd_dst_kernel<V, G_, C_><<<(unsigned)blocks, 256, dyn, stream>>>(rowptr, col, csr_to_csc, dst_is_src, Hs, \
      Ht, af_t2s, af_s2t, slope, n, c, cw, out, row_max, row_sum, gout, gHs, gHt, erec, emask, part)
  BGNN_ROW_DISPATCH(vec, g, ch, CALL);
#undef CALL
  BGNN_LAUNCH_CHECK();
  reduce_partials_kernel<<<2 * c, 256, 0, stream>>>(part, blocks, 2 * c, g_af_t2s, g_af_s2t, c);
  BGNN_LAUNCH_CHECK();
#define CALL(V, G_, C_)                                                                                            \
  gatv2_bwd_src_kernel<V, G_, C_><<<(unsigned)blocks, 256, 0, stream>>>(t_rowptr, t_col, dst_is_src, af_t2s,      \
      af_s2t, slope, n, c, cw, erec, emask, gout, gHs, gHt)
  BGNN_ROW_DISPATCH(vec, g, ch, CALL);
#undef CALL
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

}  // namespace bgnn
