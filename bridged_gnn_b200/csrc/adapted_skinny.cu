// AdaptedConv node-wise transform for NARROW outputs (classifier heads: C <= 4 classes), fused, fp32, HBM-bound.
//
// Reference: models/KTGNN.py:275-284 -- gates from [x, Delta], the two shifted copies of x, lin_s / lin_t.
// For wide outputs the host runs ONE dense contraction P = x Wcat^T (cuBLAS) followed by
// adapted_transform.cu.  With C <= 4 that contraction has 2C+2 <= 10 output columns: a GEMM library tiles it
// for 64+ columns and spends ~10x the time the bytes need, and P, dP are extra round trips.  Here every row of
// x is read once per direction and nothing but (Hs, Ht, gates) / (dx, parameter gradients) is written:
//
//   forward   p = Wcat x_i + b            Wcat = [W_s; W_t; a_g_s2t[:D]; a_g_t2s[:D]]   (O = 2C+2 rows)
//             g0 = tanh(p[2C] + kg0), g1 = tanh(p[2C+1] + kg1)
//             Hs[i] = p[0:C] + (1-c_i) g1 wd[0:C],   Ht[i] = p[C:2C] - c_i g0 wd[C:2C]
//   backward  dp[0:C] = dHs[i], dp[C:2C] = dHt[i], dp[2C] = -c_i (dHt[i].wd_t)(1-g0^2), dp[2C+1] = (1-c_i)(dHs[i].wd_s)(1-g1^2)
//             dx[i] = Wcat^T dp,   dWcat += dp x_i^T,   colsum += dp  (= d bias, d kg),
//             d wd_s += (1-c_i) g1 dHs[i],   d wd_t -= c_i g0 dHt[i]
//
// A group of G lanes owns a row (128-bit slices of x); the O dot products are reduced with a butterfly.  The
// backward keeps dWcat slices in registers over a persistent row loop and reduces them per CTA in shared
// memory, then across CTAs with reduce_columns (fixed order: deterministic).
//
// domain_colsum_kernel: sums[0] = sum of source-domain rows of x, sums[1] = target-domain rows (the Delta of
// :275-276 is their scaled difference); replaces the [2,N] x [N,D] GEMM the host used for the two means.
#include "common.cuh"
#include "kernels.h"

namespace bgnn {

constexpr int SK_THREADS = 256;
constexpr int SK_CTAS = kNumSMs * 4;

template <int G>
__device__ __forceinline__ unsigned sk_group_mask(int lane) {
  return G == 32 ? 0xffffffffu : (((1u << G) - 1u) << ((lane / G) * G));
}

// Transposed reduction: V values per lane, G lanes per row.  Each step halves both the lane distance and the number of
// values a lane still carries (the upper lanes keep the upper half), so V - 1 + log2(G / V) shuffles do what a
// butterfly per value does with V log2(G); afterwards lane l of the group holds the complete sum of value l / (G / V).
template <int S, int CNT>
__device__ __forceinline__ void sk_tr_step(float* a, int lane_g, unsigned mask) {
  if constexpr (S >= 1) {
    if constexpr (CNT > 1) {
      constexpr int H = CNT / 2;
      const bool upper = (lane_g & S) != 0;
#pragma unroll
      for (int i = 0; i < H; ++i) {
        const float send = upper ? a[i] : a[i + H];
        const float keep = upper ? a[i + H] : a[i];
        a[i] = keep + __shfl_xor_sync(mask, send, S);
      }
      sk_tr_step<S / 2, H>(a, lane_g, mask);
    } else {
      a[0] += __shfl_xor_sync(mask, a[0], S);
      sk_tr_step<S / 2, 1>(a, lane_g, mask);
    }
  }
}

__host__ __device__ constexpr int sk_pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// HEADS convs that read the SAME x (classifier heads, models/KTGNN.py:432-434) are served by one pass: their weight
// rows are stacked (wcat [HEADS*O, d], bias [HEADS*O], wd [HEADS*2C], kg [HEADS*2]) and the outputs lie side by side
// (Hs, Ht [n, HEADS*C], gates [n, HEADS*2]) -- the layout the multi-head aggregation kernel takes.
template <int C, int G, int CHD, int HEADS>
__global__ void __launch_bounds__(SK_THREADS)
adapted_skinny_fwd_kernel(const float* __restrict__ x, const uint8_t* __restrict__ is_src, const float* __restrict__ wcat,
                          const float* __restrict__ bias, const float* __restrict__ wd, const float* __restrict__ kg,
                          long long n, int d, float* __restrict__ Hs, float* __restrict__ Ht, float* __restrict__ gates) {
  constexpr int O = 2 * C + 2;
  constexpr int OT = HEADS * O;
  extern __shared__ __align__(16) float s_w[];     // [OT][d]
  for (int t = threadIdx.x; t < OT * d; t += blockDim.x) s_w[t] = __ldg(wcat + t);
  __syncthreads();
  const int lane = threadIdx.x & 31, lane_g = threadIdx.x % G;
  const unsigned mask = sk_group_mask<G>(lane);
  const long long groups = (long long)gridDim.x * (SK_THREADS / G);
  // the next row's slices are in flight while this row is reduced (one 16-byte load per thread and row would
  // otherwise leave the kernel latency-bound at ~1/3 of the HBM rate)
  long long row = (long long)blockIdx.x * (SK_THREADS / G) + threadIdx.x / G;
  float4 xn[CHD];
#pragma unroll
  for (int k = 0; k < CHD; ++k) {
    const int c0 = (lane_g + k * G) * 4;
    xn[k] = (row < n && c0 < d) ? __ldg(reinterpret_cast<const float4*>(x + row * d + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (; row < n; row += groups) {
    float4 xc[CHD];
#pragma unroll
    for (int k = 0; k < CHD; ++k) {
      const int c0 = (lane_g + k * G) * 4;
      xc[k] = xn[k];
      if (row + groups < n && c0 < d) xn[k] = __ldg(reinterpret_cast<const float4*>(x + (row + groups) * d + c0));
    }
    constexpr int V = sk_pow2_ceil(OT);
    float acc[V];
#pragma unroll
    for (int o = 0; o < V; ++o) acc[o] = 0.f;
#pragma unroll
    for (int k = 0; k < CHD; ++k) {
      const int c0 = (lane_g + k * G) * 4;
      if (c0 < d) {
        const float4 xv = xc[k];
#pragma unroll
        for (int o = 0; o < OT; ++o) {
          const float4 w = *reinterpret_cast<const float4*>(s_w + o * d + c0);
          acc[o] = fmaf(xv.x, w.x, fmaf(xv.y, w.y, fmaf(xv.z, w.z, fmaf(xv.w, w.w, acc[o]))));
        }
      }
    }
    if constexpr (V <= G && G >= 8) {
      // one value per lane after the transposed reduction; the lane that owns an output computes and stores it
      sk_tr_step<G / 2, V>(acc, lane_g, mask);
      constexpr int R = G / V;
      const int v = lane_g / R;                        // index of the value this lane holds
      const int h = v < OT ? v / O : 0, o = v - h * O;
      const int base = lane - lane_g;                  // first lane of the group inside the warp
      const float g0raw = __shfl_sync(mask, acc[0], base + (h * O + 2 * C) * R);
      const float g1raw = __shfl_sync(mask, acc[0], base + (h * O + 2 * C + 1) * R);
      if (v < OT && lane_g % R == 0) {
        const bool src = is_src[row] != 0;
        if (o < C) {
          const float fs = src ? 0.f : tanhf(g1raw + __ldg(kg + 2 * h + 1));
          Hs[(row * HEADS + h) * C + o] = fmaf(fs, __ldg(wd + h * 2 * C + o), acc[0] + (bias ? __ldg(bias + h * O + o) : 0.f));
        } else if (o < 2 * C) {
          const float ft = src ? -tanhf(g0raw + __ldg(kg + 2 * h)) : 0.f;
          Ht[(row * HEADS + h) * C + o - C] = fmaf(ft, __ldg(wd + h * 2 * C + o), acc[0] + (bias ? __ldg(bias + h * O + o) : 0.f));
        } else {
          gates[(row * HEADS + h) * 2 + (o - 2 * C)] = tanhf(acc[0] + __ldg(kg + 2 * h + (o - 2 * C)));
        }
      }
    } else {
      // few lanes per row or more values than lanes: a butterfly per value, lane 0 of the group finishes the row
#pragma unroll
      for (int o = 0; o < OT; ++o)
#pragma unroll
        for (int s = G / 2; s > 0; s >>= 1) acc[o] += __shfl_xor_sync(mask, acc[o], s);
      if (lane_g == 0) {
        const bool src = is_src[row] != 0;
#pragma unroll
        for (int h = 0; h < HEADS; ++h) {
          const float* a = acc + h * O;
          const float g0 = tanhf(a[2 * C] + __ldg(kg + 2 * h)), g1 = tanhf(a[2 * C + 1] + __ldg(kg + 2 * h + 1));
          const float fs = src ? 0.f : g1, ft = src ? -g0 : 0.f;
#pragma unroll
          for (int j = 0; j < C; ++j) {
            Hs[(row * HEADS + h) * C + j] = fmaf(fs, __ldg(wd + h * 2 * C + j), a[j] + (bias ? __ldg(bias + h * O + j) : 0.f));
            Ht[(row * HEADS + h) * C + j] = fmaf(ft, __ldg(wd + h * 2 * C + C + j), a[C + j] + (bias ? __ldg(bias + h * O + C + j) : 0.f));
          }
          gates[(row * HEADS + h) * 2] = g0;
          gates[(row * HEADS + h) * 2 + 1] = g1;
        }
      }
    }
  }
}

// part[cta][ OT*d (dWcat) | OT (column sums of dp) | HEADS*2C (d wd) ],  OT = HEADS * (2C+2).
// gm [2, d] (or null): added to every source-domain (row 0) / target-domain (row 1) row of dx -- the gradient that
// reaches x through the two domain means (Delta feeds wd and kg), folded in here instead of a separate [n, d] pass.
template <int C, int G, int CHD, int HEADS>
__global__ void __launch_bounds__(SK_THREADS, 2)
adapted_skinny_bwd_kernel(const float* __restrict__ x, const uint8_t* __restrict__ is_src, const float* __restrict__ wcat,
                          const float* __restrict__ wd, const float* __restrict__ gates, const float* __restrict__ gHs,
                          const float* __restrict__ gHt, const float* __restrict__ gm, long long n, int d,
                          float* __restrict__ gx, float* __restrict__ part) {
  constexpr int O = 2 * C + 2;
  constexpr int OT = HEADS * O;
  constexpr int GROUPS = SK_THREADS / G;
  extern __shared__ __align__(16) float s_mem[];   // [OT][d] weights, then [GROUPS][width] partials
  float* s_w = s_mem;
  const int width = OT * d + OT + HEADS * 2 * C;
  const int wpad = (width + 3) & ~3;               // 16-byte aligned partial rows (128-bit stores below)
  float* s_red = s_mem + OT * d;
  for (int t = threadIdx.x; t < OT * d; t += blockDim.x) s_w[t] = __ldg(wcat + t);
  __syncthreads();
  const int lane_g = threadIdx.x % G, grp = threadIdx.x / G;
  float4 gw[OT][CHD];
  float colsum[OT], gwd[HEADS * 2 * C];
#pragma unroll
  for (int o = 0; o < OT; ++o) {
    colsum[o] = 0.f;
#pragma unroll
    for (int k = 0; k < CHD; ++k) gw[o][k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int j = 0; j < HEADS * 2 * C; ++j) gwd[j] = 0.f;
  float wdr[HEADS * 2 * C];
#pragma unroll
  for (int j = 0; j < HEADS * 2 * C; ++j) wdr[j] = __ldg(wd + j);
  const long long groups = (long long)gridDim.x * GROUPS;
  long long row = (long long)blockIdx.x * GROUPS + grp;
  float4 xn[CHD];                                   // next row's slices, in flight while this row is processed
#pragma unroll
  for (int k = 0; k < CHD; ++k) {
    const int c0 = (lane_g + k * G) * 4;
    xn[k] = (row < n && c0 < d) ? __ldg(reinterpret_cast<const float4*>(x + row * d + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (; row < n; row += groups) {
    float4 xc[CHD];
#pragma unroll
    for (int k = 0; k < CHD; ++k) {
      const int c0 = (lane_g + k * G) * 4;
      xc[k] = xn[k];
      if (row + groups < n && c0 < d) xn[k] = __ldg(reinterpret_cast<const float4*>(x + (row + groups) * d + c0));
    }
    const bool src = is_src[row] != 0;
    float dp[OT];
#pragma unroll
    for (int h = 0; h < HEADS; ++h) {
      const float g0 = __ldg(gates + (row * HEADS + h) * 2), g1 = __ldg(gates + (row * HEADS + h) * 2 + 1);
      const float fs = src ? 0.f : g1, ft = src ? -g0 : 0.f;
      float ds = 0.f, dt = 0.f;
#pragma unroll
      for (int j = 0; j < C; ++j) {
        const float a = __ldg(gHs + (row * HEADS + h) * C + j), b = __ldg(gHt + (row * HEADS + h) * C + j);
        dp[h * O + j] = a;
        dp[h * O + C + j] = b;
        ds = fmaf(a, wdr[h * 2 * C + j], ds);
        dt = fmaf(b, wdr[h * 2 * C + C + j], dt);
        gwd[h * 2 * C + j] = fmaf(fs, a, gwd[h * 2 * C + j]);
        gwd[h * 2 * C + C + j] = fmaf(ft, b, gwd[h * 2 * C + C + j]);
      }
      dp[h * O + 2 * C] = (src ? -dt : 0.f) * (1.f - g0 * g0);
      dp[h * O + 2 * C + 1] = (src ? 0.f : ds) * (1.f - g1 * g1);
    }
#pragma unroll
    for (int o = 0; o < OT; ++o) colsum[o] += dp[o];
#pragma unroll
    for (int k = 0; k < CHD; ++k) {
      const int c0 = (lane_g + k * G) * 4;
      if (c0 < d) {
        const float4 xv = xc[k];
        float4 g = gm ? __ldg(reinterpret_cast<const float4*>(gm + (src ? 0 : d) + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int o = 0; o < OT; ++o) {
          const float4 w = *reinterpret_cast<const float4*>(s_w + o * d + c0);
          g.x = fmaf(dp[o], w.x, g.x); g.y = fmaf(dp[o], w.y, g.y); g.z = fmaf(dp[o], w.z, g.z); g.w = fmaf(dp[o], w.w, g.w);
          gw[o][k].x = fmaf(dp[o], xv.x, gw[o][k].x); gw[o][k].y = fmaf(dp[o], xv.y, gw[o][k].y);
          gw[o][k].z = fmaf(dp[o], xv.z, gw[o][k].z); gw[o][k].w = fmaf(dp[o], xv.w, gw[o][k].w);
        }
        *reinterpret_cast<float4*>(gx + row * d + c0) = g;
      }
    }
  }
  // per-CTA reduction over the groups (fixed order), one partial row per CTA
  float* mine = s_red + (size_t)grp * wpad;
#pragma unroll
  for (int o = 0; o < OT; ++o)
#pragma unroll
    for (int k = 0; k < CHD; ++k) {
      const int c0 = (lane_g + k * G) * 4;
      if (c0 < d) *reinterpret_cast<float4*>(mine + o * d + c0) = gw[o][k];
    }
  if (lane_g == 0) {
#pragma unroll
    for (int o = 0; o < OT; ++o) mine[OT * d + o] = colsum[o];
#pragma unroll
    for (int j = 0; j < HEADS * 2 * C; ++j) mine[OT * d + OT + j] = gwd[j];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < width; t += blockDim.x) {
    float acc = 0.f;
    for (int g = 0; g < GROUPS; ++g) acc += s_red[(size_t)g * wpad + t];
    part[(long long)blockIdx.x * width + t] = acc;
  }
}

// The part of the backward that does not touch x: per head (d wd [2C], d kg [2]) from (gates, dHs, dHt).  It runs
// BEFORE the pass over x so that the gradient through the domain means can ride along in that pass (gm above).
// part[cta][HEADS * (2C+2)]; thread <-> row.
template <int C, int HEADS>
__global__ void __launch_bounds__(SK_THREADS)
adapted_skinny_pre_kernel(const uint8_t* __restrict__ is_src, const float* __restrict__ wd, const float* __restrict__ gates,
                          const float* __restrict__ gHs, const float* __restrict__ gHt, long long n, float* __restrict__ part) {
  constexpr int O = 2 * C + 2;
  constexpr int OT = HEADS * O;
  __shared__ float s_red[SK_THREADS / 32][OT];
  float acc[OT];
#pragma unroll
  for (int o = 0; o < OT; ++o) acc[o] = 0.f;
  float wdr[HEADS * 2 * C];
#pragma unroll
  for (int j = 0; j < HEADS * 2 * C; ++j) wdr[j] = __ldg(wd + j);
  for (long long row = (long long)blockIdx.x * SK_THREADS + threadIdx.x; row < n; row += (long long)gridDim.x * SK_THREADS) {
    const bool src = is_src[row] != 0;
#pragma unroll
    for (int h = 0; h < HEADS; ++h) {
      const float g0 = __ldg(gates + (row * HEADS + h) * 2), g1 = __ldg(gates + (row * HEADS + h) * 2 + 1);
      const float fs = src ? 0.f : g1, ft = src ? -g0 : 0.f;
      float ds = 0.f, dt = 0.f;
#pragma unroll
      for (int j = 0; j < C; ++j) {
        const float a = __ldg(gHs + (row * HEADS + h) * C + j), b = __ldg(gHt + (row * HEADS + h) * C + j);
        ds = fmaf(a, wdr[h * 2 * C + j], ds);
        dt = fmaf(b, wdr[h * 2 * C + C + j], dt);
        acc[h * O + j] = fmaf(fs, a, acc[h * O + j]);
        acc[h * O + C + j] = fmaf(ft, b, acc[h * O + C + j]);
      }
      acc[h * O + 2 * C] += (src ? -dt : 0.f) * (1.f - g0 * g0);
      acc[h * O + 2 * C + 1] += (src ? 0.f : ds) * (1.f - g1 * g1);
    }
  }
#pragma unroll
  for (int o = 0; o < OT; ++o) {
    const float v = warp_sum(acc[o]);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][o] = v;
  }
  __syncthreads();
  if (threadIdx.x < OT) {
    float v = 0.f;
    for (int w = 0; w < SK_THREADS / 32; ++w) v += s_red[w][threadIdx.x];
    part[(long long)blockIdx.x * OT + threadIdx.x] = v;
  }
}

// out[t] = sum_p part[p][t]: strided partial sums then a fixed-shape tree (deterministic).
__global__ void __launch_bounds__(256)
sk_reduce_columns_kernel(const float* __restrict__ part, int nparts, int width, float* __restrict__ out) {
  __shared__ float red[256];
  const int t = blockIdx.x;
  float acc = 0.f;
  for (int p = threadIdx.x; p < nparts; p += 256) acc += part[(long long)p * width + t];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[t] = red[0];
}

static bool sk_config(int c, int d, int& g, int& chd) {
  if (c < 1 || c > 4 || d < 4 || d % 4 != 0 || d > 256) return false;
  const int units = d / 4;
  g = 1;
  while (g < 32 && g < units) g <<= 1;
  chd = (units + g - 1) / g;
  return chd <= 2;
}

bool adapted_skinny_supported(int c, int d) {
  int g, chd;
  return sk_config(c, d, g, chd);
}

#define SK_DISPATCH_G(C_, H_, CALL)                                        \
  switch (g) {                                                             \
    case 1: CALL(C_, 1, 1, H_); break;                                     \
    case 2: CALL(C_, 2, 1, H_); break;                                     \
    case 4: CALL(C_, 4, 1, H_); break;                                     \
    case 8: CALL(C_, 8, 1, H_); break;                                     \
    case 16: CALL(C_, 16, 1, H_); break;                                   \
    default: if (chd == 1) { CALL(C_, 32, 1, H_); } else { CALL(C_, 32, 2, H_); } break; \
  }
#define SK_DISPATCH_C(H_, CALL)                                            \
  switch (c) {                                                             \
    case 1: SK_DISPATCH_G(1, H_, CALL); break;                             \
    case 2: SK_DISPATCH_G(2, H_, CALL); break;                             \
    case 3: SK_DISPATCH_G(3, H_, CALL); break;                             \
    default: SK_DISPATCH_G(4, H_, CALL); break;                            \
  }
#define SK_DISPATCH(CALL)                                                  \
  if (heads == 1) { SK_DISPATCH_C(1, CALL); } else { SK_DISPATCH_C(2, CALL); }

bool adapted_skinny_heads_supported(int c, int d, int heads) {
  int g, chd;
  return (heads == 1 || heads == 2) && sk_config(c, d, g, chd);
}

int launch_adapted_skinny_fwd(const float* x, const uint8_t* is_src, const float* wcat, const float* bias, const float* wd,
                              const float* kg, long long n, int d, int c, int heads, float* Hs, float* Ht, float* gates,
                              cudaStream_t stream) {
  if (n <= 0) return BGNN_OK;
  int g, chd;
  if (!adapted_skinny_heads_supported(c, d, heads) || !sk_config(c, d, g, chd)) return BGNN_ERR_UNSUPPORTED;
  const size_t dyn = (size_t)heads * (2 * c + 2) * d * sizeof(float);
  long long ctas = (n * g + SK_THREADS - 1) / SK_THREADS;
  if (ctas > SK_CTAS) ctas = SK_CTAS;
#define CALL(C_, G_, K_, H_)                                                                                          \
  adapted_skinny_fwd_kernel<C_, G_, K_, H_><<<(unsigned)ctas, SK_THREADS, dyn, stream>>>(x, is_src, wcat, bias, wd, kg, n, \
                                                                                        d, Hs, Ht, gates)
  SK_DISPATCH(CALL);
#undef CALL
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

static int sk_bwd_width(int c, int d, int heads) { return heads * ((2 * c + 2) * d + (2 * c + 2) + 2 * c); }

size_t adapted_skinny_bwd_workspace_bytes(int c, int d, int heads) {
  if (!adapted_skinny_heads_supported(c, d, heads)) return 0;
  return (size_t)SK_CTAS * sk_bwd_width(c, d, heads) * sizeof(float) + 256;
}

// red [OT*d + OT + heads*2C] = (dWcat row-major, column sums of dp = per head (d bias, d kg), d wd)
int launch_adapted_skinny_bwd(const float* x, const uint8_t* is_src, const float* wcat, const float* wd, const float* gates,
                              const float* gHs, const float* gHt, const float* gm, long long n, int d, int c, int heads,
                              float* gx, float* red, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (n <= 0) return BGNN_OK;
  int g, chd;
  if (!adapted_skinny_heads_supported(c, d, heads) || !sk_config(c, d, g, chd)) return BGNN_ERR_UNSUPPORTED;
  if (ws_bytes < adapted_skinny_bwd_workspace_bytes(c, d, heads)) return BGNN_ERR_WORKSPACE;
  float* part = reinterpret_cast<float*>(ws);
  const int width = sk_bwd_width(c, d, heads);
  const size_t dyn = ((size_t)heads * (2 * c + 2) * d + (size_t)(SK_THREADS / g) * ((width + 3) & ~3)) * sizeof(float);
  if (dyn > 200 * 1024) return BGNN_ERR_UNSUPPORTED;
  long long ctas = (n * g + SK_THREADS - 1) / SK_THREADS;
  if (ctas > SK_CTAS) ctas = SK_CTAS;
#define CALL(C_, G_, K_, H_)                                                                                        \
  do {                                                                                                              \
    auto kern = adapted_skinny_bwd_kernel<C_, G_, K_, H_>;                                                          \
    BGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));               \
    kern<<<(unsigned)ctas, SK_THREADS, dyn, stream>>>(x, is_src, wcat, wd, gates, gHs, gHt, gm, n, d, gx, part);    \
  } while (0)
  SK_DISPATCH(CALL);
#undef CALL
  BGNN_LAUNCH_CHECK();
  sk_reduce_columns_kernel<<<width, 256, 0, stream>>>(part, (int)ctas, width, red);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

size_t adapted_skinny_pre_workspace_bytes(int c, int heads) { return (size_t)SK_CTAS * heads * (2 * c + 2) * sizeof(float) + 256; }

// pre [heads * (2C+2)] = per head (d wd [2C], d kg [2])
int launch_adapted_skinny_pre(const uint8_t* is_src, const float* wd, const float* gates, const float* gHs, const float* gHt,
                              long long n, int c, int heads, float* pre, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (c < 1 || c > 4 || heads < 1 || heads > 2) return BGNN_ERR_UNSUPPORTED;
  if (ws_bytes < adapted_skinny_pre_workspace_bytes(c, heads)) return BGNN_ERR_WORKSPACE;
  float* part = reinterpret_cast<float*>(ws);
  const int width = heads * (2 * c + 2);
  long long ctas = (n + SK_THREADS - 1) / SK_THREADS;
  if (ctas > SK_CTAS) ctas = SK_CTAS;
  if (ctas < 1) ctas = 1;
#define CALL(C_, H_) adapted_skinny_pre_kernel<C_, H_><<<(unsigned)ctas, SK_THREADS, 0, stream>>>(is_src, wd, gates, gHs, gHt, n, part)
#define CALL_C(H_)                    \
  switch (c) {                        \
    case 1: CALL(1, H_); break;       \
    case 2: CALL(2, H_); break;       \
    case 3: CALL(3, H_); break;       \
    default: CALL(4, H_); break;      \
  }
  if (heads == 1) { CALL_C(1); } else { CALL_C(2); }
#undef CALL_C
#undef CALL
  BGNN_LAUNCH_CHECK();
  sk_reduce_columns_kernel<<<width, 256, 0, stream>>>(part, (int)ctas, width, pre);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

// ---- domain column sums ------------------------------------------------------------------------------------
// thread -> (row slot = tid / (d/4), 128-bit column slice); persistent CTAs stride over the rows.
__global__ void __launch_bounds__(SK_THREADS)
domain_colsum_kernel(const float* __restrict__ x, const uint8_t* __restrict__ is_src, long long n, int d,
                     float* __restrict__ part) {
  extern __shared__ __align__(16) float s_sum[];   // [rows_per_pass][2][d]
  const int units = d / 4;
  const int rpp = SK_THREADS / units;              // rows per pass (>= 1: d <= 1024)
  const int slot = threadIdx.x / units, u = threadIdx.x % units;
  float4 as = make_float4(0.f, 0.f, 0.f, 0.f), at = as;
  if (slot < rpp) {
    for (long long row = (long long)blockIdx.x * rpp + slot; row < n; row += (long long)gridDim.x * rpp) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + row * d) + u);
      if (is_src[row]) { as.x += v.x; as.y += v.y; as.z += v.z; as.w += v.w; }
      else { at.x += v.x; at.y += v.y; at.z += v.z; at.w += v.w; }
    }
    *reinterpret_cast<float4*>(s_sum + ((size_t)slot * 2 + 0) * d + u * 4) = as;
    *reinterpret_cast<float4*>(s_sum + ((size_t)slot * 2 + 1) * d + u * 4) = at;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * d; t += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < rpp; ++s) acc += s_sum[(size_t)s * 2 * d + t];
    part[(long long)blockIdx.x * 2 * d + t] = acc;
  }
}

bool domain_colsum_supported(int d) { return d >= 4 && d % 4 == 0 && d <= 1024; }

size_t domain_colsum_workspace_bytes(int d) { return (size_t)SK_CTAS * 2 * d * sizeof(float) + 256; }

int launch_domain_colsum(const float* x, const uint8_t* is_src, long long n, int d, float* sums, void* ws, size_t ws_bytes,
                         cudaStream_t stream) {
  if (!domain_colsum_supported(d)) return BGNN_ERR_UNSUPPORTED;
  if (ws_bytes < domain_colsum_workspace_bytes(d)) return BGNN_ERR_WORKSPACE;
  float* part = reinterpret_cast<float*>(ws);
  const int rpp = SK_THREADS / (d / 4);
  long long ctas = n > 0 ? (n + rpp - 1) / rpp : 1;
  if (ctas > SK_CTAS) ctas = SK_CTAS;
  const size_t dyn = (size_t)rpp * 2 * d * sizeof(float);
  BGNN_CUDA_TRY(cudaFuncSetAttribute(domain_colsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  domain_colsum_kernel<<<(unsigned)ctas, SK_THREADS, dyn, stream>>>(x, is_src, n, d, part);
  BGNN_LAUNCH_CHECK();
  sk_reduce_columns_kernel<<<2 * d, 256, 0, stream>>>(part, (int)ctas, 2 * d, sums);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

}  // namespace bgnn
