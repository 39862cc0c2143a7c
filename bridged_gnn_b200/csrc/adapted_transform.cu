// Node-wise epilogue of AdaptedConv's domain-shift transform (models/KTGNN.py:275-284), fused, fp32, HBM-bound.
//
// The host computes ONE dense contraction P = x [W_s; W_t; a_g_s2t[:D]; a_g_t2s[:D]]^T (+ biases) and
//   wd = [W_s Delta ; W_t Delta],  kg = [a_g_s2t[D:].Delta, a_g_t2s[D:].Delta],  Delta = mean_src(x) - mean_tar(x).
// With c_i = 1 on source-domain rows, the reference's
//   h_s = lin_s(x + tanh(a_g_t2s.[x,Delta]) Delta (1-c)),   h_t = lin_t(x - tanh(a_g_s2t.[x,Delta]) Delta c)
// becomes, per row i (P columns: [0,C) = x W_s^T + b_s, [C,2C) = x W_t^T + b_t, 2C = s2t gate, 2C+1 = t2s gate):
//   g0 = tanh(P[i,2C] + kg0), g1 = tanh(P[i,2C+1] + kg1)
//   Hs[i,:] = P[i,0:C]  + (1-c_i) g1 wd[0:C]
//   Ht[i,:] = P[i,C:2C] -    c_i  g0 wd[C:2C]
// One pass over P instead of ~8 elementwise torch kernels; the backward kernel produces dP, d wd, d kg in one
// pass over (dHs, dHt) with deterministic two-stage column reductions.
#include "common.cuh"
#include "kernels.h"

namespace bgnn {

constexpr int AT_THREADS = 256;

template <int G>
__global__ void __launch_bounds__(AT_THREADS)
adapted_transform_fwd_kernel(const float* __restrict__ P, const uint8_t* __restrict__ is_src, const float* __restrict__ wd,
                             const float* __restrict__ kg, const float* __restrict__ bias, long long n, int c, float* __restrict__ Hs,
                             float* __restrict__ Ht, float* __restrict__ gates) {
  const int lane_g = threadIdx.x % G;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (row >= n) return;
  const int ldp = 2 * c + 2;
  const float* p = P + row * ldp;
  const bool src = is_src[row] != 0;
  const float g0 = tanhf(__ldg(p + 2 * c) + __ldg(kg));
  const float g1 = tanhf(__ldg(p + 2 * c + 1) + __ldg(kg + 1));
  const float fs = src ? 0.f : g1;        // multiplies W_s Delta
  const float ft = src ? -g0 : 0.f;       // multiplies W_t Delta
  for (int j = lane_g; j < c; j += G) {
    // bias (optional, [2c] = (b_s, b_t)): folded in here so that the contraction needs no separate bias pass
    Hs[row * c + j] = fmaf(fs, __ldg(wd + j), __ldg(p + j) + (bias ? __ldg(bias + j) : 0.f));
    Ht[row * c + j] = fmaf(ft, __ldg(wd + c + j), __ldg(p + c + j) + (bias ? __ldg(bias + c + j) : 0.f));
  }
  if (lane_g == 0) { gates[row * 2] = g0; gates[row * 2 + 1] = g1; }
}

// dP[i,0:C] = dHs[i], dP[i,C:2C] = dHt[i], dP[i,2C] = d g0 (1-g0^2), dP[i,2C+1] = d g1 (1-g1^2) with
// d g1 = (1-c_i) dHs[i].wd_s, d g0 = -c_i dHt[i].wd_t;   d wd_s = sum_i (1-c_i) g1_i dHs[i],
// d wd_t = -sum_i c_i g0_i dHt[i];   d kg = column sums of dP[:, 2C:2C+2].
// Persistent grid: each group strides over rows and keeps column partials in registers; partials are
// combined per CTA in shared memory, written to part[cta][2C+2] and reduced by reduce_columns_kernel.
template <int G, int CPL>   // CPL = columns per lane = ceil(C / G)
__global__ void __launch_bounds__(AT_THREADS)
adapted_transform_bwd_kernel(const float* __restrict__ gHs, const float* __restrict__ gHt, const float* __restrict__ gates,
                             const uint8_t* __restrict__ is_src, const float* __restrict__ wd, long long n, int c, int ldp,
                             int copy, float* __restrict__ gP, float* __restrict__ part) {
  extern __shared__ float s_part[];       // [groups][4C+2]: d wd (2C), d kg (2), column sums of dHs, dHt = d bias (2C)
  constexpr int GROUPS = AT_THREADS / G;
  const int lane_g = threadIdx.x % G, grp = threadIdx.x / G;
  const int lpart = 4 * c + 2;
  float ws[CPL], wt[CPL], as_[CPL], at_[CPL], bs_[CPL], bt_[CPL];
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    const int j = lane_g + k * G;
    ws[k] = j < c ? __ldg(wd + j) : 0.f;
    wt[k] = j < c ? __ldg(wd + c + j) : 0.f;
    as_[k] = 0.f; at_[k] = 0.f; bs_[k] = 0.f; bt_[k] = 0.f;
  }
  float k0 = 0.f, k1 = 0.f;
  const long long stride = (long long)gridDim.x * GROUPS;
  const long long iters = (n + stride - 1) / stride;      // uniform trip count: the shuffles below use the full mask
  // the next row's slices are in flight while this row is reduced (4-byte loads: latency-bound otherwise)
  float an[CPL], bn[CPL];
  {
    const long long r0 = (long long)blockIdx.x * GROUPS + grp;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int j = lane_g + k * G;
      const bool ok = r0 < n && j < c;
      an[k] = ok ? __ldg(gHs + r0 * c + j) : 0.f;
      bn[k] = ok ? __ldg(gHt + r0 * c + j) : 0.f;
    }
  }
  for (long long it = 0; it < iters; ++it) {
    const long long row_raw = (long long)blockIdx.x * GROUPS + grp + it * stride;
    const bool valid = row_raw < n;
    const long long row = valid ? row_raw : 0;
    float av[CPL], bv[CPL];
    {
      const long long rn = row_raw + stride;
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const int j = lane_g + k * G;
        av[k] = an[k];
        bv[k] = bn[k];
        const bool ok = rn < n && j < c;
        an[k] = ok ? __ldg(gHs + rn * c + j) : 0.f;
        bn[k] = ok ? __ldg(gHt + rn * c + j) : 0.f;
      }
    }
    const bool src = is_src[row] != 0;
    const float g0 = __ldg(gates + row * 2), g1 = __ldg(gates + row * 2 + 1);
    const float fs = src ? 0.f : g1, ft = src ? -g0 : 0.f;
    float ds = 0.f, dt = 0.f;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int j = lane_g + k * G;
      const float a = av[k], b = bv[k];        // zero outside the matrix
      if (copy && valid && j < c) {
        gP[row * ldp + j] = a;
        gP[row * ldp + c + j] = b;
      }
      ds = fmaf(a, ws[k], ds);
      dt = fmaf(b, wt[k], dt);
      if (valid) {
        as_[k] = fmaf(fs, a, as_[k]);
        at_[k] = fmaf(ft, b, at_[k]);
        bs_[k] += a;
        bt_[k] += b;
      }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      ds += __shfl_xor_sync(0xffffffffu, ds, o);
      dt += __shfl_xor_sync(0xffffffffu, dt, o);
    }
    if (valid && lane_g == 0) {
      const float d0 = (src ? -dt : 0.f) * (1.f - g0 * g0);
      const float d1 = (src ? 0.f : ds) * (1.f - g1 * g1);
      const int goff = copy ? 2 * c : 0;       // gates only: gP is a compact [n, ldp] buffer with (d0, d1) in front
      gP[row * ldp + goff] = d0;
      gP[row * ldp + goff + 1] = d1;
      k0 += d0; k1 += d1;
    }
  }
  float* mine = s_part + (size_t)grp * lpart;
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    const int j = lane_g + k * G;
    if (j < c) { mine[j] = as_[k]; mine[c + j] = at_[k]; mine[2 * c + 2 + j] = bs_[k]; mine[3 * c + 2 + j] = bt_[k]; }
  }
  if (lane_g == 0) { mine[2 * c] = k0; mine[2 * c + 1] = k1; }
  __syncthreads();
  for (int t = threadIdx.x; t < lpart; t += blockDim.x) {
    float acc = 0.f;
    for (int g = 0; g < GROUPS; ++g) acc += s_part[(size_t)g * lpart + t];
    part[(long long)blockIdx.x * lpart + t] = acc;
  }
}

// out[t] = sum_p part[p][t]: one CTA per column, strided partial sums then a fixed-shape tree (deterministic).
__global__ void __launch_bounds__(256)
reduce_columns_kernel(const float* __restrict__ part, int nparts, int width, float* __restrict__ out) {
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

static int pick_g(int c) {
  int g = 1;
  while (g < 32 && g < c) g <<= 1;
  return g;
}

int launch_adapted_transform_fwd(const float* P, const uint8_t* is_src, const float* wd, const float* kg, const float* bias,
                                 long long n, int c, float* Hs, float* Ht, float* gates, cudaStream_t stream) {
  if (n <= 0) return BGNN_OK;
  const int g = pick_g(c);
  const long long blocks = (n * g + AT_THREADS - 1) / AT_THREADS;
#define CALL(G_) adapted_transform_fwd_kernel<G_><<<(unsigned)blocks, AT_THREADS, 0, stream>>>(P, is_src, wd, kg, bias, n, c, Hs, Ht, gates)
  switch (g) {
    case 1: CALL(1); break;
    case 2: CALL(2); break;
    case 4: CALL(4); break;
    case 8: CALL(8); break;
    case 16: CALL(16); break;
    default: CALL(32); break;
  }
#undef CALL
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

constexpr int AT_BWD_CTAS = kNumSMs * 4;

size_t adapted_transform_bwd_workspace_bytes(int c) { return (size_t)AT_BWD_CTAS * (4 * c + 2) * sizeof(float) + 256; }

int launch_adapted_transform_bwd(const float* gHs, const float* gHt, const float* gates, const uint8_t* is_src,
                                 const float* wd, long long n, int c, int ldp, int copy, float* gP, float* g_wd_kg,
                                 void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (n <= 0) return BGNN_OK;
  if (c > 512) return BGNN_ERR_UNSUPPORTED;
  if (ldp < (copy ? 2 * c + 2 : 2)) return BGNN_ERR_INVALID_ARG;
  if (ws_bytes < adapted_transform_bwd_workspace_bytes(c)) return BGNN_ERR_WORKSPACE;
  float* part = reinterpret_cast<float*>(ws);
  const int g = pick_g(c);
  const int cpl = (c + g - 1) / g;
  const int lpart = 4 * c + 2;
  const size_t dyn = (size_t)(AT_THREADS / g) * lpart * sizeof(float);
  if (dyn > 200 * 1024) return BGNN_ERR_UNSUPPORTED;
#define CALL(G_, CPL_)                                                                                              \
  do {                                                                                                              \
    auto kern = adapted_transform_bwd_kernel<G_, CPL_>;                                                             \
    BGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));               \
    kern<<<AT_BWD_CTAS, AT_THREADS, dyn, stream>>>(gHs, gHt, gates, is_src, wd, n, c, ldp, copy, gP, part);                    \
  } while (0)
  if (g < 32) {
    switch (g) {
      case 1: CALL(1, 1); break;
      case 2: CALL(2, 1); break;
      case 4: CALL(4, 1); break;
      case 8: CALL(8, 1); break;
      default: CALL(16, 1); break;
    }
  } else {
    if (cpl <= 1) CALL(32, 1);
    else if (cpl <= 2) CALL(32, 2);
    else if (cpl <= 4) CALL(32, 4);
    else if (cpl <= 8) CALL(32, 8);
    else CALL(32, 16);
  }
#undef CALL
  BGNN_LAUNCH_CHECK();
  reduce_columns_kernel<<<lpart, 256, 0, stream>>>(part, AT_BWD_CTAS, lpart, g_wd_kg);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

}  // namespace bgnn
