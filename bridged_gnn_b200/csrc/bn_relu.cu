// BatchNorm1d (+ ReLU) over the node-feature matrix x [n, c], training mode, fp32, HBM-bound.
//
// KT-GNN normalises every hidden layer and the classifier transformer with BatchNorm1d followed by ReLU
// (models/KTGNN.py:363-366, 425-429: nn.BatchNorm1d + ReLU).  ATen runs that as 4 + 2 passes over
// [n, c] forward and 3 + 1 backward (statistics, transform, clamp; reduce, elementwise, threshold); here it is
// 2 passes each way with the ReLU folded in:
//   forward   pass 1  column sums of (x - K) and (x - K)^2 (K = first row: shifted sums, no cancellation), per-CTA
//                     partials combined in double by one CTA per column -> mean, invstd, running statistics
//             pass 2  y = max((x - mean) * (w * invstd) + b, 0)
//   backward  pass 1  g = dy * [y > 0] (the mask is recomputed from x with the forward's exact expression),
//                     column sums of g and g * xhat  -> d bias, d weight
//             pass 2  dx = (g - mean(g) - xhat * mean(g * xhat)) * w * invstd
// All reductions have a fixed shape: results are bit-reproducible.
#include "common.cuh"
#include "kernels.h"

namespace bgnn {

constexpr int BN_THREADS = 256;
constexpr int BN_CTAS = kNumSMs * 8;

struct BnGeom {
  int tpr;      // threads per row = c / 4
  int rows;     // rows per CTA pass = BN_THREADS / tpr
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// MODE 0: sums of (x - K), (x - K)^2.   MODE 1: sums of g, g * xhat.
template <int MODE>
__global__ void __launch_bounds__(BN_THREADS)
bn_reduce_kernel(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ stats, int relu,
                 long long n, int c, int tpr, int rows, float* __restrict__ part) {
  extern __shared__ float s_red[];           // [rows][2c]
  const int t = threadIdx.x;
  const bool active = t < tpr * rows;
  const int col = (t % tpr) * 4, rl = t / tpr;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
  if (active) {
    float4 k0, k1, k2, k3;                   // MODE 0: K.  MODE 1: mean, invstd, scale, bias
    if (MODE == 0) {
      k0 = ld4(x + col);
    } else {
      k0 = ld4(stats + col); k1 = ld4(stats + c + col); k2 = ld4(stats + 2 * c + col); k3 = ld4(stats + 3 * c + col);
    }
    const long long stride = (long long)gridDim.x * rows;
    for (long long r = (long long)blockIdx.x * rows + rl; r < n; r += stride) {
      const float4 v = ld4(x + r * c + col);
      if (MODE == 0) {
        const float d0 = v.x - k0.x, d1 = v.y - k0.y, d2 = v.z - k0.z, d3 = v.w - k0.w;
        a0.x += d0; a0.y += d1; a0.z += d2; a0.w += d3;
        a1.x = fmaf(d0, d0, a1.x); a1.y = fmaf(d1, d1, a1.y); a1.z = fmaf(d2, d2, a1.z); a1.w = fmaf(d3, d3, a1.w);
      } else {
        float4 g = ld4(gy + r * c + col);
        const float c0 = v.x - k0.x, c1 = v.y - k0.y, c2 = v.z - k0.z, c3 = v.w - k0.w;
        if (relu) {
          if (!(fmaf(c0, k2.x, k3.x) > 0.f)) g.x = 0.f;
          if (!(fmaf(c1, k2.y, k3.y) > 0.f)) g.y = 0.f;
          if (!(fmaf(c2, k2.z, k3.z) > 0.f)) g.z = 0.f;
          if (!(fmaf(c3, k2.w, k3.w) > 0.f)) g.w = 0.f;
        }
        a0.x += g.x; a0.y += g.y; a0.z += g.z; a0.w += g.w;
        a1.x = fmaf(g.x, c0 * k1.x, a1.x); a1.y = fmaf(g.y, c1 * k1.y, a1.y);
        a1.z = fmaf(g.z, c2 * k1.z, a1.z); a1.w = fmaf(g.w, c3 * k1.w, a1.w);
      }
    }
    float* mine = s_red + (size_t)rl * 2 * c;
    *reinterpret_cast<float4*>(mine + col) = a0;
    *reinterpret_cast<float4*>(mine + c + col) = a1;
  }
  __syncthreads();
  for (int j = t; j < 2 * c; j += BN_THREADS) {
    float acc = 0.f;
    for (int r = 0; r < rows; ++r) acc += s_red[(size_t)r * 2 * c + j];
    part[(size_t)blockIdx.x * 2 * c + j] = acc;
  }
}

__device__ __forceinline__ double bn_block_sum(double v, double* red) {
  red[threadIdx.x] = v;
  __syncthreads();
  for (int o = BN_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const double r = red[0];
  __syncthreads();
  return r;
}

// one CTA per column.  stats = [mean | invstd | scale = w * invstd | bias], each [c]
__global__ void __launch_bounds__(BN_THREADS)
bn_fwd_finalize_kernel(const float* __restrict__ part, int nparts, const float* __restrict__ x, const float* __restrict__ w,
                       const float* __restrict__ b, long long n, int c, float eps, float momentum, float* __restrict__ stats,
                       float* __restrict__ running_mean, float* __restrict__ running_var) {
  __shared__ double red[BN_THREADS];
  const int j = blockIdx.x;
  double s1 = 0.0, s2 = 0.0;
  for (int p = threadIdx.x; p < nparts; p += BN_THREADS) {
    s1 += (double)part[(size_t)p * 2 * c + j];
    s2 += (double)part[(size_t)p * 2 * c + c + j];
  }
  s1 = bn_block_sum(s1, red);
  s2 = bn_block_sum(s2, red);
  if (threadIdx.x == 0) {
    const double m1 = s1 / (double)n;
    const double mean = (double)x[j] + m1;
    double var = s2 / (double)n - m1 * m1;
    if (var < 0.0) var = 0.0;
    const double invstd = 1.0 / sqrt(var + (double)eps);
    stats[j] = (float)mean;
    stats[c + j] = (float)invstd;
    stats[2 * c + j] = (float)((w ? (double)w[j] : 1.0) * invstd);
    stats[3 * c + j] = b ? b[j] : 0.f;
    if (running_mean) running_mean[j] = (float)((1.0 - (double)momentum) * (double)running_mean[j] + (double)momentum * mean);
    if (running_var) {
      const double unbiased = n > 1 ? var * (double)n / (double)(n - 1) : var;
      running_var[j] = (float)((1.0 - (double)momentum) * (double)running_var[j] + (double)momentum * unbiased);
    }
  }
}

// gwb = [d weight | d bias]; coef = [mean(g) | mean(g * xhat)]
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_finalize_kernel(const float* __restrict__ part, int nparts, long long n, int c, float* __restrict__ gwb,
                       float* __restrict__ coef) {
  __shared__ double red[BN_THREADS];
  const int j = blockIdx.x;
  double s1 = 0.0, s2 = 0.0;
  for (int p = threadIdx.x; p < nparts; p += BN_THREADS) {
    s1 += (double)part[(size_t)p * 2 * c + j];
    s2 += (double)part[(size_t)p * 2 * c + c + j];
  }
  s1 = bn_block_sum(s1, red);
  s2 = bn_block_sum(s2, red);
  if (threadIdx.x == 0) {
    gwb[j] = (float)s2;
    gwb[c + j] = (float)s1;
    coef[j] = (float)(s1 / (double)n);
    coef[c + j] = (float)(s2 / (double)n);
  }
}

// MODE 0: y = [relu]((x - mean) * scale + bias).   MODE 1: dx = (g - coef0 - xhat * coef1) * scale.
template <int MODE>
__global__ void __launch_bounds__(BN_THREADS)
bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ stats,
                const float* __restrict__ coef, int relu, long long n, int c, int tpr, int rows, float* __restrict__ out) {
  const int t = threadIdx.x;
  if (t >= tpr * rows) return;
  const int col = (t % tpr) * 4, rl = t / tpr;
  const float4 mean = ld4(stats + col), scale = ld4(stats + 2 * c + col), bias = ld4(stats + 3 * c + col);
  float4 invstd, m0, m1;
  if (MODE == 1) { invstd = ld4(stats + c + col); m0 = ld4(coef + col); m1 = ld4(coef + c + col); }
  const long long stride = (long long)gridDim.x * rows;
#pragma unroll 2
  for (long long r = (long long)blockIdx.x * rows + rl; r < n; r += stride) {
    const float4 v = ld4(x + r * c + col);
    const float c0 = v.x - mean.x, c1 = v.y - mean.y, c2 = v.z - mean.z, c3 = v.w - mean.w;
    float4 o;
    o.x = fmaf(c0, scale.x, bias.x); o.y = fmaf(c1, scale.y, bias.y); o.z = fmaf(c2, scale.z, bias.z); o.w = fmaf(c3, scale.w, bias.w);
    if (MODE == 0) {
      if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    } else {
      float4 g = ld4(gy + r * c + col);
      if (relu) {
        if (!(o.x > 0.f)) g.x = 0.f;
        if (!(o.y > 0.f)) g.y = 0.f;
        if (!(o.z > 0.f)) g.z = 0.f;
        if (!(o.w > 0.f)) g.w = 0.f;
      }
      o.x = (g.x - m0.x - c0 * invstd.x * m1.x) * scale.x;
      o.y = (g.y - m0.y - c1 * invstd.y * m1.y) * scale.y;
      o.z = (g.z - m0.z - c2 * invstd.z * m1.z) * scale.z;
      o.w = (g.w - m0.w - c3 * invstd.w * m1.w) * scale.w;
    }
    *reinterpret_cast<float4*>(out + r * c + col) = o;
  }
}

// ------------------------------------------------------------------------------------------ host
bool bn_relu_supported(int c) { return c >= 4 && c % 4 == 0 && c <= 1024; }

static BnGeom bn_geom(int c) {
  BnGeom g;
  g.tpr = c / 4;
  g.rows = BN_THREADS / g.tpr;
  return g;
}

static int bn_ctas(long long n, int rows) {
  const long long need = (n + rows - 1) / rows;
  return (int)(need < BN_CTAS ? (need > 0 ? need : 1) : BN_CTAS);
}

size_t bn_relu_workspace_bytes(int c) { return ((size_t)BN_CTAS * 2 * c + 2 * c) * sizeof(float) + 256; }

int launch_bn_relu_fwd(const float* x, long long n, int c, const float* w, const float* b, float eps, float momentum,
                       float* running_mean, float* running_var, int relu, float* y, float* stats, void* ws, size_t ws_bytes,
                       cudaStream_t stream) {
  if (!bn_relu_supported(c)) return BGNN_ERR_UNSUPPORTED;
  if (n <= 0) return BGNN_OK;
  if (ws_bytes < bn_relu_workspace_bytes(c)) return BGNN_ERR_WORKSPACE;
  float* part = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  const BnGeom g = bn_geom(c);
  const int ctas = bn_ctas(n, g.rows);
  const size_t dyn = (size_t)g.rows * 2 * c * sizeof(float);
  BGNN_CUDA_TRY(cudaFuncSetAttribute(bn_reduce_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  bn_reduce_kernel<0><<<ctas, BN_THREADS, dyn, stream>>>(x, nullptr, nullptr, 0, n, c, g.tpr, g.rows, part);
  BGNN_LAUNCH_CHECK();
  bn_fwd_finalize_kernel<<<c, BN_THREADS, 0, stream>>>(part, ctas, x, w, b, n, c, eps, momentum, stats, running_mean, running_var);
  BGNN_LAUNCH_CHECK();
  if (y) {                                            // y == NULL: statistics only (the multi-GPU path combines them first)
    bn_apply_kernel<0><<<ctas, BN_THREADS, 0, stream>>>(x, nullptr, stats, nullptr, relu, n, c, g.tpr, g.rows, y);
    BGNN_LAUNCH_CHECK();
  }
  return BGNN_OK;
}

int launch_bn_relu_apply(const float* x, long long n, int c, const float* stats, int relu, float* y, cudaStream_t stream) {
  if (!bn_relu_supported(c)) return BGNN_ERR_UNSUPPORTED;
  if (n <= 0) return BGNN_OK;
  const BnGeom g = bn_geom(c);
  bn_apply_kernel<0><<<bn_ctas(n, g.rows), BN_THREADS, 0, stream>>>(x, nullptr, stats, nullptr, relu, n, c, g.tpr, g.rows, y);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

int launch_bn_relu_bwd(const float* gy, const float* x, long long n, int c, const float* stats, int relu, float* gx, float* gwb,
                       void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (!bn_relu_supported(c)) return BGNN_ERR_UNSUPPORTED;
  if (ws_bytes < bn_relu_workspace_bytes(c)) return BGNN_ERR_WORKSPACE;
  float* part = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  float* coef = part + (size_t)BN_CTAS * 2 * c;
  const BnGeom g = bn_geom(c);
  const int ctas = n > 0 ? bn_ctas(n, g.rows) : 0;
  if (n > 0) {
    const size_t dyn = (size_t)g.rows * 2 * c * sizeof(float);
    BGNN_CUDA_TRY(cudaFuncSetAttribute(bn_reduce_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    bn_reduce_kernel<1><<<ctas, BN_THREADS, dyn, stream>>>(x, gy, stats, relu, n, c, g.tpr, g.rows, part);
    BGNN_LAUNCH_CHECK();
  }
  bn_bwd_finalize_kernel<<<c, BN_THREADS, 0, stream>>>(part, ctas, n > 0 ? n : 1, c, gwb, coef);
  BGNN_LAUNCH_CHECK();
  if (n > 0) {
    bn_apply_kernel<1><<<ctas, BN_THREADS, 0, stream>>>(x, gy, stats, coef, relu, n, c, g.tpr, g.rows, gx);
    BGNN_LAUNCH_CHECK();
  }
  return BGNN_OK;
}

// The two halves of the backward for the destination-partitioned (multi-GPU) path, where the column sums are
// all-reduced in between: `reduce` gives this rank's sums (d weight | d bias) = (sum g xhat | sum g) under the GLOBAL
// statistics; `apply` takes coef = (mean g | mean g xhat) over ALL ranks' rows.
int launch_bn_relu_bwd_reduce(const float* gy, const float* x, long long n, int c, const float* stats, int relu, float* gwb,
                              void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (!bn_relu_supported(c)) return BGNN_ERR_UNSUPPORTED;
  if (ws_bytes < bn_relu_workspace_bytes(c)) return BGNN_ERR_WORKSPACE;
  float* part = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  float* coef = part + (size_t)BN_CTAS * 2 * c;
  const BnGeom g = bn_geom(c);
  const int ctas = n > 0 ? bn_ctas(n, g.rows) : 0;
  if (n > 0) {
    const size_t dyn = (size_t)g.rows * 2 * c * sizeof(float);
    BGNN_CUDA_TRY(cudaFuncSetAttribute(bn_reduce_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    bn_reduce_kernel<1><<<ctas, BN_THREADS, dyn, stream>>>(x, gy, stats, relu, n, c, g.tpr, g.rows, part);
    BGNN_LAUNCH_CHECK();
  }
  bn_bwd_finalize_kernel<<<c, BN_THREADS, 0, stream>>>(part, ctas, 1, c, gwb, coef);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

int launch_bn_relu_bwd_apply(const float* gy, const float* x, long long n, int c, const float* stats, int relu,
                             const float* coef, float* gx, cudaStream_t stream) {
  if (!bn_relu_supported(c)) return BGNN_ERR_UNSUPPORTED;
  if (n <= 0) return BGNN_OK;
  const BnGeom g = bn_geom(c);
  bn_apply_kernel<1><<<bn_ctas(n, g.rows), BN_THREADS, 0, stream>>>(x, gy, stats, coef, relu, n, c, g.tpr, g.rows, gx);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

}  // namespace bgnn
