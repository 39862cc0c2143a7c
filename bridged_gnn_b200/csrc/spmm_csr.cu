// CSR gather / segmented-reduce SpMM in fp32 (no tensor cores; HBM-bound).
//
//   Y[i,:] = out_scale[i] * (1/deg_i if mean) * sum_{e in row i} edge_w[e] * gather_scale[col[e]] * X[col[e],:]
//
// Replaces torch_sparse.matmul(adj_t, x, reduce='mean'|'sum') under SAGEConv
// (models/backbones.py:464-468, models/models.py:250-253) and GCNConv's weighted gather/scatter-add
// (models/backbones.py:272-274).  The backward pass is the same kernel on the transposed CSR with the
// per-row factors moved to the gather side.
//
// A group of G lanes owns one destination row and walks its edges 4 at a time with 128-bit feature
// loads; no atomics, fixed summation order (CSR order) -> deterministic.
#include "kernels.h"
#include "rowvec.cuh"

namespace bgnn {

template <int VEC, int G, int CH>
__global__ void __launch_bounds__(256)
spmm_csr_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const float* __restrict__ edge_w,
                const float* __restrict__ gather_scale, const float* __restrict__ out_scale,
                const float* __restrict__ X, long long ldx, long long n_rows, int f, int reduce_mean, float* __restrict__ Y,
                long long ldy) {
  const int lane_g = threadIdx.x % G;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (row >= n_rows) return;   // whole groups exit together
  const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
  Chunk<VEC> acc[CH];
  bool cok[CH];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) {
    cok[ch] = (ch * G + lane_g) * VEC < f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[ch].v[i] = 0.f;
  }
  constexpr int U = 4;
  for (int e = beg; e < end; e += U) {
    int j[U];
    float w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      bool ok = e + u < end;
      j[u] = ok ? __ldg(col + e + u) : -1;
      w[u] = (ok && edge_w) ? __ldg(edge_w + e + u) : 1.f;
    }
    Chunk<VEC> x[U][CH];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (gather_scale && j[u] >= 0) w[u] *= __ldg(gather_scale + j[u]);
#pragma unroll
      for (int ch = 0; ch < CH; ++ch)
        x[u][ch] = ld_chunk<VEC>(X + (long long)(j[u] < 0 ? 0 : j[u]) * ldx + (ch * G + lane_g) * VEC, cok[ch] && j[u] >= 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int ch = 0; ch < CH; ++ch)
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[ch].v[i] = fmaf(w[u], x[u][ch].v[i], acc[ch].v[i]);
  }
  float s = out_scale ? __ldg(out_scale + row) : 1.f;
  if (reduce_mean) s /= (float)max(end - beg, 1);
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[ch].v[i] *= s;
    st_chunk<VEC>(Y + row * ldy + (ch * G + lane_g) * VEC, acc[ch], cok[ch]);
  }
}

// Any width: one warp per row, 32 columns at a time.
__global__ void __launch_bounds__(256)
spmm_csr_generic_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const float* __restrict__ edge_w,
                        const float* __restrict__ gather_scale, const float* __restrict__ out_scale,
                        const float* __restrict__ X, long long ldx, long long n_rows, int f, int reduce_mean,
                        float* __restrict__ Y, long long ldy) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
  float s = out_scale ? __ldg(out_scale + row) : 1.f;
  if (reduce_mean) s /= (float)max(end - beg, 1);
  for (int c0 = 0; c0 < f; c0 += 32) {
    int c = c0 + lane;
    float acc = 0.f;
    for (int e = beg; e < end; ++e) {
      int j = __ldg(col + e);
      float w = edge_w ? __ldg(edge_w + e) : 1.f;
      if (gather_scale) w *= __ldg(gather_scale + j);
      if (c < f) acc = fmaf(w, __ldg(X + (long long)j * ldx + c), acc);
    }
    if (c < f) Y[row * ldy + c] = acc * s;
  }
}

int launch_spmm_csr(const int* rowptr, const int* col, const float* edge_w, const float* gather_scale,
                    const float* out_scale, const float* X, long long ldx, long long n_rows, int f, int reduce_mean, float* Y,
                    long long ldy, cudaStream_t stream) {
  if (n_rows <= 0 || f <= 0) return BGNN_OK;
  int vec, g, ch;
  // 128-bit loads need 16-byte aligned rows: bases and strides (a column panel of a wider matrix qualifies)
  const bool aligned = ldx % 4 == 0 && ldy % 4 == 0 && ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Y)) & 15) == 0;
  if (!pick_row_config(f, vec, g, ch) || (vec == 4 && !aligned)) {
    long long blocks = (n_rows * 32 + 255) / 256;
    spmm_csr_generic_kernel<<<(unsigned)blocks, 256, 0, stream>>>(rowptr, col, edge_w, gather_scale, out_scale, X,
                                                                    ldx, n_rows, f, reduce_mean, Y, ldy);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
  }
  long long blocks = (n_rows * g + 255) / 256;
#define CALL(V, G_, C_)                                                                                          \
  spmm_csr_kernel<V, G_, C_><<<(unsigned)blocks, 256, 0, stream>>>(rowptr, col, edge_w, gather_scale, out_scale, \
                                                                     X, ldx, n_rows, f, reduce_mean, Y, ldy)
  BGNN_ROW_DISPATCH(vec, g, ch, CALL);
#undef CALL
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

}  // namespace bgnn
