// Edge-validity filters of the bridged-graph build on the device.
//
// Replaces check_added_edges_cross_domain_validity / check_added_edges_within_domain_validity
// (main_bridged_graph.py:225-264, 123-161): a quantile of the E edge similarities (torch.quantile = full sort, and
// capped at 16 M elements), then four elementwise rules, the last one a per-edge cosine of RAW feature rows
// (a gather-dot).  Here: the quantile is an 8-bit radix SELECT (4 histogram passes over the similarities + one
// pass for the upper neighbour; no sort, no size cap) and the four rules run in ONE kernel, warp per edge, that
// gathers the two feature rows once and writes a keep flag plus the per-rule removal counts the reference prints.
#include "common.cuh"
#include "kernels.h"

namespace bgnn {

// order-preserving map float -> uint32 (negative values reversed, sign bit flipped for the others)
__device__ __forceinline__ unsigned f2key(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct SelectState {
  unsigned prefix;            // key bits fixed so far (high bits)
  unsigned pad;
  unsigned long long rank;    // rank of the wanted element among the keys that share the prefix
  unsigned long long count_le;
  unsigned min_gt;            // smallest key strictly greater than the selected one (0xffffffff: none)
  unsigned hist[256];
};

__global__ void select_init_kernel(SelectState* st, unsigned long long rank) {
  const int t = threadIdx.x;
  if (t == 0) { st->prefix = 0u; st->rank = rank; st->count_le = 0ull; st->min_gt = 0xffffffffu; }
  if (t < 256) st->hist[t] = 0u;
}

// histogram of byte `shift/8` over the keys whose higher bytes equal the prefix
__global__ void __launch_bounds__(256)
select_hist_kernel(const float* __restrict__ v, long long n, int shift, SelectState* st) {
  __shared__ unsigned sh[256];
  sh[threadIdx.x] = 0u;
  __syncthreads();
  const unsigned prefix = st->prefix;
  const unsigned himask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned k = f2key(__ldg(v + i));
    if ((k & himask) == prefix) atomicAdd(&sh[(k >> shift) & 255u], 1u);
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], sh[threadIdx.x]);
}

// one thread: bucket that holds the wanted rank -> extend the prefix, rebase the rank, clear the histogram
__global__ void select_step_kernel(SelectState* st, int shift) {
  unsigned long long r = st->rank;
  int b = 0;
  for (; b < 255; ++b) {
    const unsigned c = st->hist[b];
    if (r < c) break;
    r -= c;
  }
  st->prefix |= (unsigned)b << shift;
  st->rank = r;
  for (int i = 0; i < 256; ++i) st->hist[i] = 0u;
}

// number of keys <= selected key, smallest key above it
__global__ void __launch_bounds__(256)
select_upper_kernel(const float* __restrict__ v, long long n, SelectState* st) {
  const unsigned sel = st->prefix;
  unsigned long long cnt = 0ull;
  unsigned mn = 0xffffffffu;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned k = f2key(__ldg(v + i));
    if (k <= sel) ++cnt; else mn = min(mn, k);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (cnt) atomicAdd(&st->count_le, cnt);
    if (mn != 0xffffffffu) atomicMin(&st->min_gt, mn);
  }
}

// out = (v_(lo), v_(lo+1), lerp): torch.quantile's linear interpolation (ATen lerp: a + w (b - a) for w < 0.5,
// b - (b - a)(1 - w) otherwise)
__global__ void select_finish_kernel(const SelectState* st, unsigned long long rank_lo, long long n, float weight,
                                     float* __restrict__ out) {
  const float a = key2f(st->prefix);
  float b = a;
  if (rank_lo + 1 < (unsigned long long)n && st->count_le <= rank_lo + 1 && st->min_gt != 0xffffffffu) b = key2f(st->min_gt);
  const float diff = b - a;
  out[0] = a;
  out[1] = b;
  out[2] = weight < 0.5f ? a + weight * diff : b - diff * (1.f - weight);
}

size_t quantile_workspace_bytes() { return align_up(sizeof(SelectState), 256) + 256; }

int launch_quantile(const float* v, long long n, long long rank_lo, float weight, float* out3, void* ws, size_t ws_bytes,
                    cudaStream_t stream) {
  if (n <= 0 || rank_lo < 0 || rank_lo >= n) return BGNN_ERR_INVALID_ARG;
  Workspace w(ws, ws_bytes);
  SelectState* st = w.take<SelectState>(1);
  if (!w.ok()) return BGNN_ERR_WORKSPACE;
  long long blocks = (n + 256 * 8 - 1) / (256 * 8);
  if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
  select_init_kernel<<<1, 256, 0, stream>>>(st, (unsigned long long)rank_lo);
  BGNN_LAUNCH_CHECK();
  for (int shift = 24; shift >= 0; shift -= 8) {
    select_hist_kernel<<<(unsigned)blocks, 256, 0, stream>>>(v, n, shift, st);
    BGNN_LAUNCH_CHECK();
    select_step_kernel<<<1, 1, 0, stream>>>(st, shift);
    BGNN_LAUNCH_CHECK();
  }
  select_upper_kernel<<<(unsigned)blocks, 256, 0, stream>>>(v, n, st);
  BGNN_LAUNCH_CHECK();
  select_finish_kernel<<<1, 1, 0, stream>>>(st, (unsigned long long)rank_lo, n, weight, out3);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

// ---- the four rules, one warp per edge ---------------------------------------------------------------------
//   1. e_sim < thr_conf                                             (main_bridged_graph.py:236-239 / :134-137)
//   2. pred_a[e0] != y_a[e0]  [gated by gate_a[e1] when given]      (:242 / :141)
//      (pred_b[e1] != y_b[e1]) & gate_b[e1]                         (:243 / :142)
//   3. pred_a[e0] != pred_b[e1]                                     (:247 / :146)
//   4. cos(x_a[e0], x_b[e1]) < thres_feat_sim, ATen cosine_similarity: sum (x/max(|x|,1e-8)) (y/max(|y|,1e-8))
// counts[0..3] = edges newly removed by rule 1, 2, 3, 4 in that order (what the reference prints), counts[4] = kept.
template <int VEC>
__global__ void __launch_bounds__(256)
edge_validity_kernel(const long long* __restrict__ e0, const long long* __restrict__ e1, long long e,
                     const float* __restrict__ e_sim, const float* __restrict__ thr_conf,
                     const long long* __restrict__ pred_a, const long long* __restrict__ y_a,
                     const long long* __restrict__ pred_b, const long long* __restrict__ y_b,
                     const uint8_t* __restrict__ gate_a, const uint8_t* __restrict__ gate_b,
                     const float* __restrict__ x_a, const float* __restrict__ x_b, int d, float thres_feat_sim,
                     uint8_t* __restrict__ keep, unsigned long long* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float thr = thr_conf ? __ldg(thr_conf) : -INFINITY;
  unsigned c[5] = {0u, 0u, 0u, 0u, 0u};
  for (long long i = warp; i < e; i += nwarps) {
    const long long a = __ldg(e0 + i), b = __ldg(e1 + i);
    const float* __restrict__ pa = x_a + a * d;
    const float* __restrict__ pb = x_b + b * d;
    float sa = 0.f, sb = 0.f;
    if (VEC == 4) {
      for (int k = lane * 4; k < d; k += 128) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(pa + k)), w = __ldg(reinterpret_cast<const float4*>(pb + k));
        sa = fmaf(u.x, u.x, sa); sa = fmaf(u.y, u.y, sa); sa = fmaf(u.z, u.z, sa); sa = fmaf(u.w, u.w, sa);
        sb = fmaf(w.x, w.x, sb); sb = fmaf(w.y, w.y, sb); sb = fmaf(w.z, w.z, sb); sb = fmaf(w.w, w.w, sb);
      }
    } else {
      for (int k = lane; k < d; k += 32) { const float u = __ldg(pa + k), w = __ldg(pb + k); sa = fmaf(u, u, sa); sb = fmaf(w, w, sb); }
    }
    const float na = fmaxf(sqrtf(warp_sum(sa)), 1e-8f), nb = fmaxf(sqrtf(warp_sum(sb)), 1e-8f);
    float dot = 0.f;                                  // second sweep over the two rows: L1 hits
    if (VEC == 4) {
      for (int k = lane * 4; k < d; k += 128) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(pa + k)), w = __ldg(reinterpret_cast<const float4*>(pb + k));
        dot = fmaf(u.x / na, w.x / nb, dot); dot = fmaf(u.y / na, w.y / nb, dot);
        dot = fmaf(u.z / na, w.z / nb, dot); dot = fmaf(u.w / na, w.w / nb, dot);
      }
    } else {
      for (int k = lane; k < d; k += 32) dot = fmaf(__ldg(pa + k) / na, __ldg(pb + k) / nb, dot);
    }
    dot = warp_sum(dot);
    if (lane == 0) {
      const long long qa = __ldg(pred_a + a), qb = __ldg(pred_b + b);
      const bool gb = gate_b ? gate_b[b] != 0 : true;
      const bool ga = gate_a ? gate_a[b] != 0 : true;
      const bool r1 = __ldg(e_sim + i) < thr;
      const bool r2 = (qa != __ldg(y_a + a) && ga) || (qb != __ldg(y_b + b) && gb);
      const bool r3 = qa != qb;
      const bool r4 = dot < thres_feat_sim;
      const bool k1 = r1, k2 = k1 || r2, k3 = k2 || r3, k4 = k3 || r4;
      c[0] += k1; c[1] += (k2 && !k1); c[2] += (k3 && !k2); c[3] += (k4 && !k3); c[4] += !k4;
      keep[i] = k4 ? 0 : 1;
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < 5; ++r)
      if (c[r]) atomicAdd(counts + r, (unsigned long long)c[r]);
  }
}

__global__ void zero_counts_kernel(unsigned long long* counts) { if (threadIdx.x < 5) counts[threadIdx.x] = 0ull; }

int launch_edge_validity(const long long* e0, const long long* e1, long long e, const float* e_sim, const float* thr_conf,
                         const long long* pred_a, const long long* y_a, const long long* pred_b, const long long* y_b,
                         const uint8_t* gate_a, const uint8_t* gate_b, const float* x_a, const float* x_b, int d,
                         float thres_feat_sim, uint8_t* keep, long long* counts, cudaStream_t stream) {
  zero_counts_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<unsigned long long*>(counts));
  BGNN_LAUNCH_CHECK();
  if (e <= 0) return BGNN_OK;
  long long blocks = (e + 7) / 8;                     // 8 warps per CTA, one edge per warp and trip
  if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
  const bool v4 = d % 4 == 0 && ((reinterpret_cast<uintptr_t>(x_a) | reinterpret_cast<uintptr_t>(x_b)) & 15) == 0;
  if (v4)
    edge_validity_kernel<4><<<(unsigned)blocks, 256, 0, stream>>>(e0, e1, e, e_sim, thr_conf, pred_a, y_a, pred_b, y_b, gate_a,
        gate_b, x_a, x_b, d, thres_feat_sim, keep, reinterpret_cast<unsigned long long*>(counts));
  else
    edge_validity_kernel<1><<<(unsigned)blocks, 256, 0, stream>>>(e0, e1, e, e_sim, thr_conf, pred_a, y_a, pred_b, y_b, gate_a,
        gate_b, x_a, x_b, d, thres_feat_sim, keep, reinterpret_cast<unsigned long long*>(counts));
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

}  // namespace bgnn
