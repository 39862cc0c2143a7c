// Sub-warp row-vector helpers for the gather kernels (SpMM, fused AdaptedConv aggregation).
// A group of G lanes (G = 1..32, power of two) owns one CSR row; lane g of the group holds CH chunks
// of VEC consecutive features: columns (ch*G + g)*VEC .. +VEC-1.  VEC=4 -> 128-bit loads.
#pragma once
#include "common.cuh"

namespace bgnn {

template <int VEC>
struct Chunk {
  float v[VEC];
};

template <int VEC>
__device__ __forceinline__ Chunk<VEC> ld_chunk(const float* __restrict__ p, bool ok) {
  Chunk<VEC> c;
  if (VEC == 4) {
    float4 t = ok ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
    c.v[0] = t.x; c.v[1 % VEC] = t.y; c.v[2 % VEC] = t.z; c.v[3 % VEC] = t.w;
  } else if (VEC == 2) {
    float2 t = ok ? __ldg(reinterpret_cast<const float2*>(p)) : make_float2(0.f, 0.f);
    c.v[0] = t.x; c.v[1 % VEC] = t.y;
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) c.v[i] = ok ? __ldg(p + i) : 0.f;
  }
  return c;
}

template <int VEC>
__device__ __forceinline__ void st_chunk(float* __restrict__ p, const Chunk<VEC>& c, bool ok) {
  if (!ok) return;
  if (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(c.v[0], c.v[1 % VEC], c.v[2 % VEC], c.v[3 % VEC]);
  } else if (VEC == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(c.v[0], c.v[1 % VEC]);
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) p[i] = c.v[i];
  }
}

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Pick (VEC, G, CH) for a feature width; returns false if unsupported.
inline bool pick_row_config(int c, int& vec, int& g, int& ch) {
  if (c <= 0) return false;
  vec = (c % 4 == 0) ? 4 : 1;
  int units = (c + vec - 1) / vec;
  g = 1;
  while (g < 32 && g < units) g <<= 1;
  ch = (units + g - 1) / g;
  return ch <= 4;
}

// Dispatch a functor templated on <VEC, G, CH>.
#define BGNN_ROW_DISPATCH(vec, g, ch, CALL)                                         \
  do {                                                                              \
    if (vec == 4) {                                                                 \
      switch (g) {                                                                  \
        case 1: BGNN_ROW_DISPATCH_CH(4, 1, ch, CALL); break;                        \
        case 2: BGNN_ROW_DISPATCH_CH(4, 2, ch, CALL); break;                        \
        case 4: BGNN_ROW_DISPATCH_CH(4, 4, ch, CALL); break;                        \
        case 8: BGNN_ROW_DISPATCH_CH(4, 8, ch, CALL); break;                        \
        case 16: BGNN_ROW_DISPATCH_CH(4, 16, ch, CALL); break;                      \
        default: BGNN_ROW_DISPATCH_CH(4, 32, ch, CALL); break;                      \
      }                                                                             \
    } else {                                                                        \
      switch (g) {                                                                  \
        case 1: BGNN_ROW_DISPATCH_CH(1, 1, ch, CALL); break;                        \
        case 2: BGNN_ROW_DISPATCH_CH(1, 2, ch, CALL); break;                        \
        case 4: BGNN_ROW_DISPATCH_CH(1, 4, ch, CALL); break;                        \
        case 8: BGNN_ROW_DISPATCH_CH(1, 8, ch, CALL); break;                        \
        case 16: BGNN_ROW_DISPATCH_CH(1, 16, ch, CALL); break;                      \
        default: BGNN_ROW_DISPATCH_CH(1, 32, ch, CALL); break;                      \
      }                                                                             \
    }                                                                               \
  } while (0)

#define BGNN_ROW_DISPATCH_CH(V, G_, ch, CALL)                                       \
  do {                                                                              \
    if (G_ < 32 || ch == 1) { CALL(V, G_, 1); }                                     \
    else if (ch == 2) { CALL(V, G_, 2); }                                           \
    else if (ch == 3) { CALL(V, G_, 3); }                                           \
    else { CALL(V, G_, 4); }                                                        \
  } while (0)

}  // namespace bgnn
