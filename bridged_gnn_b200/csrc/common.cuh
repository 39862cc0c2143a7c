// Shared device/host helpers for libbgnn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define BGNN_OK 0
#define BGNN_ERR_INVALID_ARG (-1)
#define BGNN_ERR_WORKSPACE (-2)
#define BGNN_ERR_UNSUPPORTED (-3)
#define BGNN_ERR_DRIVER (-4)

#define BGNN_CUDA_TRY(expr)                          \
  do {                                               \
    cudaError_t _e = (expr);                         \
    if (_e != cudaSuccess) return (int)_e;           \
  } while (0)

#define BGNN_LAUNCH_CHECK()                          \
  do {                                               \
    cudaError_t _e = cudaGetLastError();             \
    if (_e != cudaSuccess) return (int)_e;           \
  } while (0)

namespace bgnn {

constexpr int kNumSMs = 148;  // B200

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace (256-B aligned slices).
struct Workspace {
  char* base;
  size_t size;
  size_t off;
  __host__ Workspace(void* p, size_t n) : base((char*)p), size(n), off(0) {}
  template <typename T>
  __host__ T* take(size_t count) {
    off = align_up(off, 256);
    T* r = (T*)(base + off);
    off += count * sizeof(T);
    return r;
  }
  __host__ bool ok() const { return off <= size; }
};

// Same arithmetic on both sides of the parity test: sigmoid in fp32 as 1/(1+exp(-x)).
__device__ __forceinline__ float sigmoid_f32(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------
// Thread-private top-KC list living in shared memory, organised as a binary min-heap whose root is the
// WORST kept element.  Thread t owns column t of val[KC][stride] / idx[KC][stride] (bank-conflict-free).
// Candidates must arrive in increasing index order, which makes "strictly greater than the current
// worst" equivalent to the parity key (value desc, index asc): an equal value with a larger index loses.
// ---------------------------------------------------------------------------------------------
struct ListState {
  int cnt;      // filled slots
  float thr;    // value of the worst kept element (-inf until the list is full)
};

__device__ __forceinline__ ListState list_init() {
  ListState s; s.cnt = 0; s.thr = -INFINITY; return s;
}

// a ranks below b under the parity key
__device__ __forceinline__ bool list_worse(float av, int ai, float bv, int bi) {
  return av < bv || (av == bv && ai > bi);
}

// explicit shared-window accessors: the lists live in dynamic shared memory and generic LD/ST would
// cost an address-space check per access
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ int lds_s32(uint32_t a) {
  int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_s32(uint32_t a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// val / idx: shared-window byte addresses of this thread's column; sb: byte stride between slots.
// Precondition: v > st.thr (always true while the list is not full).  O(log KC) shared-memory accesses.
__device__ __forceinline__ void list_push(uint32_t val, uint32_t idx, uint32_t sb, int kc, ListState& st, float v, int j) {
  if (st.cnt < kc) {
    int pos = st.cnt++;
    while (pos > 0) {                       // sift up: parents must be worse than children
      const int par = (pos - 1) >> 1;
      const float pv = lds_f32(val + par * sb);
      const int pi = lds_s32(idx + par * sb);
      if (!list_worse(v, j, pv, pi)) break;
      sts_f32(val + pos * sb, pv); sts_s32(idx + pos * sb, pi);
      pos = par;
    }
    sts_f32(val + pos * sb, v); sts_s32(idx + pos * sb, j);
    if (st.cnt == kc) st.thr = lds_f32(val);
  } else {
    int pos = 0;                            // replace the root, sift down
    while (true) {
      const int l = 2 * pos + 1;
      if (l >= kc) break;
      float cv = lds_f32(val + l * sb);
      int ci = lds_s32(idx + l * sb), ch = l;
      if (l + 1 < kc) {
        const float rv = lds_f32(val + (l + 1) * sb);
        const int ri = lds_s32(idx + (l + 1) * sb);
        if (list_worse(rv, ri, cv, ci)) { cv = rv; ci = ri; ch = l + 1; }
      }
      if (!list_worse(cv, ci, v, j)) break;
      sts_f32(val + pos * sb, cv); sts_s32(idx + pos * sb, ci);
      pos = ch;
    }
    sts_f32(val + pos * sb, v); sts_s32(idx + pos * sb, j);
    st.thr = lds_f32(val);
  }
}

// Out-of-line variant for call sites that must not grow (hot loops with many live registers).
static __device__ __noinline__ ListState list_insert(uint32_t val, uint32_t idx, uint32_t sb, int kc, ListState st, float v, int j) {
  list_push(val, idx, sb, kc, st, v, j);
  return st;
}

}  // namespace bgnn
