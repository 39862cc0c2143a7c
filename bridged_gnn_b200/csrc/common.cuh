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
// Thread-private top-KC list living in shared memory.  Thread t owns column t of val[KC][stride]
// / idx[KC][stride] (bank-conflict-free).  Candidates must arrive in increasing index order, which
// makes "strictly greater than the current worst" equivalent to the (value desc, index asc) key.
// ---------------------------------------------------------------------------------------------
struct ListState {
  int cnt;      // filled slots
  float thr;    // value of the worst kept element (-inf until full)
  int worst;    // slot of the worst kept element
};

__device__ __forceinline__ ListState list_init() {
  ListState s; s.cnt = 0; s.thr = -INFINITY; s.worst = 0; return s;
}

// Precondition: v > st.thr (always true while the list is not full).  Kept out of line: it runs
// ~k*ln(N/k) times per row over a whole sweep, so it must not bloat the hot loop.
static __device__ __noinline__ ListState list_insert(float* val, int* idx, int stride, int kc, ListState st, float v, int j) {
  int slot = (st.cnt < kc) ? st.cnt : st.worst;
  val[slot * stride] = v;
  idx[slot * stride] = j;
  if (st.cnt < kc) ++st.cnt;
  if (st.cnt == kc) {
    float w = val[0];
    int wi = idx[0], ws = 0;
    for (int s = 1; s < kc; ++s) {
      float x = val[s * stride];
      int i = idx[s * stride];
      if (x < w || (x == w && i > wi)) { w = x; wi = i; ws = s; }
    }
    st.thr = w; st.worst = ws;
  }
  return st;
}

}  // namespace bgnn
