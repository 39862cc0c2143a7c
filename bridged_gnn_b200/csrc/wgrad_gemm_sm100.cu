// Weight-gradient contraction on the 5th-gen tensor cores with fp32-grade accuracy (3 x TF32), sm_100a:
//     W[j, f] = sum_i G[i, j] X[i, f]          i over ~1e6 node rows, j < no <= 256, f < d <= 128
// = g_w_cat = dP^T x of AdaptedConv's node-wise part (models/KTGNN.py:277-284) and the weight gradient of the
// Linear layers of the classifier transformer (models/KTGNN.py:363).  cuBLAS
// runs these as fp32 SIMT "NT" GEMMs with a tiny output and a 1e6-long reduction (1.2 ms + 2 x 0.6 ms of the
// KT-GNN training step on one B200); here both operands cross HBM once and the reduction runs on tcgen05.
//
// The reduction index is the ROW index of both operands, so both are MN-major in shared memory.  For 32-bit
// operands tcgen05 takes MN-major tiles only in the "128-byte swizzle with 32-byte atomicity" layout
// (UMMA layout type 1, cute::UMMA::Layout_MN_SW128_32B_Atom: 4 rows x 128 B, 32-byte chunks XOR row mod 4) --
// measured here: with the ordinary SWIZZLE_128B descriptor and a transposed tf32 operand the MMA completes and
// writes NOTHING.  TMA has the matching mode (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): a box [32 columns x 32 rows]
// is a column of eight such atoms, so LBO = 4096 B (next box = next 32 columns), SBO = 512 B (next 4 rows), and one
// tf32 MMA (K = 8) eats two atoms = 1024 B.
// Both streamed operands are split on chip (hi = tf32(a), lo = tf32(a - hi)) by eight warps, element-wise and
// therefore layout-agnostic;  D[f, j] (128 TMEM lanes x nop columns) += lo.hi + hi.lo + hi.hi.
//
// Persistent CTAs stride over the 32-row blocks; the 128 x nop accumulator is double-buffered in TMEM and flushed
// every WG_FLUSH blocks into the CTA's fp32 partial [128, nop] (see the epilogue), and a second kernel adds the
// partials in CTA order, in double, into W (transposed).  Deterministic.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace bgnn {

constexpr int WG_BK = 32;              // rows per stage
constexpr int WG_BOX = 32 * WG_BK * 4; // one TMA box: 32 columns x 32 rows fp32 = 4 KB
constexpr int WG_MBOX = 4;             // X boxes per stage = 128 features (UMMA M)
constexpr int WG_SPLIT_WARPS = 8;
constexpr int WG_THREADS = (2 + WG_SPLIT_WARPS + 4) * 32;   // TMA, MMA, split warps, 4 epilogue warps
constexpr int WG_SMEM_MAX = 232448;
constexpr int WG_SMEM_FIXED = 1024 + 512;
constexpr int WG_FLUSH = 16;            // K-blocks (512 rows) accumulated in TMEM before the tile is added to the CTA's fp32 partial

__host__ __device__ constexpr uint32_t wg_idesc_tf32_mn(int m, int n) {
  // D fp32, A/B tf32, both MN-major (bits 15, 16)
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// MN-major fp32/tf32 operand, 128-byte swizzle with 32-byte atomicity: LBO = stride between 32-column atoms,
// SBO = stride between 4-row groups
__device__ __forceinline__ uint64_t wg_mn_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
  d |= (uint64_t)(WG_BOX >> 4) << 16;       // LBO: next 32-column atom column = next TMA box
  d |= (uint64_t)(512 >> 4) << 32;          // SBO: next 4-row atom
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                   // SWIZZLE_128B_BASE32B
  return d;
}

__device__ __forceinline__ float wg_tf32(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }

// stage layout: hi plane [G: nb boxes][X: 4 boxes] then the lo plane, same order.  TMA fills the first nb + mb boxes
// of the hi plane with the raw data; X boxes mb..3 (features >= d) are zeroed once and never touched again.
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_g,
                  const __grid_constant__ CUtensorMap map_g1, const __grid_constant__ CUtensorMap map_g2, int nb0, int nb1,
                  long long nblocks, int mb, int nb, int stages, int ones_col, float* __restrict__ part) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int plane = (nb + WG_MBOX) * WG_BOX;
  const int stage_bytes = 2 * plane;
  const int nop = nb * 32;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* ready_bar = bars + stages;
  uint64_t* empty_bar = bars + 2 * stages;
  uint64_t* tfull_bar = bars + 3 * stages;       // [2] accumulator buffer ready for the epilogue
  uint64_t* tempty_bar = bars + 3 * stages + 2;  // [2] drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * stages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&ready_bar[s]), WG_SPLIT_WARPS);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&tfull_bar[b]), 1); mbar_init(smem_u32(&tempty_bar[b]), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (mb < WG_MBOX) {        // zero the X boxes no TMA ever writes (both planes, every stage)
    const int zero_f4 = (WG_MBOX - mb) * WG_BOX / 16;
    for (int s = 0; s < stages; ++s)
      for (int pl = 0; pl < 2; ++pl) {
        float4* z = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes + (size_t)pl * plane + (size_t)(nb + mb) * WG_BOX);
        for (int i = threadIdx.x; i < zero_f4; i += WG_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    if (ones_col) {          // feature mb*32 := 1 in every row, so that row mb*32 of D collects the column sums of G
      __syncthreads();
      if (threadIdx.x < WG_BK)
        for (int s = 0; s < stages; ++s) {
          float* z = reinterpret_cast<float*>(smem + (size_t)s * stage_bytes + (size_t)(nb + mb) * WG_BOX);
          z[threadIdx.x * 32 + ((threadIdx.x & 3) << 3)] = 1.0f;      // column 0 of row r sits in 32-byte chunk r & 3
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_g) : "memory");
      if (nb0 < nb) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_g1) : "memory");
      if (nb0 + nb1 < nb) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_g2) : "memory");
      int s = 0;
      uint32_t ph = 0;
      for (long long b = blockIdx.x; b < nblocks; b += gridDim.x) {
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
        const uint32_t fb = smem_u32(&full_bar[s]);
        mbar_expect_tx(fb, (uint32_t)((nb + mb) * WG_BOX));
        const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
        const int row0 = (int)(b * WG_BK);
        // G = up to three column blocks side by side (each from its own matrix: [dHs | dHt | d gates] without a copy)
        for (int j = 0; j < nb; ++j) {
          const CUtensorMap* m = j < nb0 ? &map_g : (j < nb0 + nb1 ? &map_g1 : &map_g2);
          const int jj = j < nb0 ? j : (j < nb0 + nb1 ? j - nb0 : j - nb0 - nb1);
          tma_load_2d(st + j * WG_BOX, m, fb, jj * 32, row0);
        }
        for (int j = 0; j < mb; ++j) tma_load_2d(st + (nb + j) * WG_BOX, &map_x, fb, j * 32, row0);
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = wg_idesc_tf32_mn(128, nop);
      int s = 0;
      uint32_t ph = 0;
      int lb = 0;                       // this CTA's block counter; WG_FLUSH blocks share one accumulator buffer
      for (long long b = blockIdx.x; b < nblocks; b += gridDim.x, ++lb) {
        const int grp = lb / WG_FLUSH, buf = grp & 1;
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256);
        const bool first = lb % WG_FLUSH == 0;
        if (first) {
          mbar_wait(smem_u32(&tempty_bar[buf]), (((uint32_t)(grp >> 1)) & 1u) ^ 1u);
          tc_fence_after();
        }
        mbar_wait(smem_u32(&ready_bar[s]), ph);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
        const uint64_t g_hi = wg_mn_desc(st), x_hi = wg_mn_desc(st + nb * WG_BOX);
        const uint64_t g_lo = wg_mn_desc(st + plane), x_lo = wg_mn_desc(st + plane + nb * WG_BOX);
        uint32_t acc = first ? 0u : 1u;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          const uint64_t ad = (p == 0) ? x_lo : x_hi;      // A = X (M = features), B = G (N = output columns)
          const uint64_t bd = (p == 1) ? g_lo : g_hi;
#pragma unroll
          for (int ks = 0; ks < WG_BK / 8; ++ks) {
            tc_mma_tf32(d_tmem, ad + (uint64_t)(ks * (1024 >> 4)), bd + (uint64_t)(ks * (1024 >> 4)), idesc, acc);
            acc = 1u;
          }
        }
        tc_commit(smem_u32(&empty_bar[s]));
        if (lb % WG_FLUSH == WG_FLUSH - 1 || b + gridDim.x >= nblocks) tc_commit(smem_u32(&tfull_bar[buf]));
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp < 2 + WG_SPLIT_WARPS) {
    // ===================== split raw -> (hi, lo) =====================
    const int tid = threadIdx.x - 64;
    const int n_f4 = (nb + mb) * WG_BOX / 16;
    int s = 0;
    uint32_t ph = 0;
    for (long long b = blockIdx.x; b < nblocks; b += gridDim.x) {
      mbar_wait(smem_u32(&full_bar[s]), ph);
      float4* a = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes);
      float4* lo = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes + plane);
#pragma unroll 4
      for (int i = tid; i < n_f4; i += WG_SPLIT_WARPS * 32) {
        const float4 v = a[i];
        float4 h, l;
        h.x = wg_tf32(v.x); h.y = wg_tf32(v.y); h.z = wg_tf32(v.z); h.w = wg_tf32(v.w);
        l.x = wg_tf32(v.x - h.x); l.y = wg_tf32(v.y - h.y); l.z = wg_tf32(v.z - h.z); l.w = wg_tf32(v.w - h.w);
        a[i] = h;
        lo[i] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&ready_bar[s]));
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
  } else {
    // ===================== epilogue: the CTA's partial, thread <-> feature row =====================
    // The tensor core's fp32 accumulation is not round-to-nearest: over the ~2700 accumulations a CTA would chain at
    // n = 1e6 the error grew to 3e-7 of sum |g||x| (20 x an fp32 FMA chain).  Every WG_FLUSH blocks the tile is
    // therefore added (fp32, round to nearest, fixed order) to the CTA's partial in global memory -- L2-resident --
    // while the MMAs continue in the other TMEM buffer.
    const int quarter = warp & 3;
    const int f = quarter * 32 + lane;
    // partial layout [cta][column quad q = j / 4][feature f][4]: the 128 epilogue threads (one per f) touch 2 KB
    // contiguous per quad, so the read-modify-write of a flush is coalesced
    float4* out4 = reinterpret_cast<float4*>(part + (size_t)blockIdx.x * 128 * nop) + f;
    const long long mine = (long long)blockIdx.x < nblocks ? (nblocks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int ngroups = (int)((mine + WG_FLUSH - 1) / WG_FLUSH);
    for (int grp = 0; grp < ngroups; ++grp) {
      const int buf = grp & 1;
      mbar_wait(smem_u32(&tfull_bar[buf]), ((uint32_t)(grp >> 1)) & 1u);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 256);
      float r[32];
      for (int c0 = 0; c0 < nop; c0 += 32) {
        tc_ld32(taddr0 + (uint32_t)c0, r);
        tc_wait_ld();
        float4* o4 = out4 + (size_t)(c0 >> 2) * 128;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float4 v = make_float4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
          if (grp > 0) {
            const float4 old = o4[k * 128];
            v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
          }
          o4[k * 128] = v;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[buf]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// W[j * ldw + f] = sum over CTAs of part[cta](f, j), in double: eight lanes per element each add every eighth partial
// in order, then a fixed butterfly combines them (deterministic);  colsum[j] = the same for the all-ones feature row
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ part, int nparts, int nop, int no, int d, float* __restrict__ W, int ldw,
                    float* __restrict__ colsum, int f_ones) {
  const int sub = threadIdx.x & 7;
  const int idx = blockIdx.x * 32 + (threadIdx.x >> 3);     // f * no + j: consecutive groups read consecutive j
  const int total = (d + (colsum ? 1 : 0)) * no;
  const bool live = idx < total;
  int f = live ? idx / no : 0;
  const int j = live ? idx - f * no : 0;
  const bool extra = f == d;
  if (extra) f = f_ones;
  const size_t at = ((size_t)(j >> 2) * 128 + f) * 4 + (j & 3);       // [quad][feature][4] inside a CTA's partial
  double acc = 0.0;
  if (live)
    for (int p = sub; p < nparts; p += 8) acc += (double)part[(size_t)p * 128 * nop + at];
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (live && sub == 0) {
    if (extra) colsum[j] = (float)acc;
    else W[(size_t)j * ldw + f] = (float)acc;
  }
}

static int wg_make_map(CUtensorMap* m, const float* base, long long rows, int cols, int ld) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return BGNN_ERR_DRIVER;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)WG_BK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? BGNN_OK : BGNN_ERR_DRIVER;
}

static int wg_stages(int nb) {
  int st = (WG_SMEM_MAX - WG_SMEM_FIXED) / (2 * (nb + WG_MBOX) * WG_BOX);
  return st > 6 ? 6 : st;
}

bool wgrad_gemm_supported(int d, int ld_x, int no, int ld_g) {
  return d >= 1 && d <= 128 && no >= 1 && no <= 256 && ld_x % 4 == 0 && ld_g % 4 == 0 && wg_stages((no + 31) / 32) >= 2;
}

size_t wgrad_gemm_workspace_bytes(int no) { return (size_t)kNumSMs * 128 * ((no + 31) / 32 * 32) * sizeof(float) + 256; }

int launch_wgrad_gemm_cat(const float* const* G, const int* ld_g, const int* no_blk, int nblk, const float* X, int ld_x, int d,
                          long long n, float* W, int ldw, float* colsum, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (nblk < 1 || nblk > 3) return BGNN_ERR_INVALID_ARG;
  int no = 0, nbs[3] = {0, 0, 0};
  for (int i = 0; i < nblk; ++i) {
    if (no_blk[i] < 1 || ld_g[i] % 4 != 0 || ld_g[i] < no_blk[i]) return BGNN_ERR_UNSUPPORTED;
    if (i + 1 < nblk && no_blk[i] % 32 != 0) return BGNN_ERR_UNSUPPORTED;      // W rows must stay contiguous
    no += no_blk[i];
    nbs[i] = (no_blk[i] + 31) / 32;
  }
  if (!wgrad_gemm_supported(d, ld_x, no, 4) || n >= (1ll << 31)) return BGNN_ERR_UNSUPPORTED;
  if (colsum && d > 96) return BGNN_ERR_UNSUPPORTED;      // the all-ones feature needs a spare box
  if (ws_bytes < wgrad_gemm_workspace_bytes(no)) return BGNN_ERR_WORKSPACE;
  const int nb = nbs[0] + nbs[1] + nbs[2], mb = (d + 31) / 32;
  const int nop = nb * 32;
  float* part = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  const long long nblocks = (n + WG_BK - 1) / WG_BK;
  const unsigned grid = (unsigned)(nblocks < kNumSMs ? (nblocks > 0 ? nblocks : 1) : kNumSMs);
  if (n > 0) {
    CUtensorMap mx, mg[3];
    int rc;
    if ((rc = wg_make_map(&mx, X, n, d, ld_x)) != BGNN_OK) return rc;
    for (int i = 0; i < 3; ++i) {
      const int k = i < nblk ? i : 0;
      if ((rc = wg_make_map(&mg[i], G[k], n, no_blk[k], ld_g[k])) != BGNN_OK) return rc;
    }
    const int stages = wg_stages(nb);
    const size_t smem = WG_SMEM_FIXED + (size_t)stages * 2 * (nb + WG_MBOX) * WG_BOX;
    BGNN_CUDA_TRY(cudaFuncSetAttribute(wgrad_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad_gemm_kernel<<<grid, WG_THREADS, smem, stream>>>(mx, mg[0], mg[1], mg[2], nbs[0], nbs[1], nblocks, mb, nb, stages,
                                                          colsum ? 1 : 0, part);
    BGNN_LAUNCH_CHECK();
  }
  wgrad_reduce_kernel<<<((d + 1) * no + 31) / 32, 256, 0, stream>>>(part, n > 0 ? (int)grid : 0, nop, no, d, W, ldw, colsum,
                                                                      mb * 32);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

int launch_wgrad_gemm(const float* G, int ld_g, int no, const float* X, int ld_x, int d, long long n, float* W, int ldw,
                      float* colsum, void* ws, size_t ws_bytes, cudaStream_t stream) {
  return launch_wgrad_gemm_cat(&G, &ld_g, &no, 1, X, ld_x, d, n, W, ldw, colsum, ws, ws_bytes, stream);
}

}  // namespace bgnn
