// Row-panel GEMM on the 5th-gen tensor cores with fp32-grade accuracy (3 x TF32), sm_100a:
//     Y[n, NO] = A[n, K] . B[NO, K]^T          n ~ 1e6 rows streamed once, K <= 256, NO <= 256
// for the node-wise dense part of a WIDE AdaptedConv (models/KTGNN.py:277-284 after the algebra of DESIGN.md):
//     forward   P = x [W_s; W_t; a_g_s2t[:D]; a_g_t2s[:D]]^T  with the node-wise epilogue (gates, biases, rank-1
//               corrections -> Hs, Ht) fused in, so P is never written;
//     backward  dx = dP . Wcat   (plain store).
// cuBLAS runs these as fp32 SIMT GEMMs (TF32 would cost 1e-3 relative, the parity bar is 1e-5); they are the
// largest item of the KT-GNN training step outside this library.  Here the streamed operand is split ON CHIP:
//     a ~= hi + lo,  hi = a rounded to tf32, lo = (a - hi) rounded to tf32   (|a - hi - lo| <= 2^-23 |a|)
//     a.b ~= lo.b_hi + hi.b_lo + hi.b_hi       (b_hi / b_lo planes of the small operand prepared by the host)
// so every row of A crosses HBM once and the work is bound by that stream, not by the MMAs.
//
// CTA = 10 warps, persistent over 128-row tiles:
//   warp 0     TMA producer: per k-block of 32 features the raw A tile [128 x 32] and the B_hi / B_lo planes
//              [NOP x 32] into a ring of SWIZZLE_128B stages
//   warps 2-5  split: rewrite the A tile in place as hi, write lo next to it (element-wise, layout-agnostic),
//              fence.proxy.async, signal the MMA issuer
//   warp 1     TMEM allocator + single-thread tcgen05.mma (kind::tf32) issuer, 128 x NOP fp32 accumulator
//              double-buffered in TMEM
//   warps 6-9  epilogue: thread <-> row, tcgen05.ld 32 columns at a time, store (EPI 0) or AdaptedConv node-wise
//              epilogue (EPI 1)
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace bgnn {

constexpr int RG_BM = 128;
constexpr int RG_BK = 32;            // fp32 elements per k-block = 128 B = one swizzle atom row
constexpr int RG_UMMA_K = 8;         // tf32
constexpr int RG_SPLIT_WARPS = 4;
constexpr int RG_THREADS = (2 + RG_SPLIT_WARPS + 4) * 32;   // TMA, MMA, split warps, 4 epilogue warps
constexpr int RG_A_PLANE = RG_BM * RG_BK * 4;   // 16 KB
constexpr int RG_SMEM_MAX = 232448;
constexpr int RG_EPI_STAGE = 4 * 32 * 32 * 4;      // epilogue staging tiles
constexpr int RG_SMEM_FIXED = 1024 + 512 + RG_EPI_STAGE;

__host__ __device__ constexpr uint32_t rg_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// nearest tf32 (ties away from zero in magnitude): the tensor core truncates its fp32 inputs to 19 bits, so hi is
// rounded here (|a - hi| <= 2^-12 |a|) and lo = a - hi is left to the hardware's truncation (error <= 2^-22 |a|)
__device__ __forceinline__ float rg_tf32(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }

struct RgEpi {                 // EPI 1: AdaptedConv node-wise epilogue (adapted_transform.cu, fused)
  const uint8_t* is_src;       // [n]
  const float* wd;             // [2C]
  const float* kg;             // [2]
  const float* bias;           // [2C] or null
  float* Hs;                   // [n, C]
  float* Ht;                   // [n, C]
  float* gates;                // [n, 2]  (EPI 2: [n, heads * 2])
  int c;
  int heads;                   // EPI 2: narrow convs side by side, columns h * (2c+2) + [0, 2c+2) of the accumulator
  // EPI 0 (plain Linear): y = act(acc * scale[col] + bias[col]) + res[row, col] -- eval-mode BatchNorm folded into
  // (scale, bias), ReLU / Tanh and a residual applied to the accumulator tile (the embedding producers of the build)
  const float* scale;          // [no] or null
  const float* res;            // [n, ld_res] or null
  int ld_res;
  int act;                     // 0 none, 1 relu, 2 tanh
};

// Epilogue store of one 32 x 32 block: the thread <-> row registers go through a 4 KB per-warp staging tile (16-byte
// chunks XOR row & 7: conflict-free both ways) so that every global store instruction writes four full 128-byte row
// segments instead of 32 scattered 16-byte pieces.
__device__ __forceinline__ void rg_store_block(float* stg, const float* r, int lane, float* __restrict__ out, long long row0,
                                               long long n, int ld, int col0, int ncols) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<float4*>(stg + lane * 32 + ((q ^ (lane & 7)) << 2)) = make_float4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
  __syncwarp();
  const int q = lane & 7, sub = lane >> 3;
  const int col = col0 + q * 4;
  const bool vec = (ld & 3) == 0 && col + 4 <= ncols;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = it * 4 + sub;
    const float4 v = *reinterpret_cast<const float4*>(stg + rr * 32 + ((q ^ (rr & 7)) << 2));
    const long long row = row0 + rr;
    if (row < n) {
      float* y = out + row * ld + col;
      if (vec) {
        *reinterpret_cast<float4*>(y) = v;
      } else {
        if (col < ncols) y[0] = v.x;
        if (col + 1 < ncols) y[1] = v.y;
        if (col + 2 < ncols) y[2] = v.z;
        if (col + 3 < ncols) y[3] = v.w;
      }
    }
  }
  __syncwarp();
}

// BRES: the B planes of ALL k-blocks stay resident in shared memory (loaded once per CTA); otherwise they travel with
// every stage.
template <int EPI, bool BRES>
__global__ void __launch_bounds__(RG_THREADS, 1)
rowpanel_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                     const __grid_constant__ CUtensorMap map_blo, long long n, int no, int nop, int kblocks, int stages,
                     float* __restrict__ Y, int ldy, RgEpi ep) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_plane = nop * RG_BK * 4;
  const int stage_bytes = 2 * RG_A_PLANE + (BRES ? 0 : 2 * b_plane);   // A (-> hi), A lo [, B hi, B lo]; multiples of 1024
  unsigned char* stage_base = smem;
  unsigned char* bres_base = smem + (size_t)stages * stage_bytes;       // BRES: [kblocks][hi, lo]
  float* epi_stage = reinterpret_cast<float*>(bres_base + (BRES ? (size_t)kblocks * 2 * b_plane : 0));   // 4 warps x 4 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(epi_stage) + RG_EPI_STAGE);
  uint64_t* full_bar = bars;                    // [stages] TMA landed
  uint64_t* ready_bar = bars + stages;          // [stages] split done (4 warps)
  uint64_t* empty_bar = bars + 2 * stages;      // [stages] MMAs retired
  uint64_t* tfull_bar = bars + 3 * stages;      // [2]
  uint64_t* tempty_bar = bars + 3 * stages + 2; // [2]
  uint64_t* bres_bar = bars + 3 * stages + 4;   // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * stages + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ntiles = (n + RG_BM - 1) / RG_BM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&ready_bar[s]), RG_SPLIT_WARPS);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&tfull_bar[b]), 1); mbar_init(smem_u32(&tempty_bar[b]), 4); }
    mbar_init(smem_u32(bres_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_bhi) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_blo) : "memory");
      if (BRES) {
        const uint32_t bb = smem_u32(bres_bar);
        mbar_expect_tx(bb, (uint32_t)(kblocks * 2 * b_plane));
        for (int kb = 0; kb < kblocks; ++kb) {
          tma_load_2d(smem_u32(bres_base + (size_t)kb * 2 * b_plane), &map_bhi, bb, kb * RG_BK, 0);
          tma_load_2d(smem_u32(bres_base + (size_t)kb * 2 * b_plane + b_plane), &map_blo, bb, kb * RG_BK, 0);
        }
      }
      int s = 0;
      uint32_t ph = 0;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int row0 = (int)(t * RG_BM);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
          const uint32_t fb = smem_u32(&full_bar[s]);
          mbar_expect_tx(fb, (uint32_t)(RG_A_PLANE + (BRES ? 0 : 2 * b_plane)));
          const uint32_t st = smem_u32(stage_base + (size_t)s * stage_bytes);
          tma_load_2d(st, &map_a, fb, kb * RG_BK, row0);
          if (!BRES) {
            tma_load_2d(st + 2 * RG_A_PLANE, &map_bhi, fb, kb * RG_BK, 0);
            tma_load_2d(st + 2 * RG_A_PLANE + b_plane, &map_blo, fb, kb * RG_BK, 0);
          }
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = rg_idesc_tf32(RG_BM, nop);
      if (BRES) mbar_wait(smem_u32(bres_bar), 0u);
      int s = 0;
      uint32_t ph = 0;
      int lt = 0;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++lt) {
        const int buf = lt & 1;
        const uint32_t tph = (uint32_t)(lt >> 1) & 1u;
        mbar_wait(smem_u32(&tempty_bar[buf]), tph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(smem_u32(&ready_bar[s]), ph);
          tc_fence_after();
          const uint32_t st = smem_u32(stage_base + (size_t)s * stage_bytes);
          const uint64_t a_hi = make_kmajor_sw128_desc(st), a_lo = make_kmajor_sw128_desc(st + RG_A_PLANE);
          const uint32_t bst = BRES ? smem_u32(bres_base + (size_t)kb * 2 * b_plane) : st + 2 * RG_A_PLANE;
          const uint64_t b_hi = make_kmajor_sw128_desc(bst);
          const uint64_t b_lo = make_kmajor_sw128_desc(bst + b_plane);
          // small terms first, then hi.hi
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            const uint64_t ad = (p == 0) ? a_lo : a_hi;
            const uint64_t bd = (p == 1) ? b_lo : b_hi;
#pragma unroll
            for (int k4 = 0; k4 < RG_BK / RG_UMMA_K; ++k4)
              tc_mma_tf32(d_tmem, ad + (uint64_t)(k4 * 2), bd + (uint64_t)(k4 * 2), idesc, (kb | p | k4) != 0 ? 1u : 0u);
          }
          tc_commit(smem_u32(&empty_bar[s]));
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
        tc_commit(smem_u32(&tfull_bar[buf]));
      }
    }
  } else if (warp < 2 + RG_SPLIT_WARPS) {
    // ===================== split: a -> (hi, lo), element-wise on the swizzled tile =====================
    constexpr int ST = RG_SPLIT_WARPS * 32;
    const int tid = threadIdx.x - 64;               // 0..ST-1
    int s = 0;
    uint32_t ph = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(smem_u32(&full_bar[s]), ph);
        float4* a = reinterpret_cast<float4*>(stage_base + (size_t)s * stage_bytes);
        float4* lo = reinterpret_cast<float4*>(stage_base + (size_t)s * stage_bytes + RG_A_PLANE);
#pragma unroll
        for (int i = 0; i < RG_A_PLANE / 16 / ST; ++i) {
          const float4 v = a[tid + i * ST];
          float4 h, l;
          h.x = rg_tf32(v.x); h.y = rg_tf32(v.y); h.z = rg_tf32(v.z); h.w = rg_tf32(v.w);
          // a - hi is exact in fp32 and at most 2^-12 |a|; the tensor core drops its low 13 bits: <= 2^-22 |a|
          l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
          a[tid + i * ST] = h;
          lo[tid + i * ST] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core reads
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&ready_bar[s]));
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue: thread <-> row =====================
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may read
    const int r_in_tile = quarter * 32 + lane;
    int lt = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const uint32_t tph = (uint32_t)(lt >> 1) & 1u;
      mbar_wait(smem_u32(&tfull_bar[buf]), tph);
      tc_fence_after();
      const long long row = t * RG_BM + r_in_tile;
      const long long row0 = t * RG_BM + quarter * 32;
      const bool row_ok = row < n;
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 256);
      float* stg = epi_stage + quarter * (32 * 32);
      float r[32];
      if (EPI == 0) {
        for (int c0 = 0; c0 < nop; c0 += 32) {
          tc_ld32(taddr0 + (uint32_t)c0, r);
          tc_wait_ld();
          if (ep.scale) {
#pragma unroll
            for (int k = 0; k < 32; ++k) r[k] *= (c0 + k < no) ? __ldg(ep.scale + c0 + k) : 1.f;
          }
          if (ep.bias) {
#pragma unroll
            for (int k = 0; k < 32; ++k) r[k] += (c0 + k < no) ? __ldg(ep.bias + c0 + k) : 0.f;
          }
          if (ep.act == 1) {
#pragma unroll
            for (int k = 0; k < 32; ++k) r[k] = fmaxf(r[k], 0.f);
          } else if (ep.act == 2) {
#pragma unroll
            for (int k = 0; k < 32; ++k) r[k] = tanhf(r[k]);
          }
          if (ep.res && row_ok) {
            const float* rp = ep.res + row * ep.ld_res + c0;
#pragma unroll
            for (int k = 0; k < 32; ++k) r[k] += (c0 + k < no) ? __ldg(rp + k) : 0.f;
          }
          rg_store_block(stg, r, lane, Y, row0, n, ldy, c0, no);
        }
      } else if (EPI == 2) {
        // classifier heads (C <= 4, 1-2 heads): all heads * (2C+2) <= 32 columns of a row sit in one tcgen05.ld; they
        // go through the staging tile (word k of row l at k ^ l: conflict-free) so that the head / column loops can
        // index them at run time
        const int c = ep.c, o = 2 * c + 2;
        tc_ld32(taddr0, r);
        tc_wait_ld();
#pragma unroll
        for (int k = 0; k < 32; ++k) stg[lane * 32 + (k ^ lane)] = r[k];
        __syncwarp();
        if (row_ok) {
          const float* mine = stg + lane * 32;
          const bool src = ep.is_src[row] != 0;
          for (int h = 0; h < ep.heads; ++h) {
            const float g0 = tanhf(mine[(h * o + 2 * c) ^ lane] + __ldg(ep.kg + 2 * h));
            const float g1 = tanhf(mine[(h * o + 2 * c + 1) ^ lane] + __ldg(ep.kg + 2 * h + 1));
            const float fs = src ? 0.f : g1, ft = src ? -g0 : 0.f;
            for (int j = 0; j < c; ++j) {
              const float ps = mine[(h * o + j) ^ lane] + (ep.bias ? __ldg(ep.bias + h * o + j) : 0.f);
              const float pt = mine[(h * o + c + j) ^ lane] + (ep.bias ? __ldg(ep.bias + h * o + c + j) : 0.f);
              ep.Hs[(row * ep.heads + h) * c + j] = fmaf(fs, __ldg(ep.wd + h * 2 * c + j), ps);
              ep.Ht[(row * ep.heads + h) * c + j] = fmaf(ft, __ldg(ep.wd + h * 2 * c + c + j), pt);
            }
            ep.gates[(row * ep.heads + h) * 2] = g0;
            ep.gates[(row * ep.heads + h) * 2 + 1] = g1;
          }
        }
        __syncwarp();
      } else {
        // columns [0,C) = x W_s^T, [C,2C) = x W_t^T, 2C / 2C+1 = gate logits (C a multiple of 32)
        const int c = ep.c;
        tc_ld32(taddr0 + (uint32_t)(2 * c), r);
        tc_wait_ld();
        const bool src = row_ok ? (ep.is_src[row] != 0) : false;
        const float g0 = tanhf(r[0] + __ldg(ep.kg)), g1 = tanhf(r[1] + __ldg(ep.kg + 1));
        const float fs = src ? 0.f : g1, ft = src ? -g0 : 0.f;
        if (row_ok) *reinterpret_cast<float2*>(ep.gates + row * 2) = make_float2(g0, g1);
        for (int c0 = 0; c0 < 2 * c; c0 += 32) {
          tc_ld32(taddr0 + (uint32_t)c0, r);
          tc_wait_ld();
          const bool second = c0 >= c;
          const float f = second ? ft : fs;
#pragma unroll
          for (int k = 0; k < 32; ++k)      // c0 + k = column of P = index into wd / bias
            r[k] = fmaf(f, __ldg(ep.wd + c0 + k), r[k] + (ep.bias ? __ldg(ep.bias + c0 + k) : 0.f));
          rg_store_block(stg, r, lane, second ? ep.Ht : ep.Hs, row0, n, c, second ? c0 - c : c0, c);
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

// ------------------------------------------------------------------------------------------ host
// 2-D fp32 tensor map, inner dimension `cols` (elements), row stride `ld` floats, box [32, box_rows], 128-B swizzle;
// out-of-bounds elements read as zero.
static int rg_make_map(CUtensorMap* m, const float* base, long long rows, int cols, int ld, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return BGNN_ERR_DRIVER;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)RG_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? BGNN_OK : BGNN_ERR_DRIVER;
}

// NOP (padded output width = UMMA N, multiple of 16), ring depth and whether B stays resident; 0 stages = shape not supported
static int rg_plan(int k, int no, int& nop, int& kblocks, bool& bres) {
  if (k < 1 || no < 1 || no > 256) return 0;
  nop = (no + 15) / 16 * 16;
  kblocks = (k + RG_BK - 1) / RG_BK;
  const int b_all = kblocks * 2 * nop * RG_BK * 4;
  int stages = (RG_SMEM_MAX - RG_SMEM_FIXED - b_all) / (2 * RG_A_PLANE);
  bres = stages >= 2;
  static const int force_stream = getenv("BGNN_RG_STREAM") ? atoi(getenv("BGNN_RG_STREAM")) : 0;
  if (force_stream && (RG_SMEM_MAX - RG_SMEM_FIXED) / (2 * RG_A_PLANE + 2 * nop * RG_BK * 4) > stages) bres = false;
  if (!bres) stages = (RG_SMEM_MAX - RG_SMEM_FIXED) / (2 * RG_A_PLANE + 2 * nop * RG_BK * 4);
  if (stages > 6) stages = 6;
  return stages >= 2 ? stages : 0;
}

bool rowpanel_gemm_supported(int k, int ld_a, int no) {
  int nop, kb;
  bool bres;
  return ld_a % 4 == 0 && rg_plan(k, no, nop, kb, bres) > 0;
}

// bhi / blo: [nop, kblocks*32] row-major planes of B (zero padded), both tf32-exact, bhi + blo ~= B.
template <int EPI>
static int rg_launch(const float* A, long long n, int k, int ld_a, const float* bhi, const float* blo, int no, float* Y,
                     int ldy, const RgEpi& ep, cudaStream_t stream) {
  if (n <= 0) return BGNN_OK;
  if (ld_a % 4 != 0 || n >= (1ll << 31)) return BGNN_ERR_INVALID_ARG;
  int nop, kblocks;
  bool bres;
  const int stages = rg_plan(k, no, nop, kblocks, bres);
  if (stages == 0) return BGNN_ERR_UNSUPPORTED;
  CUtensorMap ma, mbh, mbl;
  int rc;
  if ((rc = rg_make_map(&ma, A, n, k, ld_a, RG_BM)) != BGNN_OK) return rc;
  if ((rc = rg_make_map(&mbh, bhi, nop, kblocks * RG_BK, kblocks * RG_BK, nop)) != BGNN_OK) return rc;
  if ((rc = rg_make_map(&mbl, blo, nop, kblocks * RG_BK, kblocks * RG_BK, nop)) != BGNN_OK) return rc;
  const size_t b_plane2 = (size_t)2 * nop * RG_BK * 4;
  const size_t smem = RG_SMEM_FIXED + (bres ? (size_t)stages * 2 * RG_A_PLANE + kblocks * b_plane2
                                            : (size_t)stages * (2 * RG_A_PLANE + b_plane2));
  auto kern = bres ? rowpanel_gemm_kernel<EPI, true> : rowpanel_gemm_kernel<EPI, false>;
  BGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long ntiles = (n + RG_BM - 1) / RG_BM;
  const unsigned grid = (unsigned)(ntiles < kNumSMs ? ntiles : kNumSMs);
  kern<<<grid, RG_THREADS, smem, stream>>>(ma, mbh, mbl, n, no, nop, kblocks, stages, Y, ldy, ep);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

int launch_rowpanel_gemm(const float* A, long long n, int k, int ld_a, const float* bhi, const float* blo, const float* scale,
                         const float* bias, int act, const float* res, int ld_res, int no, float* Y, int ldy,
                         cudaStream_t stream) {
  RgEpi ep = {};
  ep.bias = bias;
  ep.scale = scale;
  ep.act = act;
  ep.res = res;
  ep.ld_res = ld_res;
  return rg_launch<0>(A, n, k, ld_a, bhi, blo, no, Y, ldy, ep, stream);
}

// hi / lo tf32 planes of a small fp32 matrix (any strides: a transposed view costs nothing), zero padded to
// [rows_p, cols_p]: the host-side operand preparation of the kernels above in one launch
__global__ void __launch_bounds__(256)
tf32_planes_kernel(const float* __restrict__ w, int rows, int cols, long long stride_r, long long stride_c, int rows_p, int cols_p,
                   float* __restrict__ hi, float* __restrict__ lo) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows_p * cols_p) return;
  const int r = idx / cols_p, c = idx - r * cols_p;
  const float v = (r < rows && c < cols) ? __ldg(w + r * stride_r + c * stride_c) : 0.f;
  const float h = rg_tf32(v);
  hi[idx] = h;
  lo[idx] = rg_tf32(v - h);
}

int launch_tf32_planes(const float* w, int rows, int cols, long long stride_r, long long stride_c, int rows_p, int cols_p, float* hi,
                       float* lo, cudaStream_t stream) {
  if (rows < 0 || cols < 0 || rows_p < rows || cols_p < cols) return BGNN_ERR_INVALID_ARG;
  const long long total = (long long)rows_p * cols_p;
  if (total == 0) return BGNN_OK;
  if (total > (1ll << 30)) return BGNN_ERR_UNSUPPORTED;
  tf32_planes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(w, rows, cols, stride_r, stride_c, rows_p, cols_p, hi, lo);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

bool adapted_skinny_tc_supported(int c, int d, int heads) {
  return c >= 1 && c <= 4 && heads >= 1 && heads <= 2 && d % 4 == 0 && d >= 4 && d <= 256 &&
         rowpanel_gemm_supported(d, d, heads * (2 * c + 2));
}

int launch_adapted_skinny_tc_fwd(const float* x, long long n, int d, const float* wcat_hi, const float* wcat_lo, int c, int heads,
                                 const uint8_t* is_src, const float* wd, const float* kg, const float* bias, float* Hs,
                                 float* Ht, float* gates, cudaStream_t stream) {
  if (!adapted_skinny_tc_supported(c, d, heads)) return BGNN_ERR_UNSUPPORTED;
  RgEpi ep;
  ep.is_src = is_src; ep.wd = wd; ep.kg = kg; ep.bias = bias; ep.Hs = Hs; ep.Ht = Ht; ep.gates = gates; ep.c = c; ep.heads = heads;
  return rg_launch<2>(x, n, d, d, wcat_hi, wcat_lo, heads * (2 * c + 2), nullptr, 0, ep, stream);
}

bool adapted_wide_supported(int c, int d) { return c >= 32 && c % 32 == 0 && 2 * c + 2 <= 256 && d % 4 == 0 && d >= 4 && d <= 256 && rowpanel_gemm_supported(d, d, 2 * c + 2); }

int launch_adapted_wide_fwd(const float* x, long long n, int d, const float* wcat_hi, const float* wcat_lo, int c,
                            const uint8_t* is_src, const float* wd, const float* kg, const float* bias, float* Hs, float* Ht,
                            float* gates, cudaStream_t stream) {
  if (!adapted_wide_supported(c, d)) return BGNN_ERR_UNSUPPORTED;
  RgEpi ep;
  ep.is_src = is_src; ep.wd = wd; ep.kg = kg; ep.bias = bias; ep.Hs = Hs; ep.Ht = Ht; ep.gates = gates; ep.c = c; ep.heads = 1;
  return rg_launch<1>(x, n, d, d, wcat_hi, wcat_lo, 2 * c + 2, nullptr, 0, ep, stream);
}

}  // namespace bgnn
