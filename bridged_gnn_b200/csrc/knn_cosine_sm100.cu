// Cosine-similarity kNN sweep on the 5th-gen tensor cores (sm_100a): TMA -> shared memory ->
// tcgen05.mma (kind::tf32, accumulators in TMEM) -> tcgen05.ld -> fused per-row top-KC selection.
// The [nq, ndb] score matrix never leaves the SM.
//
// Replaces the pair-materialising similarity + sim_mat.topk of the reference
// (models/models.py:124-130 / 945-948, main_bridged_graph.py:45-67, 90-111) for the cosine head.
//
// Exactness.  Parity is defined on fp32 similarities, tf32 has a 10-bit mantissa.  Inputs are unit
// rows split as x = hi + lo with hi tf32-exact (knn_simt.cu: normalize_split_kernel):
//   PASSES = 3 : lo.hi + hi.lo + hi.hi accumulated in fp32 (error ~1e-6 on |cos| <= 1)
//   PASSES = 1 : hi.hi only (error <= ~2e-3)
// Either way this kernel only *nominates* KC > k candidates per row from approximate scores; the
// merge kernel re-scores the nominees exactly in fp32, selects under the parity key and certifies
// each row against (list threshold + error bound); rows that cannot be certified are re-done by the
// exact CUDA-core kernel.
//
// CTA = 6 warps, one (128-query block, db split) work unit:
//   warp 0    TMA producer: per k-block of 32 features, A planes [128 x 32] and B planes [BN x 32]
//             into a ring of SWIZZLE_128B stages
//   warp 1    TMEM allocator + single-thread tcgen05.mma issuer, 128 x BN fp32 accumulator,
//             double-buffered in TMEM so the epilogue of tile t overlaps the MMAs of tile t+1
//   warps 2-5 epilogue: thread <-> query row (TMEM lane), 32 columns per tcgen05.ld, running
//             threshold test, rare insertion into a thread-private list in shared memory
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace bgnn {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;            // fp32 elements per k-block = 128 B = one swizzle atom row
constexpr int TC_UMMA_K = 8;         // tf32
constexpr int TC_THREADS = 192;
constexpr int TC_A_PLANE = TC_BM * TC_BK * 4;   // 16 KB

template <int PASSES, int BN>
struct TcCfg {
  static constexpr int PLANES = (PASSES == 3) ? 2 : 1;
  static constexpr int B_PLANE = BN * TC_BK * 4;
  static constexpr int STAGE_BYTES = PLANES * (TC_A_PLANE + B_PLANE);
  static constexpr int TMEM_COLS = 2 * BN;   // 512 or 256: powers of two
};
constexpr int TC_SMEM_MAX = 232448;          // 227 KB opt-in limit per CTA
constexpr int TC_SMEM_FIXED = 1024 + 512;    // alignment slack + barriers / tmem slot

// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) [4,6), a/b format TF32 (2)
// [7,10)/[10,13), both K-major, N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ------------------------------------------------------------------------------------------ kernel
template <int PASSES, int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
knn_cosine_tc_kernel(const __grid_constant__ CUtensorMap map_qhi, const __grid_constant__ CUtensorMap map_qlo,
                     const __grid_constant__ CUtensorMap map_dhi, const __grid_constant__ CUtensorMap map_dlo,
                     int nq, int ndb, int kblocks, int tiles_total, int tiles_per_split, int kc, int STAGES,
                     float* __restrict__ cand_val, int* __restrict__ cand_idx) {
  using Cfg = TcCfg<PASSES, BN>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* stage_base = smem;                                   // STAGES * STAGE_BYTES, 1024-aligned
  float* lval = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);   // [kc][128]
  int* lidx = reinterpret_cast<int*>(lval + (size_t)kc * TC_BM);              // [kc][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(lidx + (size_t)kc * TC_BM);
  uint64_t* full_bar = bars;                   // [STAGES]
  uint64_t* empty_bar = bars + STAGES;         // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;     // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * TC_BM;
  const int split = blockIdx.y;
  const int tile_begin = split * tiles_per_split;
  const int tile_end = min(tiles_total, tile_begin + tiles_per_split);
  const int ntiles = tile_end - tile_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&tfull_bar[b]), 1); mbar_init(smem_u32(&tempty_bar[b]), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qhi) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dhi) : "memory");
      int it = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int db0 = (tile_begin + t) * BN;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
          const uint32_t fb = smem_u32(&full_bar[s]);
          mbar_expect_tx(fb, (uint32_t)Cfg::STAGE_BYTES);
          unsigned char* st = stage_base + (size_t)s * Cfg::STAGE_BYTES;
          const uint32_t a_hi = smem_u32(st);
          const uint32_t b_hi = a_hi + Cfg::PLANES * TC_A_PLANE;
          tma_load_2d(a_hi, &map_qhi, fb, kb * TC_BK, q0);
          tma_load_2d(b_hi, &map_dhi, fb, kb * TC_BK, db0);
          if (PASSES == 3) {
            tma_load_2d(a_hi + TC_A_PLANE, &map_qlo, fb, kb * TC_BK, q0);
            tma_load_2d(b_hi + Cfg::B_PLANE, &map_dlo, fb, kb * TC_BK, db0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(TC_BM, BN);
      int it = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        const uint32_t tph = (uint32_t)(t >> 1) & 1u;
        mbar_wait(smem_u32(&tempty_bar[buf]), tph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(smem_u32(&full_bar[s]), ph);
          tc_fence_after();
          unsigned char* st = stage_base + (size_t)s * Cfg::STAGE_BYTES;
          const uint32_t a_hi = smem_u32(st);
          const uint32_t a_lo = a_hi + TC_A_PLANE;
          const uint32_t b_hi = a_hi + Cfg::PLANES * TC_A_PLANE;
          const uint32_t b_lo = b_hi + Cfg::B_PLANE;
          // small terms first, then hi.hi
#pragma unroll
          for (int p = 0; p < PASSES; ++p) {
            const uint32_t aa = (PASSES == 3 && p == 0) ? a_lo : a_hi;
            const uint32_t bb = (PASSES == 3 && p == 1) ? b_lo : b_hi;
            const uint64_t ad = make_kmajor_sw128_desc(aa);
            const uint64_t bd = make_kmajor_sw128_desc(bb);
#pragma unroll
            for (int k4 = 0; k4 < TC_BK / TC_UMMA_K; ++k4) {
              // advance 32 B (= 8 tf32) inside the 128-B swizzle atom: +2 in 16-B units
              tc_mma_tf32(d_tmem, ad + (uint64_t)(k4 * 2), bd + (uint64_t)(k4 * 2), idesc,
                          (kb | p | k4) != 0 ? 1u : 0u);
            }
          }
          tc_commit(smem_u32(&empty_bar[s]));      // frees the smem stage when these MMAs retire
        }
        tc_commit(smem_u32(&tfull_bar[buf]));      // accumulator ready for the epilogue
      }
    }
  } else {
    // ===================== epilogue: thread <-> query row =====================
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may read
    const int r_in_tile = quarter * 32 + lane;
    const bool row_ok = q0 + r_in_tile < nq;
    float* my_val = lval + r_in_tile;
    int* my_idx = lidx + r_in_tile;
    const uint32_t my_val_s = smem_addr(my_val), my_idx_s = smem_addr(my_idx);
    ListState st = list_init();
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      const uint32_t tph = (uint32_t)(t >> 1) & 1u;
      mbar_wait(smem_u32(&tfull_bar[buf]), tph);
      tc_fence_after();
      const int db0 = (tile_begin + t) * BN;
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN);
      const bool partial = db0 + BN > ndb;          // only the last db tile has zero-filled columns
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        float r[32];
        tc_ld32(taddr0 + (uint32_t)(ch * 32), r);
        tc_wait_ld();
        if (ch == BN / 32 - 1) {
          // all of this buffer's columns are now in registers (or consumed): hand it back
          tc_fence_before();
          if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[buf]));
        }
        const int jb = db0 + ch * 32;
        if (partial) {
#pragma unroll
          for (int c = 0; c < 32; ++c) r[c] = (jb + c < ndb) ? r[c] : -INFINITY;
        }
        float gm[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float m01 = fmaxf(r[g * 8 + 0], r[g * 8 + 1]), m23 = fmaxf(r[g * 8 + 2], r[g * 8 + 3]);
          float m45 = fmaxf(r[g * 8 + 4], r[g * 8 + 5]), m67 = fmaxf(r[g * 8 + 6], r[g * 8 + 7]);
          gm[g] = fmaxf(fmaxf(m01, m23), fmaxf(m45, m67));
        }
        const float mx = fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3]));
        if (row_ok && mx > st.thr) {
          // rare path, register-resident: per group of 8 columns, repeatedly pick the first column (index
          // order) that still beats the running threshold and push it into the heap
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (gm[g] > st.thr) {
              int last = -1;
              while (true) {
                float cv = -INFINITY;
                int cc = -1;
#pragma unroll
                for (int c = 7; c >= 0; --c) {
                  const bool p = (c > last) && (r[g * 8 + c] > st.thr);
                  cv = p ? r[g * 8 + c] : cv;
                  cc = p ? c : cc;
                }
                if (cc < 0) break;
                list_push(my_val_s, my_idx_s, TC_BM * 4, kc, st, cv, jb + g * 8 + cc);
                last = cc;
              }
            }
          }
        }
      }
    }
    if (row_ok) {
      const long long base = ((long long)split * nq + (q0 + r_in_tile)) * kc;
      for (int s = 0; s < kc; ++s) {
        const bool f = s < st.cnt;
        cand_val[base + s] = f ? my_val[s * TC_BM] : -INFINITY;
        cand_idx[base + s] = f ? my_idx[s * TC_BM] : -1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host
// Row-major [rows, dpad] fp32, box = [box_rows, 32 features], 128-B swizzle, zero fill out of bounds.
static int make_map(CUtensorMap* m, const float* base, long long rows, int dpad, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return BGNN_ERR_DRIVER;
  cuuint64_t dims[2] = {(cuuint64_t)dpad, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)dpad * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? BGNN_OK : BGNN_ERR_DRIVER;
}

// Work decomposition.  Lists of kc > k nominees per (row, split) live in shared memory next to the
// operand ring, so kc, the db tile width and the ring depth are traded against each other here.
TcPlan tc_plan(int nq, int ndb, int d, int k, int passes) {
  (void)d;
  TcPlan p;
  p.bn = 0;
  int kc = (passes == 3) ? (k + 4) : max(k + 12, (3 * k) / 2 + 2);
  kc = (kc + 3) / 4 * 4;
  p.kc = kc;
  const int planes = (passes == 3) ? 2 : 1;
  const int list_bytes = kc * TC_BM * 8;
  for (int bn = 256; bn >= 128; bn >>= 1) {
    const int stage_bytes = planes * (TC_A_PLANE + bn * TC_BK * 4);
    const int stages = (TC_SMEM_MAX - TC_SMEM_FIXED - list_bytes) / stage_bytes;
    if (stages >= 2) { p.bn = bn; p.stages = min(stages, 8); break; }
  }
  if (p.bn == 0 || kc > BGNN_MERGE_MAX_CAND) { p.bn = 0; return p; }   // caller falls back to the CUDA-core sweep
  const int tiles = (ndb + p.bn - 1) / p.bn;
  const int qblocks = (nq + TC_BM - 1) / TC_BM;
  int ns = (2 * kNumSMs + qblocks - 1) / qblocks;      // aim for >= 2 CTAs per SM when nq is small
  ns = max(1, min(min(ns, tiles), min(8, BGNN_MERGE_MAX_CAND / kc)));
  p.tiles_per_split = (tiles + ns - 1) / ns;
  p.nsplit = (tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.nlists = p.nsplit;
  p.pair = 0;
  p.kps = 1;
  return p;
}

template <int PASSES, int BN>
static int launch_cfg(const float* qhi, const float* qlo, int nq, const float* dhi, const float* dlo, int ndb, int dpad,
                      const TcPlan& plan, float* cand_val, int* cand_idx, cudaStream_t stream) {
  using Cfg = TcCfg<PASSES, BN>;
  CUtensorMap mq_hi, mq_lo, md_hi, md_lo;
  int rc;
  if ((rc = make_map(&mq_hi, qhi, nq, dpad, TC_BM)) != BGNN_OK) return rc;
  if ((rc = make_map(&md_hi, dhi, ndb, dpad, BN)) != BGNN_OK) return rc;
  if (PASSES == 3) {
    if ((rc = make_map(&mq_lo, qlo, nq, dpad, TC_BM)) != BGNN_OK) return rc;
    if ((rc = make_map(&md_lo, dlo, ndb, dpad, BN)) != BGNN_OK) return rc;
  } else {
    mq_lo = mq_hi; md_lo = md_hi;
  }
  const size_t smem = TC_SMEM_FIXED + (size_t)plan.stages * Cfg::STAGE_BYTES + (size_t)plan.kc * TC_BM * 8;
  if (smem > (size_t)TC_SMEM_MAX || plan.stages < 2) return BGNN_ERR_UNSUPPORTED;
  auto kern = knn_cosine_tc_kernel<PASSES, BN>;
  BGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = (ndb + BN - 1) / BN;
  dim3 grid((nq + TC_BM - 1) / TC_BM, plan.nsplit);
  kern<<<grid, TC_THREADS, smem, stream>>>(mq_hi, mq_lo, md_hi, md_lo, nq, ndb, dpad / TC_BK, tiles,
                                           plan.tiles_per_split, plan.kc, plan.stages, cand_val, cand_idx);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

int launch_knn_cosine_tc(const float* qhi, const float* qlo, int nq, const float* dhi, const float* dlo, int ndb, int d,
                         int passes, const TcPlan& plan, float* cand_val, int* cand_idx, cudaStream_t stream) {
  if (nq <= 0) return BGNN_OK;
  if (d % TC_BK != 0) return BGNN_ERR_INVALID_ARG;   // caller pads to a multiple of 32
  const int bn = plan.bn;
  if (bn != 128 && bn != 256) return BGNN_ERR_UNSUPPORTED;
  if (passes == 3) {
    return bn == 256 ? launch_cfg<3, 256>(qhi, qlo, nq, dhi, dlo, ndb, d, plan, cand_val, cand_idx, stream)
                     : launch_cfg<3, 128>(qhi, qlo, nq, dhi, dlo, ndb, d, plan, cand_val, cand_idx, stream);
  }
  return bn == 256 ? launch_cfg<1, 256>(qhi, qlo, nq, dhi, dlo, ndb, d, plan, cand_val, cand_idx, stream)
                   : launch_cfg<1, 128>(qhi, qlo, nq, dhi, dlo, ndb, d, plan, cand_val, cand_idx, stream);
}

}  // namespace bgnn
