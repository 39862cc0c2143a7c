// Cosine-similarity kNN sweep, single-pass fp16 tensor-core nomination (sm_100a):
//   TMA -> shared memory -> tcgen05.mma kind::f16 (fp32 accumulators in TMEM) -> tcgen05.ld ->
//   fused per-row top-KC selection.  The [nq, ndb] score matrix never leaves the SM.
//
// Replaces the pair-materialising similarity + sim_mat.topk of the reference for the cosine head
// (models/models.py:124-130 / 945-948, main_bridged_graph.py:45-67, 90-111).
//
// Exactness.  The sweep scores unit rows rounded to fp16 (10 explicit mantissa bits, the same as tf32,
// at half the operand bytes and twice the MMA rate).  For unit rows a, b with fp16 roundings ha, hb:
//     |a.b - ha.hb| <= 2^-10 sum|a_i||b_i| + 2^-25 (sum|a_i| + sum|b_i|) + O(2^-22)  <=  2^-10 + 2^-24 sqrt(d)
// so the approximate score is within delta_f16(d) of the exact fp32 cosine.  The kernel only NOMINATES
// candidates (two lists of KC per row: one per epilogue warp of a lane quarter); knn_select.cu re-scores
// them exactly in fp32 with the reference fmaf chain, selects under the parity key and certifies the row
// against (largest discarded approximate score + delta); uncertified rows are redone exactly on CUDA cores.
//
// CTA = 10 warps, one (128-query block, db split) work unit, 1 CTA / SM:
//   warp 0     TMA producer: the query block A [128 x dpad] once (resident), then a ring of B k-blocks
//              [BN x 64] fp16, SWIZZLE_128B
//   warp 1     TMEM allocator + single-thread tcgen05.mma issuer; 128 x BN fp32 accumulator,
//              double-buffered in TMEM: the epilogue of tile t overlaps the MMAs of tile t+1
//   warps 2-9  epilogue: two warps per TMEM lane quarter, each taking every other 32-column chunk;
//              thread <-> query row, running-threshold test on registers, rare insertion into a
//              thread-private list in shared memory
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace bgnn {

constexpr int F16_BM = 128;
constexpr int F16_BK = 64;                          // fp16 elements per k-block = 128 B = one swizzle row
constexpr int F16_UMMA_K = 16;
constexpr int F16_THREADS = 320;
constexpr int F16_A_KBLOCK = F16_BM * F16_BK * 2;   // 16 KB
constexpr int F16_SMEM_MAX = 232448;
constexpr int F16_SMEM_FIXED = 1024 + 512 + 1024;   // alignment slack + barriers / tmem slot + shared thresholds
constexpr int F16_FCAP = 4;                         // pending candidates per (row, epilogue warp) before a drain
constexpr int F16_FIFO_BYTES = 2 * F16_FCAP * F16_BM * 8;
constexpr int F16_DRAIN_TILES = 8;                  // all lanes drain together every so many tiles

__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ float max3f(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) [4,6), a/b format F16 (0)
// [7,10)/[10,13), both K-major, N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// x -> unit row in fp32 (row stride ld, zero padded) and its fp16 rounding (row stride ldh, zero padded)
__global__ void __launch_bounds__(256)
normalize_f16_kernel(const float* __restrict__ x, long long n, int d, int ld, int ldh, int normalize,
                     float* __restrict__ xn, __half* __restrict__ xh) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const float* p = x + row * d;
  float denom = 1.f;
  if (normalize) {
    float ss = 0.f;
    for (int c = lane; c < d; c += 32) { float v = __ldg(p + c); ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    denom = fmaxf(sqrtf(ss), 1e-8f);
  }
  for (int c = lane; c < ldh; c += 32) {
    float v = 0.f;
    if (c < d) { v = __ldg(p + c); if (normalize) v = v / denom; }
    if (c < ld) xn[row * ld + c] = v;
    xh[row * ldh + c] = __float2half_rn(v);
  }
}

int launch_normalize_f16(const float* x, long long n, int d, int ld, int ldh, int normalize, float* xn, void* xh,
                         cudaStream_t stream) {
  if (n <= 0) return BGNN_OK;
  long long blocks = (n + 7) / 8;
  normalize_f16_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, n, d, ld, ldh, normalize, xn, (__half*)xh);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

template <int BN>
__global__ void __launch_bounds__(F16_THREADS, 1)
knn_cosine_f16_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_db,
                      int nq, int ndb, int kblocks, int tiles_total, int tiles_per_split, int kc, int stages,
                      const float* __restrict__ thr_init, float* __restrict__ cand_val, int* __restrict__ cand_idx,
                      int dbg) {
  constexpr int B_STAGE = BN * F16_BK * 2;
  constexpr int NBUF = 512 / BN;                 // accumulator buffers in TMEM: 2 x 256 or 4 x 128 columns
  constexpr int TMEM_COLS = 512;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* a_base = smem;                                          // kblocks * 16 KB, resident
  unsigned char* b_base = a_base + (size_t)kblocks * F16_A_KBLOCK;       // stages * B_STAGE ring
  float* lval = reinterpret_cast<float*>(b_base + (size_t)stages * B_STAGE);   // [2][kc][128]
  int* lidx = reinterpret_cast<int*>(lval + (size_t)2 * kc * F16_BM);           // [2][kc][128]
  float* ffv = reinterpret_cast<float*>(lidx + (size_t)2 * kc * F16_BM);     // [2][FCAP][128] pending values
  int* ffi = reinterpret_cast<int*>(ffv + 2 * F16_FCAP * F16_BM);             // [2][FCAP][128] pending indices
  float* thr_sh = reinterpret_cast<float*>(ffi + 2 * F16_FCAP * F16_BM);      // [2][128] list thresholds, shared by the warp pair
  uint64_t* bars = reinterpret_cast<uint64_t*>(thr_sh + 2 * F16_BM);
  uint64_t* full_bar = bars;                      // [stages]
  uint64_t* empty_bar = bars + stages;            // [stages]
  uint64_t* tfull_bar = bars + 2 * stages;               // [NBUF]
  uint64_t* tempty_bar = bars + 2 * stages + NBUF;       // [NBUF]
  uint64_t* a_bar = bars + 2 * stages + 2 * NBUF;        // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 2 * NBUF + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * F16_BM;
  const int split = blockIdx.y;
  const int tile_begin = split * tiles_per_split;
  const int tile_end = min(tiles_total, tile_begin + tiles_per_split);
  const int ntiles = tile_end - tile_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int b = 0; b < NBUF; ++b) { mbar_init(smem_u32(&tfull_bar[b]), 1); mbar_init(smem_u32(&tempty_bar[b]), 8); }
    mbar_init(smem_u32(a_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_db) : "memory");
      const uint32_t ab = smem_u32(a_bar);
      mbar_expect_tx(ab, (uint32_t)(kblocks * F16_A_KBLOCK));
      for (int kb = 0; kb < kblocks; ++kb)
        tma_load_2d(smem_u32(a_base + (size_t)kb * F16_A_KBLOCK), &map_q, ab, kb * F16_BK, q0);
      int it = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int db0 = (tile_begin + t) * BN;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % stages;
          const uint32_t ph = (uint32_t)(it / stages) & 1u;
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
          const uint32_t fb = smem_u32(&full_bar[s]);
          mbar_expect_tx(fb, (uint32_t)B_STAGE);
          tma_load_2d(smem_u32(b_base + (size_t)s * B_STAGE), &map_db, fb, kb * F16_BK, db0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(F16_BM, BN);
      mbar_wait(smem_u32(a_bar), 0u);
      tc_fence_after();
      int it = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int buf = t % NBUF;
        const uint32_t tph = (uint32_t)(t / NBUF) & 1u;
        mbar_wait(smem_u32(&tempty_bar[buf]), tph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % stages;
          const uint32_t ph = (uint32_t)(it / stages) & 1u;
          mbar_wait(smem_u32(&full_bar[s]), ph);
          tc_fence_after();
          const uint64_t ad = make_kmajor_sw128_desc(smem_u32(a_base + (size_t)kb * F16_A_KBLOCK));
          const uint64_t bd = make_kmajor_sw128_desc(smem_u32(b_base + (size_t)s * B_STAGE));
#pragma unroll
          for (int k4 = 0; k4 < F16_BK / F16_UMMA_K; ++k4) {
            // advance 32 B (= 16 fp16) inside the 128-B swizzle row: +2 in 16-B units
            if (!(dbg & 4)) tc_mma_f16(d_tmem, ad + (uint64_t)(k4 * 2), bd + (uint64_t)(k4 * 2), idesc, (kb | k4) != 0 ? 1u : 0u);
          }
          tc_commit(smem_u32(&empty_bar[s]));      // frees the ring slot when these MMAs retire
        }
        tc_commit(smem_u32(&tfull_bar[buf]));      // accumulator ready for the epilogue
      }
    }
  } else {
    // ===================== epilogue: thread <-> query row, warp pair <-> lane quarter =====================
    const int quarter = warp & 3;                  // TMEM lanes this warp may read: 32*quarter ..
    const int half = (warp - 2) >> 2;              // which of the two warps of the quarter
    const int r_in_tile = quarter * 32 + lane;
    const bool row_ok = q0 + r_in_tile < nq;
    float* my_val = lval + (size_t)half * kc * F16_BM + r_in_tile;
    int* my_idx = lidx + (size_t)half * kc * F16_BM + r_in_tile;
    const uint32_t my_val_s = smem_addr(my_val), my_idx_s = smem_addr(my_idx);
    // The two warps of a lane quarter keep separate lists for the same rows but share their thresholds:
    // a column is kept only if it beats max(own, partner) threshold.  Sound for the certification in
    // knn_select.cu: every column either warp discards scores <= the larger of the two final list minima.
    const uint32_t my_thr_s = smem_addr(thr_sh + half * F16_BM + r_in_tile);
    const uint32_t other_thr_s = smem_addr(thr_sh + (half ^ 1) * F16_BM + r_in_tile);
    sts_f32(my_thr_s, -INFINITY);
    asm volatile("bar.sync 1, 256;" ::: "memory");     // epilogue warps only
    ListState st = list_init();
    // effective threshold = max(own list, partner list, seed): the seed is a score that at least kc_seed db
    // rows of a sample reach (knn_seed_thr_kernel), so the lists skip most of their start-up insertions
    float thr = (thr_init && row_ok) ? __ldg(thr_init + q0 + r_in_tile) : -INFINITY;
    bool partial = false;
    // Heap insertions are deferred: a column that beats the threshold is appended to a small per-thread
    // FIFO (two stores), and the FIFOs are drained into the heaps by ALL lanes of the warp together every
    // F16_DRAIN_TILES tiles (or by one lane alone when its FIFO is full).  In steady state a chunk holds a
    // candidate for one or two of the 32 rows only, so inline insertion ran the long heap code at 1/32 lane
    // occupancy (profiles/r01g); drained together it is shared by every row that has something pending.
    // The threshold a lane filters with is then slightly stale, which only lets a few more columns through.
    const uint32_t my_fv_s = smem_addr(ffv + (size_t)half * F16_FCAP * F16_BM + r_in_tile);
    const uint32_t my_fi_s = smem_addr(ffi + (size_t)half * F16_FCAP * F16_BM + r_in_tile);
    int fcnt = 0;
    auto drain = [&]() {
      for (int s = 0; s < fcnt; ++s) {
        const float v = lds_f32(my_fv_s + s * (F16_BM * 4));
        const int j = lds_s32(my_fi_s + s * (F16_BM * 4));
        if (v > st.thr) list_push(my_val_s, my_idx_s, F16_BM * 4, kc, st, v, j);   // FIFO order = index order
      }
      fcnt = 0;
      thr = fmaxf(thr, st.thr);
      sts_f32(my_thr_s, st.thr);
    };
    // one 32-column chunk held in registers: group maxima vs the running threshold; the rare path stays in
    // registers too -- per group of 8 columns, pick the first column (index order) that beats the threshold,
    // queue it, and rescan only if the group held more than one candidate
    auto process_chunk = [&](float (&r)[32], int jb) {
      if (partial) {
#pragma unroll
        for (int c = 0; c < 32; ++c) r[c] = (jb + c < ndb) ? r[c] : -INFINITY;
      }
      float gm[4];                                  // 3-input maxima (FMNMX3): 4 instructions per group of 8
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float m012 = max3f(r[g * 8 + 0], r[g * 8 + 1], r[g * 8 + 2]);
        const float m345 = max3f(r[g * 8 + 3], r[g * 8 + 4], r[g * 8 + 5]);
        gm[g] = max3f(m012, m345, fmaxf(r[g * 8 + 6], r[g * 8 + 7]));
      }
      const float mx = max3f(gm[0], gm[1], fmaxf(gm[2], gm[3]));
      if (row_ok && mx > thr) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (gm[g] > thr) {
            int last = -1;
            while (true) {
              float cv = -INFINITY;
              int cc = -1, cnt = 0;
#pragma unroll
              for (int c = 7; c >= 0; --c) {
                const bool p = (c > last) && (r[g * 8 + c] > thr);
                cv = p ? r[g * 8 + c] : cv;
                cc = p ? c : cc;
                cnt += p ? 1 : 0;
              }
              if (cc < 0) break;
              if (fcnt == F16_FCAP) drain();
              if (cv > thr) {                       // the drain may have raised the threshold
                sts_f32(my_fv_s + fcnt * (F16_BM * 4), cv);
                sts_s32(my_fi_s + fcnt * (F16_BM * 4), jb + g * 8 + cc);
                ++fcnt;
              }
              if (cnt == 1) break;
              last = cc;
            }
          }
        }
      }
    };
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t % NBUF;
      const uint32_t tph = (uint32_t)(t / NBUF) & 1u;
      if ((t % F16_DRAIN_TILES) == F16_DRAIN_TILES - 1 && __any_sync(0xffffffffu, fcnt > 0)) drain();
      mbar_wait(smem_u32(&tfull_bar[buf]), tph);
      tc_fence_after();
      const int db0 = (tile_begin + t) * BN;
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN);
      partial = db0 + BN > ndb;                     // only the last db tile has zero-filled columns
      if (dbg & 2) {                                // bottleneck experiments: hand the buffer straight back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[buf]));
        continue;
      }
      // Software pipeline over this warp's chunks (half, half+2, ...): the tcgen05.ld of the next chunk is
      // in flight while the current one is reduced.  The TMEM buffer goes back to the MMA issuer as soon as
      // the last chunk has landed in registers.
      constexpr int NCH = BN / 64;                  // chunks per warp per tile (even)
      float ra[32], rb[32];
      tc_ld32(taddr0 + (uint32_t)(half * 32), ra);
#pragma unroll 1
      for (int i = 0; i < NCH; i += 2) {
        const int ch_a = half + 2 * i, ch_b = ch_a + 2;
        tc_wait_ld();
        tc_ld32(taddr0 + (uint32_t)(ch_b * 32), rb);
        thr = fmaxf(thr, lds_f32(other_thr_s));
        if (!(dbg & 1)) process_chunk(ra, db0 + ch_a * 32);
        tc_wait_ld();
        if (i + 2 < NCH) {
          tc_ld32(taddr0 + (uint32_t)((ch_b + 2) * 32), ra);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[buf]));
        }
        thr = fmaxf(thr, lds_f32(other_thr_s));
        if (!(dbg & 1)) process_chunk(rb, db0 + ch_b * 32);
      }
    }
    drain();
    if (row_ok) {
      const long long base = (((long long)split * 2 + half) * nq + (q0 + r_in_tile)) * kc;
      for (int s = 0; s < kc; ++s) {
        const bool f = s < st.cnt;
        cand_val[base + s] = f ? my_val[s * F16_BM] : -INFINITY;
        cand_idx[base + s] = f ? my_idx[s * F16_BM] : -1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host
// Row-major [rows, ldh] fp16, box = [box_rows, 64 features], 128-B swizzle, zero fill out of bounds.
static int make_map_f16(CUtensorMap* m, const void* base, long long rows, int ldh, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return BGNN_ERR_DRIVER;
  cuuint64_t dims[2] = {(cuuint64_t)ldh, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ldh * 2};
  cuuint32_t box[2] = {(cuuint32_t)F16_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? BGNN_OK : BGNN_ERR_DRIVER;
}

// Work decomposition: kc nominees per (row, half-list) next to the resident query block and the B ring.
TcPlan tc_plan_f16(int nq, int ndb, int d, int k, int kc_fixed) {
  TcPlan p;
  p.bn = 0;
  const int ldh = (d + F16_BK - 1) / F16_BK * F16_BK;
  int kc = kc_fixed > 0 ? kc_fixed : (k + 4 + 3) / 4 * 4;
  p.kc = kc;
  const int a_bytes = (ldh / F16_BK) * F16_A_KBLOCK;
  const int list_bytes = 2 * kc * F16_BM * 8 + F16_FIFO_BYTES;
  static const int force_bn = getenv("BGNN_F16_BN") ? atoi(getenv("BGNN_F16_BN")) : 0;   // tuning experiments
  for (int bn = (force_bn == 128 ? 128 : 256); bn >= 128; bn >>= 1) {
    const int stage_bytes = bn * F16_BK * 2;
    const int stages = (F16_SMEM_MAX - F16_SMEM_FIXED - a_bytes - list_bytes) / stage_bytes;
    if (stages >= 3) { p.bn = bn; p.stages = min(stages, 8); break; }
  }
  if (p.bn == 0) return p;                                  // caller falls back to another sweep
  const int tiles = (ndb + p.bn - 1) / p.bn;
  const int qblocks = (nq + F16_BM - 1) / F16_BM;
  int ns = (2 * kNumSMs + qblocks - 1) / qblocks;            // fill the machine when nq is small
  ns = max(1, min(min(ns, tiles), min(8, BGNN_MERGE_MAX_CAND / (2 * kc))));
  if (ns < 1) { p.bn = 0; return p; }
  p.tiles_per_split = (tiles + ns - 1) / ns;
  p.nsplit = (tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.nlists = 2 * p.nsplit;
  if (p.nlists * kc > BGNN_MERGE_MAX_CAND) p.bn = 0;
  return p;
}

template <int BN>
static int launch_f16_cfg(const void* qh, int nq, const void* dh, int ndb, int ldh, const TcPlan& plan,
                          const float* thr_init, float* cand_val, int* cand_idx, cudaStream_t stream) {
  CUtensorMap mq, md;
  int rc;
  if ((rc = make_map_f16(&mq, qh, nq, ldh, F16_BM)) != BGNN_OK) return rc;
  if ((rc = make_map_f16(&md, dh, ndb, ldh, BN)) != BGNN_OK) return rc;
  const int kblocks = ldh / F16_BK;
  const size_t smem = F16_SMEM_FIXED + (size_t)kblocks * F16_A_KBLOCK + (size_t)plan.stages * BN * F16_BK * 2 +
                      (size_t)2 * plan.kc * F16_BM * 8 + F16_FIFO_BYTES;
  if (smem > (size_t)F16_SMEM_MAX || plan.stages < 2) return BGNN_ERR_UNSUPPORTED;
  auto kern = knn_cosine_f16_kernel<BN>;
  BGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = (ndb + BN - 1) / BN;
  dim3 grid((nq + F16_BM - 1) / F16_BM, plan.nsplit);
  static const int dbg = getenv("BGNN_F16_DBG") ? atoi(getenv("BGNN_F16_DBG")) : 0;   // bottleneck experiments only
  kern<<<grid, F16_THREADS, smem, stream>>>(mq, md, nq, ndb, kblocks, tiles, plan.tiles_per_split, plan.kc, plan.stages,
                                            thr_init, cand_val, cand_idx, dbg);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

int launch_knn_cosine_f16(const void* qh, int nq, const void* dh, int ndb, int ldh, const TcPlan& plan,
                          const float* thr_init, float* cand_val, int* cand_idx, cudaStream_t stream) {
  if (nq <= 0) return BGNN_OK;
  if (ldh % F16_BK != 0) return BGNN_ERR_INVALID_ARG;
  if (plan.bn == 256) return launch_f16_cfg<256>(qh, nq, dh, ndb, ldh, plan, thr_init, cand_val, cand_idx, stream);
  if (plan.bn == 128) return launch_f16_cfg<128>(qh, nq, dh, ndb, ldh, plan, thr_init, cand_val, cand_idx, stream);
  return BGNN_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------ threshold seeding
// A streaming top-KC list performs ~KC ln(n/KC) insertions, most of them while its threshold is still loose.
// Seeding: sweep a strided SAMPLE of the db first (same kernel, tiny lists), take per query row a score that at
// least kSeedKc sampled rows reach, and start the full sweep from that threshold.  Soundness is unchanged: the
// merge kernel treats the seed as one more "largest discarded score" bound and certifies or falls back.
int knn_seed_rows(int ndb, int k) {
  static const int off = getenv("BGNN_F16_NOSEED") ? atoi(getenv("BGNN_F16_NOSEED")) : 0;   // tuning experiments
  if (off) return 0;
  // expected number of db rows above the seed >= (ndb / S) * kSeedKc; keep it >= 16 (k + 4) so that rows with
  // fewer than k + 1 survivors (-> exact fallback) stay below ~2e-4
  long long s = ndb / (4ll * (k + 4));
  const long long cap = ndb / 256 > 8192 ? ndb / 256 : 8192;
  if (s > cap) s = cap;
  s = s / 256 * 256;
  return s >= 1024 ? (int)s : 0;
}

__global__ void __launch_bounds__(256)
gather_sample_f16_kernel(const uint4* __restrict__ src, long long ndb, int row_vecs, int srows, uint4* __restrict__ dst) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)srows * row_vecs) return;
  const long long i = t / row_vecs, v = t % row_vecs;
  const long long r = i * ndb / srows;
  dst[t] = __ldg(src + r * row_vecs + v);
}

// seed[row] = max over the FULL lists of the row of the list minimum (each full list holds kc sampled rows that
// score at least its minimum); -inf when no list is full.
__global__ void __launch_bounds__(256)
knn_seed_thr_kernel(const float* __restrict__ cand_val, const int* __restrict__ cand_idx, int nlists, int kc, int nq,
                    float* __restrict__ seed) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nq) return;
  float best = -INFINITY;
  for (int l = 0; l < nlists; ++l) {
    const long long base = ((long long)l * nq + row) * kc;
    float lmin = INFINITY;
    bool full = true;
    for (int s = 0; s < kc; ++s) {
      if (cand_idx[base + s] < 0) full = false;
      lmin = fminf(lmin, cand_val[base + s]);
    }
    if (full) best = fmaxf(best, lmin);
  }
  seed[row] = best;
}

int launch_knn_seed_f16(const void* qh, int nq, const void* dh, int ndb, int ldh, int srows, const TcPlan& seed_plan,
                        void* sample, float* cand_val, int* cand_idx, float* seed, cudaStream_t stream) {
  if (nq <= 0 || srows <= 0) return BGNN_OK;
  const int row_vecs = ldh * 2 / 16;
  const long long total = (long long)srows * row_vecs;
  gather_sample_f16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(dh), ndb, row_vecs, srows, reinterpret_cast<uint4*>(sample));
  BGNN_LAUNCH_CHECK();
  int rc = launch_knn_cosine_f16(qh, nq, sample, srows, ldh, seed_plan, nullptr, cand_val, cand_idx, stream);
  if (rc != BGNN_OK) return rc;
  knn_seed_thr_kernel<<<(nq + 255) / 256, 256, 0, stream>>>(cand_val, cand_idx, seed_plan.nlists, seed_plan.kc, nq, seed);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

}  // namespace bgnn
