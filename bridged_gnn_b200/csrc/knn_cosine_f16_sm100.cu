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
// CTA = 2 + 4*EW warps (EW = epilogue warps per TMEM lane quarter: 4, or 2 for CTA pairs / the seed sweep), one
// (128-query block, db split) work unit, 1 CTA / SM:
//   warp 0     TMA producer: the query block A [128 x dpad] once (resident), then a ring of B k-blocks
//              [BN x 64] fp16, SWIZZLE_128B
//   warp 1     TMEM allocator + single-thread tcgen05.mma issuer; 128 x BN fp32 accumulator,
//              double-buffered in TMEM: the epilogue of tile t overlaps the MMAs of tile t+1
//   warps 2..   epilogue: EW warps per TMEM lane quarter, each taking every EW-th 32-column chunk;
//              thread <-> query row, running-threshold test on registers, rare insertion into a
//              thread-private list in shared memory (one list per row and epilogue warp of the quarter)
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace bgnn {

constexpr int F16_BM = 128;
constexpr int F16_BK = 64;                          // fp16 elements per k-block = 128 B = one swizzle row
constexpr int F16_UMMA_K = 16;
constexpr int f16_threads(int ew) { return (2 + 4 * ew) * 32; }   // producer + MMA issuer + 4*EW epilogue warps
constexpr int F16_A_KBLOCK = F16_BM * F16_BK * 2;   // 16 KB
constexpr int F16_SMEM_MAX = 232448;
constexpr int f16_smem_fixed(int ew) { return 1024 + 512 + ew * F16_BM * 4; }   // alignment slack + barriers / tmem slot + shared thresholds
constexpr int F16_FCAP = 4;                         // pending candidates per (row, epilogue warp) before a drain
// EW = 2: the pending FIFOs live in shared memory; EW = 4: two pending candidates per thread in registers (the
// shared memory goes to the four lists per row)
constexpr int f16_fifo_bytes(int ew) { return ew == 2 ? 2 * F16_FCAP * F16_BM * 8 : 0; }
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

__device__ __forceinline__ void tc_mma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
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

// ---- epilogue slow path, out of line -------------------------------------------------------------------
// The hot loop of an epilogue warp is the max tree of a chunk; everything a threshold hit needs (finding the
// columns, the pending FIFO, the heap) lives in two __noinline__ functions so that the unrolled per-tile loop
// stays a few hundred instructions (four inlined copies of it thrashed the instruction cache).
struct EpiState {
  int cnt;      // heap fill
  int fcnt;     // pending FIFO fill
  float lthr;   // worst kept heap value (-inf until the heap is full)
  float thr;    // threshold the thread filters with = max(own heap, partner heaps, seed), slightly stale
  float f0v, f1v;   // register FIFO (REGF): up to two pending candidates
  int f0i, f1i;
};
struct EpiAddr {                // shared-window byte addresses of this thread's columns
  uint32_t val, idx;            // heap [kc][128]
  uint32_t fv, fi;              // FIFO [F16_FCAP][128]
  uint32_t thr;                 // this warp's published heap threshold of the row
  int kc;
};

// Nominee heap: a binary min-heap on the APPROXIMATE score alone (root = worst kept nominee).  The shared list_push of
// common.cuh orders by the parity key (value desc, index asc), which the exact CUDA-core sweep needs; here the lists only
// nominate -- the threshold test is strict, an evicted nominee scores no more than the list's new minimum either way, and
// knn_select.cu re-scores and orders exactly -- so the index compare, and with it most of the dependent instructions of a
// sift level, can go: at one or two active lanes a level cost ~185 cycles (2300 cycles per drain, profiles/r02w).
// Both children's values AND indices are loaded together (one shared-memory round trip per level), and the new root is
// tracked in a register instead of being read back.  An unsorted list with a tracked minimum (two stores + one pass
// over the kc values per insertion) was measured slower than a heap (kc / 4 dependent rounds of loads, r02x).
__device__ __forceinline__ void nominee_push(uint32_t val, uint32_t idx, int kc, ListState& st, float v, int j) {
  constexpr uint32_t SB = F16_BM * 4;               // byte stride between the slots of a thread's column
  int pos;
  float root = v;
  if (st.cnt < kc) {
    pos = st.cnt++;
    while (pos > 0) {                               // sift up: a parent must not score more than its children
      const int par = (pos - 1) >> 1;
      const float pv = lds_f32(val + par * SB);
      const int pi = lds_s32(idx + par * SB);
      if (!(v < pv)) break;
      sts_f32(val + pos * SB, pv);
      sts_s32(idx + pos * SB, pi);
      pos = par;
    }
    sts_f32(val + pos * SB, v);
    sts_s32(idx + pos * SB, j);
    if (st.cnt == kc) st.thr = lds_f32(val);
    return;
  }
  pos = 0;                                          // replace the root, sift down
  while (true) {
    const int l = 2 * pos + 1;
    if (l >= kc) break;
    const int r = min(l + 1, kc - 1);               // == l when there is no right child
    const float lv = lds_f32(val + l * SB), rv = lds_f32(val + r * SB);
    const int li = lds_s32(idx + l * SB), ri = lds_s32(idx + r * SB);
    const bool right = rv < lv;
    const float cv = right ? rv : lv;
    if (!(cv < v)) break;
    if (pos == 0) root = cv;
    sts_f32(val + pos * SB, cv);
    sts_s32(idx + pos * SB, right ? ri : li);
    pos = right ? r : l;
  }
  sts_f32(val + pos * SB, v);
  sts_s32(idx + pos * SB, j);
  st.thr = root;
}

// FIFO -> heap
template <bool REGF>
static __device__ __noinline__ EpiState epi_drain(EpiState e, EpiAddr a) {
  ListState st;
  st.cnt = e.cnt;
  st.thr = e.lthr;
  if (REGF) {
    for (int s = 0; s < e.fcnt; ++s) {
      const float v = s == 0 ? e.f0v : e.f1v;
      const int j = s == 0 ? e.f0i : e.f1i;
      if (v > st.thr) nominee_push(a.val, a.idx, a.kc, st, v, j);
    }
  } else {
    for (int s = 0; s < e.fcnt; ++s) {
      const float v = lds_f32(a.fv + s * (F16_BM * 4));
      const int j = lds_s32(a.fi + s * (F16_BM * 4));
      if (v > st.thr) nominee_push(a.val, a.idx, a.kc, st, v, j);
    }
  }
  e.cnt = st.cnt;
  e.lthr = st.thr;
  e.fcnt = 0;
  e.thr = fmaxf(e.thr, st.thr);
  sts_f32(a.thr, st.thr);
  return e;
}

// queue every column of one group of 8 that beats the threshold, in index order
template <bool REGF>
static __device__ __noinline__ EpiState epi_scan8(EpiState e, EpiAddr a, int jbase, float v0, float v1, float v2,
                                                  float v3, float v4, float v5, float v6, float v7) {
  constexpr int FCAP = REGF ? 2 : F16_FCAP;
  const float v[8] = {v0, v1, v2, v3, v4, v5, v6, v7};
  int last = -1;
  while (true) {
    float cv = -INFINITY;
    int cc = -1, cnt = 0;
#pragma unroll
    for (int c = 7; c >= 0; --c) {
      const bool p = (c > last) && (v[c] > e.thr);
      cv = p ? v[c] : cv;
      cc = p ? c : cc;
      cnt += p ? 1 : 0;
    }
    if (cc < 0) break;
    if (e.fcnt == FCAP) e = epi_drain<REGF>(e, a);
    if (cv > e.thr) {                               // the drain may have raised the threshold
      if (REGF) {
        if (e.fcnt == 0) { e.f0v = cv; e.f0i = jbase + cc; } else { e.f1v = cv; e.f1i = jbase + cc; }
      } else {
        sts_f32(a.fv + e.fcnt * (F16_BM * 4), cv);
        sts_s32(a.fi + e.fcnt * (F16_BM * 4), jbase + cc);
      }
      ++e.fcnt;
    }
    if (cnt == 1) break;
    last = cc;
  }
  return e;
}

// PAIR: two CTAs of a cluster (two neighbouring query blocks) share every db tile -- each loads HALF of it, and
// the leader's tcgen05.mma.cta_group::2 (M = 256) reads both halves.  L2 -> shared-memory traffic and the
// shared-memory fill rate per SM halve; both bounded the single-CTA kernel (profiles/README.md, r01g).
// MODE: 0 = sweep, 1 = seed sweep over the db sample (segment maxima only), 2 = sweep with the BGNN_F16_DBG
// bottleneck experiments / wait-cycle instrumentation compiled in (tools/diag_knn_*.sh).
// EW: epilogue warps per TMEM lane quarter = lists per (row, db split).  The epilogue is bound by the LATENCY of its
// rare paths (one lane of 32 walks the hit path while the tile's other rows wait: 24 % of the warps' time; the heap
// drains: 14 %; profiles/r02w_knn_f16_ncu.txt).  EW = 4 halves the columns, the registers and the hit probability per
// warp and puts four epilogue warps on every scheduler, but the hits per warp only drop by a fifth and the tcgen05.ld's
// contend: measured equal to EW = 2 (42.2 vs 42.6 ms), kept as an experiment (BGNN_F16_EW=4).
template <int BN, bool PAIR, int MODE, int EW>
__global__ void __launch_bounds__(f16_threads(EW), 1)
knn_cosine_f16_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_db,
                      int nq, int ndb, int kblocks, int tiles_total, int tiles_per_split, int kc, int stages, int kps,
                      const float* __restrict__ thr_init, float* __restrict__ cand_val, int* __restrict__ cand_idx,
                      int seed_segs, int dbg_arg, int drain_tiles) {
  constexpr bool SEED = MODE == 1;
  constexpr bool REGF = EW == 4;                   // pending candidates in registers
  static_assert(EW == 2 || EW == 4, "two or four epilogue warps per lane quarter");
  static_assert(BN % (32 * EW) == 0, "every epilogue warp takes whole 32-column chunks");
  const int dbg = (MODE == 2) ? dbg_arg : 0;       // folds every experiment branch away outside MODE 2
  constexpr int B_ROWS = PAIR ? BN / 2 : BN;     // db rows of a tile this CTA loads
  constexpr int B_STAGE = B_ROWS * F16_BK * 2;
  constexpr int NBUF = 512 / BN;                 // accumulator buffers in TMEM: 2 x 256 or 4 x 128 columns
  constexpr int TMEM_COLS = 512;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* a_base = smem;                                          // kblocks * 16 KB, resident
  unsigned char* b_base = a_base + (size_t)kblocks * F16_A_KBLOCK;       // stages * B_STAGE ring
  constexpr int FIFO_ELEMS = REGF ? 0 : 2 * F16_FCAP * F16_BM;
  float* lval = reinterpret_cast<float*>(b_base + (size_t)stages * kps * B_STAGE);   // [EW][kc][128]
  int* lidx = reinterpret_cast<int*>(lval + (size_t)EW * kc * F16_BM);          // [EW][kc][128]
  float* ffv = reinterpret_cast<float*>(lidx + (size_t)EW * kc * F16_BM);    // [2][FCAP][128] pending values (EW = 2)
  int* ffi = reinterpret_cast<int*>(ffv + FIFO_ELEMS);                        // [2][FCAP][128] pending indices (EW = 2)
  float* thr_sh = reinterpret_cast<float*>(ffi + FIFO_ELEMS);                 // [EW][128] list thresholds, shared by the warps of a quarter
  uint64_t* bars = reinterpret_cast<uint64_t*>(thr_sh + EW * F16_BM);
  uint64_t* full_bar = bars;                      // [stages]
  uint64_t* empty_bar = bars + stages;            // [stages]
  uint64_t* tfull_bar = bars + 2 * stages;               // [NBUF]
  uint64_t* tempty_bar = bars + 2 * stages + NBUF;       // [NBUF]
  uint64_t* a_bar = bars + 2 * stages + 2 * NBUF;        // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 2 * NBUF + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;   // 0 = leader: issues the MMAs, owns full / tempty / a barriers
  const int q0 = blockIdx.x * F16_BM;
  const int split = blockIdx.y;
  const int tile_begin = split * tiles_per_split;
  const int tile_end = min(tiles_total, tile_begin + tiles_per_split);
  const int ntiles = tile_end - tile_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int b = 0; b < NBUF; ++b) { mbar_init(smem_u32(&tfull_bar[b]), 1); mbar_init(smem_u32(&tempty_bar[b]), (PAIR ? 8 : 4) * EW); }
    mbar_init(smem_u32(a_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {                                      // the same warp of both CTAs
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();  // barrier inits visible to the peer before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_db) : "memory");
      // PAIR: both CTAs load (their query block, their half of every db tile) but the bytes are counted on the
      // LEADER's barriers, which alone announces the totals
      const uint32_t ab = smem_u32(a_bar);
      if (rank == 0) mbar_expect_tx(ab, (uint32_t)((PAIR ? 2 : 1) * kblocks * F16_A_KBLOCK));
      const uint32_t ab_l = PAIR ? mapa_rank(ab, 0) : ab;
      for (int kb = 0; kb < kblocks; ++kb) {
        if (PAIR) tma_load_2d_pair(smem_u32(a_base + (size_t)kb * F16_A_KBLOCK), &map_q, ab_l, kb * F16_BK, q0);
        else tma_load_2d(smem_u32(a_base + (size_t)kb * F16_A_KBLOCK), &map_q, ab, kb * F16_BK, q0);
      }
      // Ring of `stages` slots; a slot holds `kps` consecutive k-blocks of one db tile and is announced by ONE
      // barrier (per-slot bookkeeping of this thread, not L2 bandwidth, bounded the operand stream: r01g notes).
      int s = 0, nloads = 0;
      uint32_t ph = 0;
      long long w_empty = 0, t_start = clock64();
      for (int t = 0; t < ntiles; ++t) {
        const int db0 = (tile_begin + t) * BN + (int)rank * B_ROWS;
        for (int kb0 = 0; kb0 < kblocks; kb0 += kps) {
          const long long c0 = (dbg & 8) ? clock64() : 0;
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
          if (dbg & 8) w_empty += clock64() - c0;
          const uint32_t fb = smem_u32(&full_bar[s]);
          if (rank == 0) mbar_expect_tx(fb, (uint32_t)((PAIR ? 2 : 1) * kps * B_STAGE));
          const uint32_t fb_l = PAIR ? mapa_rank(fb, 0) : fb;
          const uint32_t dst = smem_u32(b_base + (size_t)s * kps * B_STAGE);
          for (int kk = 0; kk < kps; ++kk) {
            if (PAIR) tma_load_2d_pair(dst + kk * B_STAGE, &map_db, fb_l, (kb0 + kk) * F16_BK, db0);
            else tma_load_2d(dst + kk * B_STAGE, &map_db, fb, (kb0 + kk) * F16_BK, db0);
          }
          ++nloads;
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
      }
      if ((dbg & 8) && blockIdx.x < 2 && blockIdx.y == 0)
        printf("cta %d producer: total %lld cyc, waiting for a free slot %lld cyc (%d slot loads)\n", (int)blockIdx.x,
               clock64() - t_start, w_empty, nloads);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(PAIR ? 2 * F16_BM : F16_BM, BN);
      mbar_wait(smem_u32(a_bar), 0u);
      tc_fence_after();
      int s = 0;
      uint32_t ph = 0;
      const uint64_t ad0 = make_kmajor_sw128_desc(smem_u32(a_base));
      const uint64_t bd0 = make_kmajor_sw128_desc(smem_u32(b_base));
      long long w_full = 0, w_tempty = 0, t_start = clock64();
      for (int t = 0; t < ntiles; ++t) {
        const int buf = t % NBUF;
        const uint32_t tph = (uint32_t)(t / NBUF) & 1u;
        long long c0 = (dbg & 8) ? clock64() : 0;
        mbar_wait(smem_u32(&tempty_bar[buf]), tph ^ 1u);
        if (dbg & 8) w_tempty += clock64() - c0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        for (int kb0 = 0; kb0 < kblocks; kb0 += kps) {
          c0 = (dbg & 8) ? clock64() : 0;
          mbar_wait(smem_u32(&full_bar[s]), ph);
          if (dbg & 8) w_full += clock64() - c0;
          tc_fence_after();
          // descriptors advance in 16-byte units: a k-block of A is 16 KB, of B B_STAGE bytes; 16 fp16 (one MMA
          // K step) are 32 B inside the 128-B swizzle row
          uint64_t ad = ad0 + (uint64_t)(kb0 * (F16_A_KBLOCK >> 4));
          uint64_t bd = bd0 + (uint64_t)(s * kps * (B_STAGE >> 4));
          for (int kk = 0; kk < kps; ++kk, ad += (F16_A_KBLOCK >> 4), bd += (B_STAGE >> 4)) {
#pragma unroll
            for (int k4 = 0; k4 < F16_BK / F16_UMMA_K; ++k4) {
              if (dbg & 4) continue;
              const uint32_t acc = (kb0 + kk + k4) != 0 ? 1u : 0u;
              if (PAIR) tc_mma_f16_pair(d_tmem, ad + (uint64_t)(k4 * 2), bd + (uint64_t)(k4 * 2), idesc, acc);
              else tc_mma_f16(d_tmem, ad + (uint64_t)(k4 * 2), bd + (uint64_t)(k4 * 2), idesc, acc);
            }
          }
          // frees the ring slot (in both CTAs of a pair) when these MMAs retire
          if (PAIR) tc_commit_pair(smem_u32(&empty_bar[s])); else tc_commit(smem_u32(&empty_bar[s]));
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
        // accumulator ready for the epilogue (of both CTAs)
        if (PAIR) tc_commit_pair(smem_u32(&tfull_bar[buf])); else tc_commit(smem_u32(&tfull_bar[buf]));
      }
      if ((dbg & 8) && blockIdx.x < 2 && blockIdx.y == 0)
        printf("cta %d mma issuer: total %lld cyc, waiting for operands %lld cyc, for a free accumulator %lld cyc (%d tiles)\n",
               (int)blockIdx.x, clock64() - t_start, w_full, w_tempty, ntiles);
    }
  } else {
    // ===================== epilogue: thread <-> query row, warp pair <-> lane quarter =====================
    const int quarter = warp & 3;                  // TMEM lanes this warp may read: 32*quarter ..
    const int half = (warp - 2) >> 2;              // which of the EW warps of the quarter
    const int r_in_tile = quarter * 32 + lane;
    const bool row_ok = q0 + r_in_tile < nq;
    float* my_val = lval + (size_t)half * kc * F16_BM + r_in_tile;
    int* my_idx = lidx + (size_t)half * kc * F16_BM + r_in_tile;
    const uint32_t my_val_s = smem_addr(my_val), my_idx_s = smem_addr(my_idx);
    // The EW warps of a lane quarter keep separate lists for the same rows but share their thresholds:
    // a column is kept only if it beats the largest of the row's list thresholds.  Sound for the certification in
    // knn_select.cu: every column any warp discards scores <= the largest of the final minima of the FULL lists
    // (a threshold is published only once its list is full, and thresholds only rise).
    const uint32_t row_thr_s = smem_addr(thr_sh + r_in_tile);        // + l * 512: list l of this row
    const uint32_t my_thr_s = row_thr_s + (uint32_t)(half * F16_BM * 4);
    sts_f32(my_thr_s, -INFINITY);
    asm volatile("bar.sync 1, %0;" ::"n"(128 * EW) : "memory");     // epilogue warps only
    // effective threshold = max(own list, partner list, seed): the seed is a score that at least kSeedKc db
    // rows of a sample reach (knn_seed_thr_kernel), so the lists skip most of their start-up insertions
    EpiState es;
    es.cnt = 0;
    es.fcnt = 0;
    es.lthr = -INFINITY;
    es.thr = (thr_init && row_ok) ? __ldg(thr_init + q0 + r_in_tile) : -INFINITY;
    es.f0v = es.f1v = 0.f;
    es.f0i = es.f1i = 0;
    // Heap insertions are deferred: a column that beats the threshold is appended to a small per-thread
    // FIFO (two stores), and the FIFOs are drained into the heaps by ALL lanes of the warp together every
    // F16_DRAIN_TILES tiles (or by one lane alone when its FIFO is full).  In steady state a chunk holds a
    // candidate for one or two of the 32 rows only, so inline insertion ran the long heap code at 1/32 lane
    // occupancy (profiles/r01g); drained together it is shared by every row that has something pending.
    // The threshold a lane filters with is then slightly stale, which only lets a few more columns through.
    EpiAddr ea;
    ea.val = my_val_s;
    ea.idx = my_idx_s;
    ea.fv = REGF ? 0u : smem_addr(ffv + (size_t)half * F16_FCAP * F16_BM + r_in_tile);
    ea.fi = REGF ? 0u : smem_addr(ffi + (size_t)half * F16_FCAP * F16_BM + r_in_tile);
    ea.thr = my_thr_s;
    ea.kc = kc;
    float seg_max = -INFINITY, seed_min = INFINITY;   // seed mode only (see below)
    // Seed mode (the sweep over the db sample): no lists at all.  The tiles of this CTA are cut into seed_segs
    // segments; the smallest of the segment maxima is a score that at least seed_segs sampled rows reach, which
    // is all the seed has to guarantee -- maxima only instead of a top-k.
    const int seg_tiles = SEED ? max(1, ntiles / max(seed_segs, 1)) : 0;
    int seg_left = seg_tiles, segs_done = 0;
    int drain_left = drain_tiles;
    long long w_tfull = 0, w_ld = 0, w_proc = 0, w_drain = 0, w_hit = 0, e_start = clock64();
    int n_hit_tiles = 0;
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t % NBUF;
      const uint32_t tph = (uint32_t)(t / NBUF) & 1u;
      long long c0 = (dbg & 8) ? clock64() : 0;
      if (!SEED && --drain_left == 0) {
        drain_left = drain_tiles;
        if (__any_sync(0xffffffffu, es.fcnt > 0)) es = epi_drain<REGF>(es, ea);
      }
      if (dbg & 8) { const long long c1 = clock64(); w_drain += c1 - c0; c0 = c1; }
      mbar_wait(smem_u32(&tfull_bar[buf]), tph);
      if (dbg & 8) { const long long c1 = clock64(); w_tfull += c1 - c0; c0 = c1; }
      tc_fence_after();
      const int db0 = (tile_begin + t) * BN;
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN);
      if (dbg & 2) {                                // bottleneck experiments: hand the buffer straight back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(mapa_rank(smem_u32(&tempty_bar[buf]), 0)); else mbar_arrive(smem_u32(&tempty_bar[buf]));
        }
        continue;
      }
      // All of this warp's chunks of the tile (half, half+2, ...) are pulled into registers up front and the
      // TMEM buffer goes straight back to the MMA issuer: how long the selection below takes (it varies a lot
      // from warp to warp and tile to tile) no longer decides when the MMAs of tile t+2 may start.
      constexpr int NCH = BN / (32 * EW);           // chunks per warp per tile
      float rr[NCH][32];
#pragma unroll
      for (int i = 0; i < NCH; ++i) tc_ld32(taddr0 + (uint32_t)((half + EW * i) * 32), rr[i]);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                               // the MMA issuer lives in the leader CTA
        if (PAIR) mbar_arrive_cluster(mapa_rank(smem_u32(&tempty_bar[buf]), 0)); else mbar_arrive(smem_u32(&tempty_bar[buf]));
      }
      if (dbg & 8) { const long long c1 = clock64(); w_ld += c1 - c0; c0 = c1; }
      if (db0 + BN > ndb) {                          // only the last db tile has zero-filled columns
#pragma unroll
        for (int i = 0; i < NCH; ++i)
#pragma unroll
          for (int c = 0; c < 32; ++c) rr[i][c] = (db0 + (half + EW * i) * 32 + c < ndb) ? rr[i][c] : -INFINITY;
      }
      if (!(dbg & 1)) {
        // chunk maxima first (3-input maxima, 4 instructions per group of 8, independent of each other), then one
        // test per tile.  Only the chunk maxima stay live: a hit recomputes the group maxima of the chunks concerned
        // (kept per group they cost 8-16 registers, which the compiler spilled on every tile).
        float cm[NCH];
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          float gmx[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float m012 = max3f(rr[i][g * 8 + 0], rr[i][g * 8 + 1], rr[i][g * 8 + 2]);
            const float m345 = max3f(rr[i][g * 8 + 3], rr[i][g * 8 + 4], rr[i][g * 8 + 5]);
            gmx[g] = max3f(m012, m345, fmaxf(rr[i][g * 8 + 6], rr[i][g * 8 + 7]));
          }
          cm[i] = max3f(gmx[0], gmx[1], fmaxf(gmx[2], gmx[3]));
        }
        float mx = cm[0];
#pragma unroll
        for (int i = 1; i < NCH; ++i) mx = fmaxf(mx, cm[i]);
        if (SEED) {
          seg_max = fmaxf(seg_max, mx);
          if (--seg_left == 0) {                     // a segment is complete (tiles past the last full one are ignored)
            seg_left = seg_tiles;
            if (segs_done++ < seed_segs) seed_min = fminf(seed_min, seg_max);
            seg_max = -INFINITY;
          }
        } else {
#pragma unroll
          for (int l = 0; l < EW; ++l) es.thr = fmaxf(es.thr, lds_f32(row_thr_s + (uint32_t)(l * F16_BM * 4)));
          const long long ch0 = (dbg & 8) ? clock64() : 0;
          if ((dbg & 8) && __any_sync(0xffffffffu, row_ok && mx > es.thr)) ++n_hit_tiles;
          if (row_ok && mx > es.thr) {
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
              if (!(cm[i] > es.thr)) continue;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                // The copies are made opaque to the compiler INSIDE the branch: without that it if-converts the
                // select chains of all groups of a tile into every hit (290 instructions per hit at one active
                // lane, profiles/r01p_knn_f16_ncu.txt)
                float v[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) { v[c] = rr[i][g * 8 + c]; asm volatile("" : "+f"(v[c])); }
                const float gmax = max3f(max3f(v[0], v[1], v[2]), max3f(v[3], v[4], v[5]), fmaxf(v[6], v[7]));
                if (gmax > es.thr) {
                  // the usual case inline: exactly one column of the group beats the threshold -- it is the
                  // group maximum, only its position is missing -- and the FIFO has room
                  // count and position in one sum (3-input adds, depth 2) instead of two chains of 8 dependent selects
                  int tk[8];
#pragma unroll
                  for (int c = 0; c < 8; ++c) tk[c] = (v[c] > es.thr) ? (0x100 | c) : 0;
                  const int pk = (tk[0] + tk[1] + tk[2]) + (tk[3] + tk[4] + tk[5]) + (tk[6] + tk[7]);
                  const int first = pk & 0xff;       // the column's position when exactly one beats the threshold
                  const int jb = db0 + (half + EW * i) * 32 + g * 8;
                  if ((pk >> 8) == 1 && es.fcnt < (REGF ? 2 : F16_FCAP)) {
                    if (REGF) {
                      if (es.fcnt == 0) { es.f0v = gmax; es.f0i = jb + first; } else { es.f1v = gmax; es.f1i = jb + first; }
                    } else {
                      sts_f32(ea.fv + es.fcnt * (F16_BM * 4), gmax);
                      sts_s32(ea.fi + es.fcnt * (F16_BM * 4), jb + first);
                    }
                    ++es.fcnt;
                  } else {
                    es = epi_scan8<REGF>(es, ea, jb, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
                  }
                }
              }
            }
          }
          if (dbg & 8) { __syncwarp(); w_hit += clock64() - ch0; }
        }
      }
      if (dbg & 8) w_proc += clock64() - c0;
    }
    if ((dbg & 8) && blockIdx.x < 1 && blockIdx.y == 0 && lane == 0)
      printf("cta %d epilogue warp %d: total %lld cyc, waiting for a tile %lld, tmem loads %lld, selection %lld (of which hit path %lld in %d tiles), drains %lld (%d tiles)\n",
             (int)blockIdx.x, warp, clock64() - e_start, w_tfull, w_ld, w_proc, w_hit, n_hit_tiles, w_drain, ntiles);
    if (SEED) {
      // one value per (list, row): -inf when this CTA saw no complete segment
      if (row_ok) cand_val[((long long)split * EW + half) * nq + (q0 + r_in_tile)] = (seed_min < INFINITY) ? seed_min : -INFINITY;
    } else {
    es = epi_drain<REGF>(es, ea);
    if (row_ok) {
      const long long base = (((long long)split * EW + half) * nq + (q0 + r_in_tile)) * kc;
      for (int s = 0; s < kc; ++s) {
        const bool f = s < es.cnt;
        cand_val[base + s] = f ? my_val[s * F16_BM] : -INFINITY;
        cand_idx[base + s] = f ? my_idx[s * F16_BM] : -1;
      }
    }
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();  // nobody leaves while the peer may still signal or be read
  if (warp == 1) {
    tc_fence_after();
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                   : "memory");
    else
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

// List length for EW lists per (row, split).  The lists of a row cover disjoint column classes (32-column chunks dealt
// round-robin to the EW warps), so the row's best k + 1 columns spread over them like Binomial(k + 1, 1 / EW) unless the
// db order correlates with the chunk pattern; a list that would have to hold more than kc of them makes the row fail
// certification (knn_select.cu) and the row is redone exactly -- correct, but slow when there are many.  kc is the
// smallest length whose overflow probability per row stays below `target` (EW = 2 keeps k + 4 >= k + 1: no overflow).
static int f16_list_len(int k, int ew, double target) {
  const int n = k + 1;
  if (ew == 2) return (k + 4 + 3) / 4 * 4;
  const double p = 1.0 / ew;
  for (int kc = (n + ew - 1) / ew; kc < n; ++kc) {
    // P(X > kc), X ~ Binomial(n, p)
    double tail = 0.0, term = 1.0;
    for (int x = 0; x < n; ++x) term *= (1.0 - p);               // P(X = 0)
    double px = term;
    for (int x = 0; x <= n; ++x) {
      if (x > kc) tail += px;
      px = px * (double)(n - x) / (double)(x + 1) * p / (1.0 - p);
    }
    if (ew * tail <= target) return kc;
  }
  return n;
}

// Work decomposition: kc nominees per (row, list) next to the resident query block and the B ring.
TcPlan tc_plan_f16(int nq, int ndb, int d, int k, int kc_fixed) {
  TcPlan p;
  p.bn = 0;
  p.ew = 2;
  const int ldh = (d + F16_BK - 1) / F16_BK * F16_BK;
  const int a_bytes = (ldh / F16_BK) * F16_A_KBLOCK;
  static const int force_bn = getenv("BGNN_F16_BN") ? atoi(getenv("BGNN_F16_BN")) : 0;   // tuning experiments
  // CTA pairs by default: every db tile crosses L2 -> shared memory once per 256 query rows instead of once per 128
  // (the single-CTA sweep without its epilogue stops at 69 % of the tensor pipe = the chip's ~6300 B/clk L2 throughput;
  // pairs reach 75 %), BGNN_F16_PAIR=0 switches them off.  BGNN_F16_EW=4: four epilogue warps / lists per lane quarter
  // (single CTAs only) -- measured equal to two (42.2 vs 42.6 ms) and the shorter lists fail certification more often.
  static const int pair_env = getenv("BGNN_F16_PAIR") ? atoi(getenv("BGNN_F16_PAIR")) : 1;
  static const int force_ew = getenv("BGNN_F16_EW") ? atoi(getenv("BGNN_F16_EW")) : 0;       // 2 | 4
  const int qblocks = (nq + F16_BM - 1) / F16_BM;
  p.pair = (pair_env && force_ew != 4 && qblocks >= 2) ? 1 : 0;
  const int ew_first = (force_ew == 4 && kc_fixed == 0) ? 4 : 2;
  for (int ew = ew_first; ew >= 2 && p.bn == 0; ew -= 2) {
    // expected uncertified rows of the whole call <= ~1.5 (a handful go through knn_exact_rows at ~0.2 ms each)
    const double target = fmin(6e-6, 1.5 / (double)(nq > 0 ? nq : 1));
    const int kc = kc_fixed > 0 ? kc_fixed : f16_list_len(k, ew, target);
    const int list_bytes = ew * kc * F16_BM * 8 + f16_fifo_bytes(ew);
    for (int bn = (force_bn == 128 ? 128 : 256); bn >= 128; bn >>= 1) {
      if (bn % (32 * ew) != 0) continue;
      const int stage_bytes = (p.pair ? bn / 2 : bn) * F16_BK * 2;
      const int units = (F16_SMEM_MAX - f16_smem_fixed(ew) - a_bytes - list_bytes) / stage_bytes;   // k-blocks that fit
      if (units < 3) continue;
      // one barrier per slot of kps k-blocks: as many k-blocks of a tile per slot as still leave >= 2 slots
      const int kblocks = ldh / F16_BK;
      int kps = kblocks;
      while (kps > 1 && (kblocks % kps != 0 || units / kps < 2)) --kps;
      p.bn = bn;
      p.kps = kps;
      p.stages = min(units / kps, 8);
      p.kc = kc;
      p.ew = ew;
      break;
    }
  }
  if (p.bn == 0) return p;                                  // caller falls back to another sweep
  const int kc = p.kc;
  const int tiles = (ndb + p.bn - 1) / p.bn;
  int ns = (2 * kNumSMs + qblocks - 1) / qblocks;            // fill the machine when nq is small
  ns = max(1, min(min(ns, tiles), min(8, BGNN_MERGE_MAX_CAND / (p.ew * kc))));
  if (ns < 1) { p.bn = 0; return p; }
  p.tiles_per_split = (tiles + ns - 1) / ns;
  p.nsplit = (tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.nlists = p.ew * p.nsplit;
  if (p.nlists * kc > BGNN_MERGE_MAX_CAND) p.bn = 0;
  return p;
}

static int f16_dbg_env() {
  static const int dbg = getenv("BGNN_F16_DBG") ? atoi(getenv("BGNN_F16_DBG")) : 0;   // bottleneck experiments only
  return dbg;
}

template <int BN, bool PAIR, int MODE, int EW>
static int launch_f16_cfg(const void* qh, int nq, const void* dh, int ndb, int ldh, const TcPlan& plan,
                          const float* thr_init, float* cand_val, int* cand_idx, int seed_segs, cudaStream_t stream) {
  CUtensorMap mq, md;
  int rc;
  constexpr int B_ROWS = PAIR ? BN / 2 : BN;
  if ((rc = make_map_f16(&mq, qh, nq, ldh, F16_BM)) != BGNN_OK) return rc;
  if ((rc = make_map_f16(&md, dh, ndb, ldh, B_ROWS)) != BGNN_OK) return rc;
  const int kblocks = ldh / F16_BK;
  const size_t smem = f16_smem_fixed(EW) + (size_t)kblocks * F16_A_KBLOCK + (size_t)plan.stages * plan.kps * B_ROWS * F16_BK * 2 +
                      (size_t)EW * plan.kc * F16_BM * 8 + f16_fifo_bytes(EW);
  if (smem > (size_t)F16_SMEM_MAX || plan.stages < 2 || plan.kps < 1 || kblocks % plan.kps != 0) return BGNN_ERR_UNSUPPORTED;
  auto kern = knn_cosine_f16_kernel<BN, PAIR, MODE, EW>;
  BGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = (ndb + BN - 1) / BN;
  const int qblocks = (nq + F16_BM - 1) / F16_BM;
  const int dbg = f16_dbg_env();
  static const int drain = getenv("BGNN_F16_DRAIN") ? max(1, atoi(getenv("BGNN_F16_DRAIN"))) : F16_DRAIN_TILES;   // tuning experiments
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(PAIR ? (qblocks + 1) / 2 * 2 : qblocks, plan.nsplit);   // a pair = two neighbouring query blocks
  cfg.blockDim = dim3(f16_threads(EW));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  BGNN_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, mq, md, nq, ndb, kblocks, tiles, plan.tiles_per_split, plan.kc, plan.stages,
                                   plan.kps, thr_init, cand_val, cand_idx, seed_segs, dbg, drain));
  return BGNN_OK;
}

static int launch_f16_any(const void* qh, int nq, const void* dh, int ndb, int ldh, const TcPlan& plan,
                          const float* thr_init, float* cand_val, int* cand_idx, int seed_segs, cudaStream_t stream) {
  if (nq <= 0) return BGNN_OK;
  if (ldh % F16_BK != 0) return BGNN_ERR_INVALID_ARG;
  const int mode = seed_segs > 0 ? 1 : (f16_dbg_env() ? 2 : 0);
#define F16_GO(BN_, PAIR_, MODE_, EW_) \
  launch_f16_cfg<BN_, PAIR_, MODE_, EW_>(qh, nq, dh, ndb, ldh, plan, thr_init, cand_val, cand_idx, seed_segs, stream)
#define F16_GO_MODE(BN_, PAIR_, EW_) \
  (mode == 1 ? F16_GO(BN_, PAIR_, 1, EW_) : mode == 2 ? F16_GO(BN_, PAIR_, 2, EW_) : F16_GO(BN_, PAIR_, 0, EW_))
  if (plan.ew == 4) {                                  // single CTAs only (tc_plan_f16)
    if (plan.pair) return BGNN_ERR_UNSUPPORTED;
    if (plan.bn == 256) return F16_GO_MODE(256, false, 4);
    if (plan.bn == 128) return F16_GO_MODE(128, false, 4);
  } else if (plan.ew == 2) {
    if (plan.bn == 256) return plan.pair ? F16_GO_MODE(256, true, 2) : F16_GO_MODE(256, false, 2);
    if (plan.bn == 128) return plan.pair ? F16_GO_MODE(128, true, 2) : F16_GO_MODE(128, false, 2);
  }
#undef F16_GO_MODE
#undef F16_GO
  return BGNN_ERR_UNSUPPORTED;
}

int launch_knn_cosine_f16(const void* qh, int nq, const void* dh, int ndb, int ldh, const TcPlan& plan,
                          const float* thr_init, float* cand_val, int* cand_idx, cudaStream_t stream) {
  return launch_f16_any(qh, nq, dh, ndb, ldh, plan, thr_init, cand_val, cand_idx, 0, stream);
}

// ------------------------------------------------------------------------------------------ threshold seeding
// A streaming top-KC list performs ~KC ln(n/KC) insertions, most of them while its threshold is still loose.
// Seeding: sweep a strided SAMPLE of the db first (same kernel in seed mode: segment maxima, no lists), take
// per query row a score that at least kSeedKc sampled rows reach, and start the full sweep from that threshold.  Soundness is unchanged: the
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

// seed[row] = max over the row's lists of the list's segment-maxima minimum (each at least kSeedKc sampled
// rows reach); -inf when no list saw a complete segment.
__global__ void __launch_bounds__(256)
knn_seed_thr_kernel(const float* __restrict__ list_seed, int nlists, int nq, float* __restrict__ seed) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nq) return;
  float best = -INFINITY;
  for (int l = 0; l < nlists; ++l) best = fmaxf(best, list_seed[(long long)l * nq + row]);
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
  int rc = launch_f16_any(qh, nq, sample, srows, ldh, seed_plan, nullptr, cand_val, cand_idx, kSeedKc, stream);
  if (rc != BGNN_OK) return rc;
  knn_seed_thr_kernel<<<(nq + 255) / 256, 256, 0, stream>>>(cand_val, seed_plan.nlists, nq, seed);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

}  // namespace bgnn
