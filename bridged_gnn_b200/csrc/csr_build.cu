// Edge list -> destination-major CSR on the device.
//
// Replaces, for the kernels below, what the reference does on every forward with
// torch_sparse.SparseTensor(row=edge_index[1], col=edge_index[0]) (models/backbones.py:464) and what
// torch_geometric.utils.coalesce does on the CPU (main_bridged_graph.py:75, 113, 193): sort by
// (dst, src), optionally drop duplicate edges, emit rowptr/col plus the permutation back to the
// caller's edge order.  The sort itself is CUB's radix sort (plumbing, run once per graph and cached
// by the host layer); the rest is hand-written.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "kernels.h"

namespace bgnn {

// key = (dst << nb) | src with nb = bits needed for a node id, so the radix sort touches 2*nb bits only.
// An edge with a node id outside [0, n) would spill into the other half of the key and later serve as a gather index:
// it is counted in *bad and parked under destination n, behind the last row, where no kernel reads it.
__global__ void make_keys_kernel(const long long* __restrict__ src, const long long* __restrict__ dst, long long e, long long n,
                                 int nb, unsigned long long* __restrict__ keys, unsigned int* __restrict__ vals,
                                 unsigned long long* __restrict__ bad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= e) return;
  long long s = src[i], d = dst[i];
  if (s < 0 || s >= n || d < 0 || d >= n) {
    atomicAdd(bad, 1ull);
    s = 0;
    d = n;
  }
  keys[i] = ((unsigned long long)d << nb) | (unsigned long long)s;
  vals[i] = (unsigned int)i;
}

__global__ void flag_heads_kernel(const unsigned long long* __restrict__ keys, long long e, int dedup,
                                  unsigned int* __restrict__ flags) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= e) return;
  flags[i] = (!dedup || i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

__global__ void compact_kernel(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ vals,
                               const unsigned int* __restrict__ flags, const unsigned int* __restrict__ pos,
                               long long e, int nb, int* __restrict__ col, long long* __restrict__ perm,
                               unsigned long long* __restrict__ ckeys, long long* __restrict__ e_out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= e) return;
  if (flags[i]) {
    unsigned int p = pos[i];
    col[p] = (int)(keys[i] & ((1ull << nb) - 1ull));
    if (perm) perm[p] = (long long)vals[i];
    ckeys[p] = keys[i];
  }
  if (i == e - 1) *e_out = (long long)pos[i] + (long long)flags[i];
}

// rowptr[r] = first position whose dst >= r  (r = 0..n)
__global__ void rowptr_kernel(const unsigned long long* __restrict__ ckeys, const long long* __restrict__ e_out,
                              long long n, int nb, int* __restrict__ rowptr) {
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n) return;
  long long lo = 0, hi = *e_out;
  while (lo < hi) {
    long long mid = (lo + hi) >> 1;
    if ((long long)(ckeys[mid] >> nb) < r) lo = mid + 1; else hi = mid;
  }
  rowptr[r] = (int)lo;
}

__global__ void zero_rowptr_kernel(long long n, int* rowptr, long long* e_out) {
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r <= n) rowptr[r] = 0;
  if (r == 0) { e_out[0] = 0; e_out[1] = 0; }
}

__global__ void zero_bad_kernel(long long* e_out) { e_out[1] = 0; }

static size_t cub_temp_bytes(long long e) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const unsigned int*)nullptr, (unsigned int*)nullptr, e);
  cub::DeviceScan::ExclusiveSum(nullptr, b, (const unsigned int*)nullptr, (unsigned int*)nullptr, e);
  return a > b ? a : b;
}

size_t csr_build_workspace_bytes(long long e) {
  if (e <= 0) return 256;
  size_t per = 8 + 8 + 4 + 4 + 4 + 4;  // keys x2, vals x2, flags, pos
  return (size_t)e * per + cub_temp_bytes(e) + 8 * 256;
}

int launch_edges_to_csr(const long long* src, const long long* dst, long long e, long long n, int dedup, int* rowptr,
                        int* col, long long* perm, long long* e_out, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (n < 0 || e < 0 || n >= (1ll << 31) || e >= (1ll << 31)) return BGNN_ERR_INVALID_ARG;
  const int T = 256;
  if (e == 0) {
    zero_rowptr_kernel<<<(unsigned)((n + 1 + T - 1) / T), T, 0, stream>>>(n, rowptr, e_out);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
  }
  Workspace w(ws, ws_bytes);
  auto* k0 = w.take<unsigned long long>(e);
  auto* k1 = w.take<unsigned long long>(e);
  auto* v0 = w.take<unsigned int>(e);
  auto* v1 = w.take<unsigned int>(e);
  auto* flags = w.take<unsigned int>(e);
  auto* pos = w.take<unsigned int>(e);
  size_t tb = cub_temp_bytes(e);
  auto* temp = w.take<char>(tb);
  if (!w.ok()) return BGNN_ERR_WORKSPACE;
  unsigned blocks = (unsigned)((e + T - 1) / T);
  int nb = 1;
  while ((1ll << nb) < n) ++nb;
  zero_bad_kernel<<<1, 1, 0, stream>>>(e_out);
  BGNN_LAUNCH_CHECK();
  make_keys_kernel<<<blocks, T, 0, stream>>>(src, dst, e, n, nb, k0, v0, reinterpret_cast<unsigned long long*>(e_out + 1));
  BGNN_LAUNCH_CHECK();
  // 2 nb + 1 key bits: destination n (the parking row of invalid edges) needs one bit more than a node id when n = 2^nb
  BGNN_CUDA_TRY(cub::DeviceRadixSort::SortPairs(temp, tb, k0, k1, v0, v1, e, 0, 2 * nb + 1, stream));
  flag_heads_kernel<<<blocks, T, 0, stream>>>(k1, e, dedup, flags);
  BGNN_LAUNCH_CHECK();
  BGNN_CUDA_TRY(cub::DeviceScan::ExclusiveSum(temp, tb, flags, pos, e, stream));
  compact_kernel<<<blocks, T, 0, stream>>>(k1, v1, flags, pos, e, nb, col, perm, k0, e_out);
  BGNN_LAUNCH_CHECK();
  rowptr_kernel<<<(unsigned)((n + 1 + T - 1) / T), T, 0, stream>>>(k0, e_out, n, nb, rowptr);
  BGNN_LAUNCH_CHECK();
  return BGNN_OK;
}

// ---- one-call graph preparation of the aggregation kernels ---------------------------------------------------
// Everything KT-GNN's forward / backward need from a NEW edge list, straight from the caller's edge_index:
//   * the self-loop rewrite of graph_partition (models/KTGNN.py:385-398: remove_self_loops, then add_self_loops)
//     happens while the sort keys are made -- an edge with src == dst is parked behind the last row like an invalid
//     one, and n keys (v, v) are appended -- so no filtered / concatenated copy of the edge list is ever written;
//   * destination-major CSR: ONE keys-only radix sort of (dst, src) (no permutation back to the input is needed: the
//     aggregation has no per-edge inputs);
//   * transposed CSR + the CSR -> CSC slot map: a STABLE sort of the CSR-ordered edges on the source bits alone
//     (half the radix passes of a second full sort; entries of a source row stay in destination order), whose
//     payload -- the CSR position -- is the slot map;
//   * both processing orders (rows by descending degree).
// 168 B of sort traffic per edge against 288 B for two full pair sorts, and none of the index / cat / nonzero
// passes of the host-side partition (2 of the 6.3 ms a new 2.1e7-edge graph cost, tools/profile_e2e.py).
__global__ void prep_keys_kernel(const long long* __restrict__ src, const long long* __restrict__ dst, long long e,
                                 long long n, int nb, int rewrite, unsigned long long* __restrict__ keys,
                                 unsigned long long* __restrict__ bad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long m = e + (rewrite ? n : 0);
  if (i >= m) return;
  unsigned long long k;
  if (i < e) {
    long long s = src[i], d = dst[i];
    if (s < 0 || s >= n || d < 0 || d >= n) {
      atomicAdd(bad, 1ull);
      s = 0;
      d = n;
    } else if (rewrite && s == d) {                  // a self loop of the input: dropped (re-added below)
      s = 0;
      d = n;
    }
    k = ((unsigned long long)d << nb) | (unsigned long long)s;
  } else {
    const unsigned long long v = (unsigned long long)(i - e);
    k = (v << nb) | v;
  }
  keys[i] = k;
}

// col of the CSR, keys (src, dst) of the transposed sort with the CSR position as payload; parked entries keep
// source n (behind the last transposed row)
__global__ void prep_split_kernel(const unsigned long long* __restrict__ keys, long long m, long long n, int nb,
                                  int* __restrict__ col, unsigned long long* __restrict__ tkeys,
                                  unsigned int* __restrict__ vals) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const unsigned long long k = keys[i], mask = (1ull << nb) - 1ull;
  const unsigned long long d = k >> nb, s = k & mask;
  col[i] = (int)s;
  if (tkeys) {
    tkeys[i] = d >= (unsigned long long)n ? ((unsigned long long)n << nb) : ((s << nb) | d);
    vals[i] = (unsigned int)i;
  }
}

// rowptr[r] = first position whose row field (key >> nb) is >= r, r = 0..n; count = rowptr[n] (parked entries follow)
__global__ void prep_rowptr_kernel(const unsigned long long* __restrict__ keys, long long m, long long n, int nb,
                                   int* __restrict__ rowptr, long long* __restrict__ count) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n) return;
  long long lo = 0, hi = m;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if ((long long)(keys[mid] >> nb) < r) lo = mid + 1; else hi = mid;
  }
  rowptr[r] = (int)lo;
  if (r == n && count) *count = lo;
}

__global__ void prep_transposed_kernel(const unsigned long long* __restrict__ tkeys, const unsigned int* __restrict__ vals,
                                       long long m, long long n, int nb, int* __restrict__ t_col,
                                       int* __restrict__ csr_to_csc) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= m) return;
  const unsigned long long k = tkeys[q];
  if ((k >> nb) >= (unsigned long long)n) return;    // parked
  t_col[q] = (int)(k & ((1ull << nb) - 1ull));
  csr_to_csc[vals[q]] = (int)q;
}

static size_t prep_sort_temp_bytes(long long m) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, a, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, m);
  cub::DeviceRadixSort::SortPairs(nullptr, b, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const unsigned int*)nullptr, (unsigned int*)nullptr, m);
  return a > b ? a : b;
}

size_t graph_prepare_workspace_bytes(long long e, long long n, int rewrite) {
  const long long m = e + (rewrite ? n : 0);
  if (m <= 0) return 256;
  const size_t deg = rows_by_degree_workspace_bytes(n);
  return (size_t)m * (8 + 8 + 4 + 4) + prep_sort_temp_bytes(m) + deg + 8 * 256;
}

int launch_graph_prepare(const long long* src, const long long* dst, long long e, long long n, int rewrite, int* rowptr,
                         int* col, int* t_rowptr, int* t_col, int* csr_to_csc, int* order, int* t_order, long long* e_out,
                         void* ws, size_t ws_bytes, cudaStream_t stream) {
  const long long m = e + (rewrite ? n : 0);
  if (n < 0 || e < 0 || n >= (1ll << 31) || m >= (1ll << 31)) return BGNN_ERR_INVALID_ARG;
  if ((t_rowptr != nullptr) != (t_col != nullptr) || (t_col != nullptr) != (csr_to_csc != nullptr)) return BGNN_ERR_INVALID_ARG;
  if (t_order && !t_rowptr) return BGNN_ERR_INVALID_ARG;
  const int T = 256;
  if (m == 0) {
    zero_rowptr_kernel<<<(unsigned)((n + 1 + T - 1) / T), T, 0, stream>>>(n, rowptr, e_out);
    BGNN_LAUNCH_CHECK();
    if (t_rowptr) BGNN_CUDA_TRY(cudaMemsetAsync(t_rowptr, 0, (size_t)(n + 1) * sizeof(int), stream));
    return BGNN_OK;
  }
  Workspace w(ws, ws_bytes);
  auto* k0 = w.take<unsigned long long>(m);
  auto* k1 = w.take<unsigned long long>(m);
  auto* v0 = w.take<unsigned int>(m);
  auto* v1 = w.take<unsigned int>(m);
  const size_t tb = prep_sort_temp_bytes(m);
  auto* temp = w.take<char>(tb);
  const size_t db = rows_by_degree_workspace_bytes(n);
  auto* dws = w.take<char>(db);
  if (!w.ok()) return BGNN_ERR_WORKSPACE;
  const unsigned blocks = (unsigned)((m + T - 1) / T), rblocks = (unsigned)((n + 1 + T - 1) / T);
  int nb = 1;
  while ((1ll << nb) < n) ++nb;
  zero_bad_kernel<<<1, 1, 0, stream>>>(e_out);
  BGNN_LAUNCH_CHECK();
  prep_keys_kernel<<<blocks, T, 0, stream>>>(src, dst, e, n, nb, rewrite, k0, reinterpret_cast<unsigned long long*>(e_out + 1));
  BGNN_LAUNCH_CHECK();
  size_t tbv = tb;
  BGNN_CUDA_TRY(cub::DeviceRadixSort::SortKeys(temp, tbv, k0, k1, m, 0, 2 * nb + 1, stream));
  prep_rowptr_kernel<<<rblocks, T, 0, stream>>>(k1, m, n, nb, rowptr, e_out);
  BGNN_LAUNCH_CHECK();
  prep_split_kernel<<<blocks, T, 0, stream>>>(k1, m, n, nb, col, t_col ? k0 : nullptr, v0);
  BGNN_LAUNCH_CHECK();
  int rc;
  if (order && (rc = launch_rows_by_degree(rowptr, n, 0, order, dws, db, stream)) != BGNN_OK) return rc;
  if (t_col) {
    tbv = tb;
    // stable: within a source row the CSR order (dst ascending) survives; source n (parked) needs bit 2 nb
    BGNN_CUDA_TRY(cub::DeviceRadixSort::SortPairs(temp, tbv, k0, k1, v0, v1, m, nb, 2 * nb + 1, stream));
    prep_rowptr_kernel<<<rblocks, T, 0, stream>>>(k1, m, n, nb, t_rowptr, nullptr);
    BGNN_LAUNCH_CHECK();
    prep_transposed_kernel<<<blocks, T, 0, stream>>>(k1, v1, m, n, nb, t_col, csr_to_csc);
    BGNN_LAUNCH_CHECK();
    if (t_order && (rc = launch_rows_by_degree(t_rowptr, n, 0, t_order, dws, db, stream)) != BGNN_OK) return rc;
  }
  return BGNN_OK;
}

// ---- rows by descending degree ---------------------------------------------------------------------------
// The row-parallel gather kernels take an optional processing order: longest rows first, so that a hub row of
// a kNN graph starts at once instead of forming the tail of the launch, and rows sharing a warp have similar
// lengths.  Stable radix sort of (max_deg - deg, row): equal degrees keep ascending row order.
// Rows shorter than min_degree all get the largest key: they follow the long rows in their natural order, which
// keeps the row-level loads of narrow feature rows coalesced (min_degree = 0: full sort).
__global__ void degree_keys_kernel(const int* __restrict__ rowptr, long long n, int min_degree,
                                   unsigned int* __restrict__ keys, unsigned int* __restrict__ vals) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int deg = rowptr[i + 1] - rowptr[i];
  keys[i] = deg >= min_degree ? 0x7ffffffeu - (unsigned int)deg : 0x7fffffffu;
  vals[i] = (unsigned int)i;
}

static size_t degree_sort_temp_bytes(long long n) {
  size_t a = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const unsigned int*)nullptr, (unsigned int*)nullptr,
                                  (const unsigned int*)nullptr, (unsigned int*)nullptr, n);
  return a;
}

size_t rows_by_degree_workspace_bytes(long long n) {
  if (n <= 0) return 256;
  return (size_t)n * 12 + degree_sort_temp_bytes(n) + 4 * 256;
}

int launch_rows_by_degree(const int* rowptr, long long n, int min_degree, int* order, void* ws, size_t ws_bytes,
                          cudaStream_t stream) {
  if (n < 0 || n >= (1ll << 31)) return BGNN_ERR_INVALID_ARG;
  if (n == 0) return BGNN_OK;
  Workspace w(ws, ws_bytes);
  auto* k0 = w.take<unsigned int>(n);
  auto* k1 = w.take<unsigned int>(n);
  auto* v0 = w.take<unsigned int>(n);
  size_t tb = degree_sort_temp_bytes(n);
  auto* temp = w.take<char>(tb);
  if (!w.ok()) return BGNN_ERR_WORKSPACE;
  degree_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(rowptr, n, min_degree, k0, v0);
  BGNN_LAUNCH_CHECK();
  BGNN_CUDA_TRY(cub::DeviceRadixSort::SortPairs(temp, tb, k0, k1, v0, reinterpret_cast<unsigned int*>(order), n, 0, 31, stream));
  return BGNN_OK;
}

}  // namespace bgnn
