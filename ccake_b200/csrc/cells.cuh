// Phase 01: neighbour search as a cell-list build.
// Replaces Cabana::VerletList / LinkedCellList (reference src/system_state.cpp:651-719):
// instead of a CSR neighbour list (N * N_ngb int32, ~2 GB for config 1) the particles
// themselves are sorted by cell key, so that every later gather walks contiguous memory.
//
//   cell key  : k = (ix * ny + iy) * nz + iz,  i_d = floor((x_d - min_d) * rdx_d) clamped,
//               same grid as the reference (cell edge >= 2h, grid = 2x (Milne) or 6x
//               (Cartesian) the IC box);
//   sort      : stable LSD radix sort of (key, index), 8 bits per pass, hand-written
//               (block histogram -> exclusive scan -> stable block scatter);
//   cell_start: histogram of keys + exclusive scan;
//   permute   : every persistent column is gathered through the permutation.
// All kernels stream: they are graded against HBM bandwidth.
#pragma once

#include "state.cuh"

namespace ccb {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;                      // keys per thread per tile
constexpr int kSortTile = kSortThreads * kSortItems;
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

// ---- cell keys + per-cell histogram ---------------------------------------------
// keys outside [key_lo, key_hi) become `sentinel` (= ncells, sorts last) when drop_outside is set: particles that
// have just been handed to a neighbouring rank.
__global__ void cell_key_kernel(DeviceView dv, CellGrid g, uint32_t* __restrict__ key, uint32_t* __restrict__ idx,
                                uint32_t* __restrict__ cell_count, int64_t n, int drop_outside, uint32_t key_lo, uint32_t key_hi,
                                uint32_t sentinel) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double x = dv.f[CCB_F_POSITION][(int64_t)g.axis[d] * dv.cap + i];
    int ci = (int)floor((x - g.min[d]) * g.rdx[d]);
    // More than one cell outside the neighbour grid (NaN included): the clamped cell no longer bounds the particle's
    // neighbours. The reference indexes out of bounds there (src/system_state.cpp:636-638); here it is an error.
    if (!(ci >= -1 && ci <= g.nc[d])) {
      if (atomicCAS(&dv.err[0], 0, (int)CCB_ERR_OUT_OF_GRID) == 0) dv.err[1] = dv.id[i];
    }
    ci = ci < 0 ? 0 : (ci > g.nc[d] - 1 ? g.nc[d] - 1 : ci);
    c[d] = ci;
  }
  uint32_t k = (uint32_t)(((int64_t)c[0] * g.nc[1] + c[1]) * g.nc[2] + c[2]);
  if (drop_outside && (k < key_lo || k >= key_hi)) k = sentinel;
  key[i] = k;
  idx[i] = (uint32_t)i;
  atomicAdd(&cell_count[k], 1u);
}

// ---- slab decomposition helpers -----------------------------------------------------------------
// migration message = the migrating particles' permuted columns, column-major [column][count], + 4 int columns
struct MigrateList {
  int ncols;
  double* col[96];
  int32_t* icol[4];
};
__global__ void unpack_rows(MigrateList ml, int64_t start, int64_t count, const double* __restrict__ dbuf, const int32_t* __restrict__ ibuf) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  for (int c = 0; c < ml.ncols; ++c) ml.col[c][start + k] = dbuf[(int64_t)c * count + k];
  for (int c = 0; c < 4; ++c) ml.icol[c][start + k] = ibuf[(int64_t)c * count + k];
}
// Particles whose cell column has left the slab [col_lo, col_hi), found BEFORE the stage's sort so that one sort places
// leavers (dropped) and arrivals (appended) together: list[0] = device indices of those bound for rank - 1, list[1] for
// rank + 1; info[0..1] count them. A side without a neighbour keeps its particles (their keys stay on the grid).
__global__ void find_leavers(const double* __restrict__ x, double xmin, double rdx, int ncol, int col_lo, int col_hi, int has_left,
                             int has_right, int64_t n, uint32_t* __restrict__ list_left, uint32_t* __restrict__ list_right,
                             int32_t* __restrict__ info) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = (int)floor((x[i] - xmin) * rdx);   // cell_key_kernel's formula
  c = c < 0 ? 0 : (c > ncol - 1 ? ncol - 1 : c);
  if (has_left && c < col_lo) list_left[atomicAdd(&info[0], 1)] = (uint32_t)i;
  else if (has_right && c >= col_hi) list_right[atomicAdd(&info[1], 1)] = (uint32_t)i;
}
// The atomics above fill the lists in arbitrary order; ranking the (distinct) indices makes the order of the migration
// message, hence the order of the arrivals behind the neighbour's particles, hence every later sum, reproducible.
__global__ void rank_sort_u32(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int count) {
  __shared__ uint32_t tile[256];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t mine = t < count ? in[t] : 0xffffffffu;
  int rank = 0;
  for (int base = 0; base < count; base += 256) {
    __syncthreads();
    tile[threadIdx.x] = base + threadIdx.x < count ? in[base + threadIdx.x] : 0xffffffffu;
    __syncthreads();
    const int m = min(256, count - base);
    for (int k = 0; k < m; ++k) rank += tile[k] < mine;
  }
  if (t < count) out[rank] = mine;
}
__global__ void pack_rows_indexed(MigrateList ml, const uint32_t* __restrict__ list, int64_t count, double* __restrict__ dbuf,
                                  int32_t* __restrict__ ibuf) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const int64_t i = list[k];
  for (int c = 0; c < ml.ncols; ++c) dbuf[(int64_t)c * count + k] = ml.col[c][i];
  for (int c = 0; c < 4; ++c) ibuf[(int64_t)c * count + k] = ml.icol[c][i];
}
// any non-zero entry in [lo, hi) of the `count` received gradient records of `rec_doubles` doubles each -> *flag = 1
__global__ void ghost_q_flag(const double* __restrict__ rec, int64_t count, int rec_doubles, int lo, int hi, int* __restrict__ flag) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  bool nz = false;
  for (int c = lo; c < hi; ++c) nz |= (rec[k * rec_doubles + c] != 0.0);
  if (nz) *flag = 1;
}
// ---- peer-memory halo: ready / done flags between neighbouring ranks ------------------------------------------------
// One thread publishes `value` (epoch * 2 + this rank's sparsity flag) into a flag word of each neighbour (peer stores over
// NVLink; nullptr = no neighbour), after making this rank's records visible system-wide.
__global__ void p2p_signal(int* peer_a, int* peer_b, int epoch, const int* sparsity_or_null) {
  const int v = epoch * 2 + ((sparsity_or_null && *sparsity_or_null) ? 1 : 0);
  __threadfence_system();
  if (peer_a) *reinterpret_cast<volatile int*>(peer_a) = v;
  if (peer_b) *reinterpret_cast<volatile int*>(peer_b) = v;
  __threadfence_system();
}
// One thread waits until both own flag words (nullptr = no neighbour) carry at least `epoch`; the neighbours' sparsity bits are
// OR-ed into *ghost_flag (nullptr = not wanted). Gives up with CCB_ERR_COMM after ~60 s of polling rather than hanging the GPU (a
// neighbour that is merely slow -- first-touch IPC mapping, a profiler, heavy imbalance -- stays well inside that); the boundary
// launch of the gather that follows on the same stream sees the error word and does no work (gather_kernel).
__global__ void p2p_wait(const int* flag_a, const int* flag_b, int epoch, int* ghost_flag, int32_t* err) {
  int bits = 0;
  const int* f[2] = {flag_a, flag_b};
  for (int k = 0; k < 2; ++k) {
    if (!f[k]) continue;
    long long spins = 0;
    int v;
    while (((v = *reinterpret_cast<const volatile int*>(f[k])) >> 1) < epoch) {
      __nanosleep(200);
      if (++spins > 300000000LL) { atomicCAS(&err[0], 0, (int)CCB_ERR_COMM); return; }
    }
    if ((v >> 1) == epoch) bits |= v & 1; else bits |= 1;  // a later epoch already: assume the general case
  }
  if (ghost_flag && bits) *ghost_flag = 1;
  __threadfence_system();
}
// the eight record offsets exchange_column_tables needs on the host, gathered for ONE device-to-host copy
__global__ void table_ends(const uint32_t* __restrict__ own_first, const uint32_t* __restrict__ own_last, const uint32_t* __restrict__ tab_left,
                           const uint32_t* __restrict__ tab_right, int64_t P, int has_left, int has_right, const uint32_t* __restrict__ n_keyed,
                           uint32_t* __restrict__ out) {
  if (threadIdx.x != 0) return;
  out[8] = n_keyed[0];
  out[0] = own_first[0]; out[1] = own_first[P]; out[2] = own_last[0]; out[3] = own_last[P];
  out[4] = has_left ? tab_left[0] : 0u; out[5] = has_left ? tab_left[P] : 0u;
  out[6] = has_right ? tab_right[0] : 0u; out[7] = has_right ? tab_right[P] : 0u;
}
// received column table (record offsets in the sender's numbering) -> offsets into this rank's ghost region
__global__ void rebase_ghost_table(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int64_t nent, uint32_t base) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nent) out[k] = base + (in[k] - in[0]);
}

// ---- device-wide exclusive scan (uint32), three small kernels --------------------
__global__ void scan_block_sums(const uint32_t* __restrict__ in, uint32_t* __restrict__ sums, int64_t n) {
  __shared__ uint32_t sh[kScanThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  uint32_t acc = 0;
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = base + (int64_t)k * kScanThreads + threadIdx.x;
    if (i < n) acc += in[i];
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) t += sh[w];
    sums[blockIdx.x] = t;
  }
}
// single block: exclusive scan of the block sums in place
__global__ void scan_sums_inplace(uint32_t* __restrict__ sums, int64_t nb) {
  __shared__ uint32_t sh[1024];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < nb; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const uint32_t v = (i < nb) ? sums[i] : 0u;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const uint32_t t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0u;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nb) sums[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
}
// per tile: exclusive scan with the tile's offset; element order inside a tile is
// (k, thread) -> base + k * threads + thread, i.e. plain index order.
__global__ void scan_apply(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, const uint32_t* __restrict__ sums,
                           int64_t n, int write_total) {
  __shared__ uint32_t wsum[kScanThreads / 32];
  __shared__ uint32_t running;
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  if (threadIdx.x == 0) running = sums[blockIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = base + (int64_t)k * kScanThreads + threadIdx.x;
    const uint32_t v = (i < n) ? in[i] : 0u;
    uint32_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < warp; ++w) woff += wsum[w];
    const uint32_t r = running;
    if (i < n) out[i] = r + woff + x - v;
    __syncthreads();
    if (threadIdx.x == kScanThreads - 1) running = r + woff + x;
    __syncthreads();
  }
  // out[n] = grand total (cell_start needs ncells + 1 entries)
  if (write_total && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = running;
}

// single block: exclusive scan of up to 1024 * kScanSmallPer elements (in place is fine); one launch instead of three.
// Warp w owns the contiguous chunk [w * per * 32, (w + 1) * per * 32) and walks it in rows of 32 (coalesced loads and
// stores), scanning each row with shuffles and carrying the running total from row to row.
constexpr int kScanSmallPer = 32;
__global__ void __launch_bounds__(1024) scan_small(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n, int write_total) {
  __shared__ uint32_t wsum[32];
  const int per = (int)((n + 1023) / 1024);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = (int64_t)warp * per * 32;
  uint32_t v[kScanSmallPer];
#pragma unroll
  for (int k = 0; k < kScanSmallPer; ++k) {
    const int64_t e = base + (int64_t)k * 32 + lane;
    v[k] = (k < per && e < n) ? in[e] : 0u;
  }
  uint32_t carry = 0;
#pragma unroll
  for (int k = 0; k < kScanSmallPer; ++k) {
    if (k < per) {
      uint32_t x = v[k];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      const uint32_t row_total = __shfl_sync(0xffffffffu, x, 31);
      v[k] = x - v[k] + carry;  // exclusive within the warp's chunk
      carry += row_total;
    }
  }
  if (lane == 0) wsum[warp] = carry;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = wsum[lane];
    uint32_t z = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, z, o);
      if (lane >= o) z += y;
    }
    wsum[lane] = z - w;  // exclusive warp offsets
    if (write_total && lane == 31) out[n] = z;
  }
  __syncthreads();
  const uint32_t off = wsum[warp];
#pragma unroll
  for (int k = 0; k < kScanSmallPer; ++k) {
    const int64_t e = base + (int64_t)k * 32 + lane;
    if (k < per && e < n) out[e] = v[k] + off;
  }
}

// ---- LSD radix sort, one 8-bit digit per pass -----------------------------------
__global__ void radix_hist(const uint32_t* __restrict__ key, uint32_t* __restrict__ hist, int64_t n, int shift, int nblocks) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
  for (int k = 0; k < kSortItems; ++k) {
    const int64_t i = base + (int64_t)k * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(key[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];  // digit-major
}

// Stable scatter. Within a tile elements are visited in index order (round k, thread t);
// the rank of an element among equal digits is (#equal in earlier rounds) + (#equal in
// lower warps of this round) + (#equal in lower lanes of this warp).
__global__ void radix_scatter(const uint32_t* __restrict__ key_in, const uint32_t* __restrict__ idx_in,
                              uint32_t* __restrict__ key_out, uint32_t* __restrict__ idx_out,
                              const uint32_t* __restrict__ offs, int64_t n, int shift, int nblocks) {
  __shared__ uint32_t base[256];
  __shared__ uint32_t wcount[kSortThreads / 32][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  base[threadIdx.x] = offs[(int64_t)threadIdx.x * nblocks + blockIdx.x];
  for (int w = 0; w < kSortThreads / 32; ++w) wcount[w][threadIdx.x] = 0;
  __syncthreads();
  const int64_t tile = (int64_t)blockIdx.x * kSortTile;
  for (int k = 0; k < kSortItems; ++k) {
    const int64_t i = tile + (int64_t)k * kSortThreads + threadIdx.x;
    const bool valid = i < n;
    uint32_t kk = 0, vv = 0, d = 256u;  // 256 = "no digit" for the tail
    if (valid) { kk = key_in[i]; vv = idx_in[i]; d = (kk >> shift) & 255u; }
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank_in_warp == 0) wcount[warp][d] = __popc(peers);
    __syncthreads();
    if (valid) {
      uint32_t o = base[d] + rank_in_warp;
      for (int w = 0; w < warp; ++w) o += wcount[w][d];
      key_out[o] = kk;
      idx_out[o] = vv;
    }
    __syncthreads();
    {
      uint32_t t = 0;
      for (int w = 0; w < kSortThreads / 32; ++w) { t += wcount[w][threadIdx.x]; wcount[w][threadIdx.x] = 0; }
      base[threadIdx.x] += t;
    }
    __syncthreads();
  }
}

// ---- permutation of the persistent columns --------------------------------------
struct PermuteList {
  int ncols;
  const double* src[96];
  double* dst[96];
  double* dst2[96];   // second destination of a column (the RK snapshot y0 taken with the first sort of a step), or nullptr
  int nints;
  const int32_t* isrc[4];
  int32_t* idst[4];
};
__global__ void permute_columns(PermuteList pl, const uint32_t* __restrict__ perm, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t j = perm[i];
  for (int c = 0; c < pl.ncols; ++c) {
    const double v = pl.src[c][j];
    pl.dst[c][i] = v;
    if (pl.dst2[c]) pl.dst2[c][i] = v;
  }
  for (int c = 0; c < pl.nints; ++c) pl.idst[c][i] = pl.isrc[c][j];
}

// ---- task list of the gathers ------------------------------------------------------------------------------------------
// A task = up to `chunk` consecutive particles of ONE cell (a cell with more particles is cut into chunks), so that every
// target of a gather CTA has the same 3^D-cell stencil. task_flag marks the first particle of every task, an exclusive scan
// numbers them, task_scatter writes the list in cell order and the three counts the gather kernels read:
// info[0] = tasks, info[1] = tasks before particle b_lo (= the slab's first cell column), info[2] = tasks before b_hi.
__global__ void task_flag(const uint32_t* __restrict__ key, const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ flag, int64_t n,
                          uint32_t chunk) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t rank = (uint32_t)i - cell_start[key[i]];
  flag[i] = (rank % chunk) == 0u ? 1u : 0u;
}
// pos = exclusive scan of the flags (n + 1 entries). lo / hi point at the cell-table entries that bound the interior of a slab
// (nullptr = no neighbour on that side = no boundary column): they are read on the device, the host never sees them.
__global__ void task_scatter(const uint32_t* __restrict__ pos, int32_t* __restrict__ tasks, int64_t n, const uint32_t* __restrict__ lo,
                             const uint32_t* __restrict__ hi, int32_t* __restrict__ info) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    int64_t b_lo = lo ? (int64_t)lo[0] : 0, b_hi = hi ? (int64_t)hi[0] : n;
    if (b_hi < b_lo) b_hi = b_lo;  // a slab of one column with a neighbour on either side: every task is a boundary task
    info[0] = (int32_t)pos[n];
    info[1] = (int32_t)pos[b_lo < n ? b_lo : n];
    info[2] = (int32_t)pos[b_hi < n ? b_hi : n];
  }
  if (i < n && pos[i + 1] != pos[i]) tasks[pos[i]] = (int32_t)i;
}

// whole blocks of columns at once (ccb_upload_state / ccb_download_state): blockIdx.y = column
struct ColumnList {
  int ncols;
  double* col[32];
};
__global__ void gather_columns_by_id(const double* __restrict__ stage, ColumnList cl, const int32_t* __restrict__ id, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cl.col[blockIdx.y][i] = stage[(int64_t)blockIdx.y * n + id[i]];
}
__global__ void scatter_columns_by_id(ColumnList cl, double* __restrict__ stage, const int32_t* __restrict__ id, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stage[(int64_t)blockIdx.y * n + id[i]] = cl.col[blockIdx.y][i];
}

// rows of a row-major host array (staged on the device as src[n][w]) -> `ncols` columns dst + k * stride, in the current particle
// order: dst[k][i] = src[id[i]][c0 + k * cstep] (id == nullptr: identity). Replaces per-column transposes on the host.
__global__ void gather_rows_by_id(const double* __restrict__ src, int w, int c0, int cstep, int ncols, double* __restrict__ dst, int64_t stride,
                                  const int32_t* __restrict__ id, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t row = id ? id[i] : i;
  for (int k = 0; k < ncols; ++k) dst[(int64_t)k * stride + i] = src[row * w + c0 + k * cstep];
}

// the reverse: `w` columns in the current order -> row-major dst[n][w] in particle-id order
__global__ void scatter_rows_by_id(const double* __restrict__ src, int64_t stride, int w, double* __restrict__ dst, const int32_t* __restrict__ id, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t row = id ? id[i] : i;
  for (int k = 0; k < w; ++k) dst[row * w + k] = src[(int64_t)k * stride + i];
}

// 16 independent fp64 FMA chains per thread: the DFMA issue-rate ceiling (ccb_measure_dfma_peak)
__global__ void dfma_peak_kernel(double* out, int iters, double m) {
  double a[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = 1.0 + 1e-9 * (threadIdx.x + k);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = fma(a[k], m, 1e-12);
  }
  double s = 0.;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += a[k];
  if (s == 123.456) out[threadIdx.x] = s;
}

// integer columns (current order) <-> particle-id order
__global__ void scatter_int_by_id(const int32_t* __restrict__ src, int32_t* __restrict__ dst, const int32_t* __restrict__ id, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[id[i]] = src[i];
}
__global__ void gather_int_by_id(const int32_t* __restrict__ src, int32_t* dst, const int32_t* id, int64_t n) {  // dst may alias id
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[id[i]];
}

}  // namespace ccb
