// C ABI of the B200-native CCAKE hot path (include/ccake_b200.h): context, device
// memory, phase orchestration. One translation unit; the kernels live in
// cells.cuh (phase 01), gather.cuh (03, 09) and local.cuh (02, 04, 05, 08, 10 + RK).
//
// There is deliberately NO host fallback: every entry point that computes launches
// CUDA kernels on the context's stream and returns CCB_ERR_CUDA if no device is usable.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "cells.cuh"
#include "comm.h"
#include "eos_preinverted.cuh"
#include "eos_table_host.h"
#include "freeze_out.cuh"
#include "launchers.h"

using namespace ccb;

namespace {
enum { PH_NEIGHBOUR = 0, PH_GAMMA, PH_SMOOTH_FIELDS, PH_THERMO, PH_LOCAL, PH_SMOOTH_GRAD, PH_TIME_DERIV, PH_OTHER, PH_SLAB_MIGRATE, PH_SLAB_TABLES, PH_COUNT };
}

struct ccb_ctx {
  ccb_config cfg;
  int D = 2;
  bool milne = true;
  cudaStream_t stream = nullptr;
  int64_t n = 0, cap = 0;
  double t = 0.0;
  DeviceView dv;
  CellGrid grid;
  // allocations
  double* slab = nullptr;       // all public columns
  double* y0_slab = nullptr;    // snapshot of evolved columns
  double* alt_slab = nullptr;   // ping-pong target of the permutation (persistent + y0 columns)
  double* alt[CCB_F_COUNT];
  double* alt_y0[CCB_F_COUNT];
  // rk_order 4: the weighted sum of stage derivatives (the reference's `k` cache, src/evolver.cpp:38-60, 172-211): one column per
  // evolved column; it follows the particles through every sort like y0 does
  double *k_slab = nullptr, *alt_k_slab = nullptr;
  double* kacc[CCB_F_COUNT];
  double* alt_k[CCB_F_COUNT];
  int32_t* ints = nullptr;      // id, btrack, freeze, eos_id + alternates (8 columns)
  int32_t* alt_int[4];
  uint32_t *key_a = nullptr, *key_b = nullptr, *idx_a = nullptr, *idx_b = nullptr, *hist = nullptr, *cell_count = nullptr,
           *scan_sums = nullptr;
  int64_t ncells_alloc = 0, hist_alloc = 0, sums_alloc = 0;
  double* rec = nullptr;        // gather records
  int32_t* tasks = nullptr;     // gather task list (cell order), one entry per chunk of <= task_nt particles of one cell
  uint32_t* task_pos = nullptr; // cap + 1: task flags, then their exclusive scan
  int32_t* d_taskinfo = nullptr;// [0..2] task counts (cells.cuh: task_scatter), [4] / [5] task counters of the interior / boundary launch
  int32_t* d_small = nullptr;   // [1..2] error code and first offending particle id
  double* d_stage = nullptr;    // staging for id-ordered transfers (16 columns) and reductions
  double* init_tmp = nullptr;
  double* d_reduce = nullptr;   // per-block partials of ccb_reduce
  bool cells_valid = false;   // the cell table and sort order match the current positions (gathers may run)
  // the neighbour list of the stage as tile bitmasks (GatherParams::hitmask): stored by the field gather, read by the gradient gather
  unsigned long long* hitmask = nullptr;
  int hit_kmax = 0;
  int64_t hit_cap = 0;
  bool hit_valid = false, hit_off = false;
  // ccb_step: the RK snapshot y0 = y(t_n) is taken BY the first sort of the step (permute_columns writes the sorted evolved
  // columns twice) instead of copying 26 columns and then permuting both copies
  bool y0_pending = false;
  bool cells_built = false;   // a cell build has run on this particle set (dv.key matches the current order)
  bool have_state = false;
  std::string err;
  int64_t launches = 0;
  // phase timing (CCAKE_PROFILE analogue)
  bool profiling = false;
  struct EventPair { cudaEvent_t a, b; int phase; };
  std::vector<EventPair> ev_pool;
  size_t ev_used = 0;
  double phase_ms[PH_COUNT];
  int64_t phase_calls[PH_COUNT];
  KernelNorm kD, k2, k1;
  double efcheck = 0.0;
  // tabulated EoS: host copy (start-up only), device copies, the fall-back chain
  std::unique_ptr<ccb::HostTable> table;
  double *d_hot_s = nullptr, *d_hot_e = nullptr, *d_full = nullptr;
  ccb::EosChain chain;
  // pre-inverted EoS (online_inverter_enabled = 0): device copies of the tiles / the flat table, and the flat table's host copy
  std::vector<double*> pre_alloc;
  ccb::PreTile* d_pre_tiles = nullptr;
  ccb::PreTable pre_table{nullptr, 0};
  ccb::PreFlat pre_flat_dev{};
  std::vector<std::vector<double>> pre_flat_host;   // s_axis, nB_axis, T, muB, muS, muQ, e, p, cs2, dwds, dwdB (fm units)
  bool have_pre = false;
  // freeze-out (freeze_out.cuh): snapshot / results rows by particle id; FreezeOut::frzc, taup, taupp; SystemState::number_part_fo
  ccb::FreezeArrays frz{};
  double* frz_slab = nullptr;
  int frzc = 0;
  double taup = 0.0, taupp = 0.0;
  int64_t number_part_fo = 0;
  bool have_table = false;   // the EoS chain is ready: table loaded, or an analytic single-member chain (tanh_conformal, conformal_diagonal)
  bool use_chain = false;    // phase 04 runs through the chain kernels (every eos_type but the closed-form conformal one)
  // launch shapes of the gradient / field gather. nt (targets = threads per task) is common to both: it is the chunk size of
  // the task list
  ccb::GatherShape shape_grad{64, 64}, shape_fields{64, 64};
  int sms = 148;
  // slab decomposition (nranks > 1): own cell columns [slab_lo, slab_hi) along x, neighbours rank -+ 1
  std::unique_ptr<ccb::Link> comm;
  int slab_lo = 0, slab_hi = 0;
  int64_t rec_cap = 0;                       // records the gather buffer holds (own + ghosts)
  uint32_t* ghost_tab = nullptr;             // 2 x (P + 1) rebased column tables + 2 x (P + 1) receive scratch
  int32_t* d_info = nullptr;                 // small device scratch for counts
  uint32_t* mig_list = nullptr;              // migrants' device indices, ranked (cap mig_cap)
  uint32_t* d_ends = nullptr;                // 8 table ends of exchange_column_tables, fetched in one copy
  bool init_diffusion = false;               // the IC rows carried q^mu_a (ccb_ic::diffusion)
  int* d_qflag = nullptr;                    // DeviceView::qflag
  // peer-memory halo (CCB_HALO_P2P=1): the boundary tasks read the neighbours' record arrays over NVLink; flag words order the accesses
  bool p2p = false;
  int* d_flags = nullptr;                    // own flag words: [0] ready from left, [1] ready from right, [2] done from left, [3] done from right
  void* peer_rec[2] = {nullptr, nullptr};
  void* peer_flags[2] = {nullptr, nullptr};
  int halo_epoch = 0;                        // one epoch per gather
  cudaStream_t comm_stream = nullptr;        // halo exchanges run here, next to the interior tasks of the gather
  cudaEvent_t ev_pack = nullptr, ev_halo = nullptr;
  int64_t ghost_n[2] = {0, 0};               // ghost records per side
  int64_t send_start[2] = {0, 0}, send_n[2] = {0, 0};  // own edge columns (sorted order): halo sends
  double* mig_d[2] = {nullptr, nullptr};     // migration buffers [send, recv]
  int32_t* mig_i[2] = {nullptr, nullptr};
  int64_t mig_cap = 0;
  int64_t migrated_out = 0, migrated_in = 0; // running totals (instrumentation)
};

namespace {

#define CUDA_OK(call)                                                                 \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                 \
      return CCB_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

#define LAUNCH(kernel, grid, block, ...)                                              \
  do {                                                                                \
    kernel<<<(grid), (block), 0, ctx->stream>>>(__VA_ARGS__);                         \
    ctx->launches++;                                                                  \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) {                                                         \
      ctx->err = std::string(#kernel) + ": " + cudaGetErrorString(e__);               \
      return CCB_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
constexpr int kStageCols = 26;  // widest id-ordered transfer: the whole evolved state

int field_offset(int field) {
  int o = 0;
  for (int f = 0; f < field; ++f) o += kFieldWidth[f];
  return o;
}
int total_width() { return field_offset(CCB_F_COUNT); }

// Phase timing without stalling the stream: every phase records a (start, stop) event pair from a
// pool; the pairs are resolved into per-phase totals when the host next synchronises.
using EventPair = ccb_ctx::EventPair;
struct PhaseTimer {
  ccb_ctx* ctx; int ph; size_t slot;
  PhaseTimer(ccb_ctx* c, int p) : ctx(c), ph(p), slot(0) {
    if (!ctx->profiling) return;
    if (ctx->ev_used == ctx->ev_pool.size()) {
      EventPair e; e.phase = p;
      cudaEventCreate(&e.a); cudaEventCreate(&e.b);
      ctx->ev_pool.push_back(e);
    }
    slot = ctx->ev_used++;
    ctx->ev_pool[slot].phase = p;
    cudaEventRecord(ctx->ev_pool[slot].a, ctx->stream);
  }
  ~PhaseTimer() { if (ctx->profiling) cudaEventRecord(ctx->ev_pool[slot].b, ctx->stream); }
};

void resolve_phase_events(ccb_ctx* ctx) {
  if (ctx->ev_used == 0) return;
  cudaStreamSynchronize(ctx->stream);
  for (size_t k = 0; k < ctx->ev_used; ++k) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev_pool[k].a, ctx->ev_pool[k].b) == cudaSuccess) {
      ctx->phase_ms[ctx->ev_pool[k].phase] += ms;
      ctx->phase_calls[ctx->ev_pool[k].phase]++;
    }
  }
  ctx->ev_used = 0;
}

// The neighbour grid of the reference (src/system_state.cpp:651-700 + Cabana's Cartesian grid), stored in KEY order
// (CellGrid, state.cuh): the slab axis is the slowest digit of the cell key.
void setup_grid(ccb_ctx* ctx) {
  const ccb_config& c = ctx->cfg;
  double min_pos[3] = {c.xmin, c.ymin, c.etamin};
  const double h = c.hT;
  for (int d = 0; d < 3; ++d) min_pos[d] *= 2.;
  if (c.coordinate_system == CCB_COORD_CARTESIAN)
    for (int d = 0; d < 3; ++d) min_pos[d] *= 3.;
  const double radius = 2 * h;
  for (int d = ctx->D; d < 3; ++d) min_pos[d] = -h;
  CellGrid& g = ctx->grid;
  int order[3], no = 0;
  order[no++] = c.slab_axis;
  for (int d = 0; d < 3; ++d) if (d != c.slab_axis) order[no++] = d;
  g.ncells = 1;
  for (int k = 0; k < 3; ++k) {
    const int d = order[k];
    const double max_pos = -min_pos[d];
    const double delta = 1. * radius;
    int nc = (int)std::floor((max_pos - min_pos[d]) * (1.0 / delta));
    if (nc < 1) nc = 1;
    const double dx = (max_pos - min_pos[d]) / nc;
    g.axis[k] = d;
    g.min[k] = min_pos[d];
    g.rdx[k] = 1.0 / dx;
    g.nc[k] = nc;
    g.ncells *= nc;
  }
}
// cell key of the library (digits in key order) -> cell id in the reference's numbering (ix ny + iy) nz + iz
int32_t reference_cell_id(const CellGrid& g, uint32_t key) {
  int c[3], ncs[3];
  uint32_t k = key;
  for (int q = 2; q >= 0; --q) { c[g.axis[q]] = (int)(k % (uint32_t)g.nc[q]); ncs[g.axis[q]] = g.nc[q]; k /= (uint32_t)g.nc[q]; }
  return (int32_t)((c[0] * ncs[1] + c[1]) * ncs[2] + c[2]);
}

int free_particles(ccb_ctx* ctx) {
  if (ctx->hitmask) { cudaFree(ctx->hitmask); ctx->hitmask = nullptr; }
  ctx->hit_valid = false; ctx->hit_cap = 0; ctx->hit_kmax = 0;
  cudaFree(ctx->slab); cudaFree(ctx->y0_slab); cudaFree(ctx->alt_slab); cudaFree(ctx->ints);
  cudaFree(ctx->k_slab); cudaFree(ctx->alt_k_slab); ctx->k_slab = ctx->alt_k_slab = nullptr;
  cudaFree(ctx->key_a); cudaFree(ctx->key_b); cudaFree(ctx->idx_a); cudaFree(ctx->idx_b); cudaFree(ctx->hist);
  cudaFree(ctx->rec); cudaFree(ctx->tasks); cudaFree(ctx->task_pos); cudaFree(ctx->d_stage); cudaFree(ctx->init_tmp);
  cudaFree(ctx->frz_slab); cudaFree(ctx->frz.print); cudaFree(ctx->frz.count);
  ctx->frz_slab = nullptr; ctx->frz = FreezeArrays{};
  ctx->task_pos = nullptr;
  for (int k = 0; k < 2; ++k) { cudaFree(ctx->mig_d[k]); cudaFree(ctx->mig_i[k]); ctx->mig_d[k] = nullptr; ctx->mig_i[k] = nullptr; }
  cudaFree(ctx->mig_list); ctx->mig_list = nullptr;
  ctx->mig_cap = 0;
  ctx->slab = ctx->y0_slab = ctx->alt_slab = nullptr; ctx->ints = nullptr;
  ctx->key_a = ctx->key_b = ctx->idx_a = ctx->idx_b = ctx->hist = nullptr;
  ctx->rec = nullptr; ctx->tasks = nullptr; ctx->d_stage = nullptr; ctx->init_tmp = nullptr;
  return CCB_OK;
}

__global__ void iota_kernel(int32_t* id, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) id[i] = (int32_t)i;
}
__global__ void fill_kernel(double* p, double v, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

int allocate_particles(ccb_ctx* ctx, int64_t n) {
  free_particles(ctx);
  ctx->n = n;
  // with a slab decomposition particles arrive from the neighbours: 25 % + 8192 of slack
  const int64_t want = ctx->cfg.nranks > 1 ? n + n / 4 + 8192 : n;
  ctx->cap = ((want + 63) / 64) * 64;
  if (ctx->cap == 0) ctx->cap = 64;
  const int64_t cap = ctx->cap;
  const int W = total_width();
  int wy0 = 0, wp = 0;
  for (int k = 0; k < kNumEvolvedFields; ++k) wy0 += kFieldWidth[kEvolvedFields[k]];
  for (int f = 0; f < kNumPersistentFields; ++f) wp += kFieldWidth[f];
  CUDA_OK(cudaMalloc(&ctx->slab, sizeof(double) * W * cap));
  CUDA_OK(cudaMalloc(&ctx->y0_slab, sizeof(double) * wy0 * cap));
  CUDA_OK(cudaMalloc(&ctx->alt_slab, sizeof(double) * (wp + wy0) * cap));
  CUDA_OK(cudaMalloc(&ctx->ints, sizeof(int32_t) * 8 * cap));
  CUDA_OK(cudaMemsetAsync(ctx->slab, 0, sizeof(double) * W * cap, ctx->stream));
  CUDA_OK(cudaMemsetAsync(ctx->y0_slab, 0, sizeof(double) * wy0 * cap, ctx->stream));
  CUDA_OK(cudaMemsetAsync(ctx->alt_slab, 0, sizeof(double) * (wp + wy0) * cap, ctx->stream));
  CUDA_OK(cudaMemsetAsync(ctx->ints, 0, sizeof(int32_t) * 8 * cap, ctx->stream));
  DeviceView& dv = ctx->dv;
  dv.n = n; dv.cap = cap;
  int oy = 0, oa = 0;
  for (int f = 0; f < CCB_F_COUNT; ++f) {
    dv.f[f] = ctx->slab + (int64_t)field_offset(f) * cap;
    dv.y0[f] = nullptr; ctx->alt[f] = nullptr; ctx->alt_y0[f] = nullptr;
    if (f < kNumPersistentFields) { ctx->alt[f] = ctx->alt_slab + (int64_t)oa * cap; oa += kFieldWidth[f]; }
  }
  for (int k = 0; k < kNumEvolvedFields; ++k) {
    const int f = kEvolvedFields[k];
    dv.y0[f] = ctx->y0_slab + (int64_t)oy * cap; oy += kFieldWidth[f];
    ctx->alt_y0[f] = ctx->alt_slab + (int64_t)oa * cap; oa += kFieldWidth[f];
  }
  for (int f = 0; f < CCB_F_COUNT; ++f) { ctx->kacc[f] = nullptr; ctx->alt_k[f] = nullptr; }
  if (ctx->cfg.rk_order == 4) {
    CUDA_OK(cudaMalloc(&ctx->k_slab, sizeof(double) * wy0 * cap));
    CUDA_OK(cudaMalloc(&ctx->alt_k_slab, sizeof(double) * wy0 * cap));
    CUDA_OK(cudaMemsetAsync(ctx->k_slab, 0, sizeof(double) * wy0 * cap, ctx->stream));
    CUDA_OK(cudaMemsetAsync(ctx->alt_k_slab, 0, sizeof(double) * wy0 * cap, ctx->stream));
    int ok = 0;
    for (int k = 0; k < kNumEvolvedFields; ++k) {
      const int f = kEvolvedFields[k];
      ctx->kacc[f] = ctx->k_slab + (int64_t)ok * cap; ctx->alt_k[f] = ctx->alt_k_slab + (int64_t)ok * cap; ok += kFieldWidth[f];
    }
  }
  dv.id = ctx->ints; dv.btrack = ctx->ints + cap; dv.freeze = ctx->ints + 2 * cap; dv.eos_id = ctx->ints + 3 * cap;
  for (int k = 0; k < 4; ++k) ctx->alt_int[k] = ctx->ints + (4 + k) * cap;
  CUDA_OK(cudaMalloc(&ctx->key_a, sizeof(uint32_t) * cap)); CUDA_OK(cudaMalloc(&ctx->key_b, sizeof(uint32_t) * cap));
  CUDA_OK(cudaMalloc(&ctx->idx_a, sizeof(uint32_t) * cap)); CUDA_OK(cudaMalloc(&ctx->idx_b, sizeof(uint32_t) * cap));
  ctx->hist_alloc = 256 * cdiv(cap, kSortTile);
  CUDA_OK(cudaMalloc(&ctx->hist, sizeof(uint32_t) * ctx->hist_alloc));
  CUDA_OK(cudaMalloc(&ctx->rec, sizeof(double) * REC_G_MAX * cap));
  ctx->rec_cap = cap;
  CUDA_OK(cudaMalloc(&ctx->tasks, sizeof(int32_t) * cap));
  CUDA_OK(cudaMalloc(&ctx->task_pos, sizeof(uint32_t) * (cap + 1)));
  CUDA_OK(cudaMalloc(&ctx->d_stage, sizeof(double) * kStageCols * cap));
  dv.key = ctx->key_a;
  dv.cell_start = ctx->cell_count;
  dv.rec_fields = ctx->rec; dv.rec_grad = ctx->rec;
  dv.ghost_rec[0] = dv.ghost_rec[1] = ctx->rec;
  dv.tasks = ctx->tasks; dv.err = ctx->d_small + 1;
  const int64_t need_sums = cdiv(std::max<int64_t>(std::max<int64_t>(ctx->hist_alloc, ctx->grid.ncells + 1), cap + 1), kScanTile) + 1;
  if (need_sums > ctx->sums_alloc) {
    cudaFree(ctx->scan_sums);
    CUDA_OK(cudaMalloc(&ctx->scan_sums, sizeof(uint32_t) * need_sums));
    ctx->sums_alloc = need_sums;
  }
  if (n > 0) LAUNCH(iota_kernel, cdiv(n, 256), 256, dv.id, n);
  dv.eos_type_is_table = ctx->cfg.eos_type == CCB_EOS_TABLE ? 1 : 0;
  {  // freeze-out: the counter always (the initial status check runs in every run); the snapshot and results rows when particlization is on
    CUDA_OK(cudaMalloc(&ctx->frz.count, sizeof(unsigned long long)));
    ctx->frzc = 0; ctx->taup = ctx->taupp = 0.0; ctx->number_part_fo = 0;
    if (ctx->cfg.particlization_enabled) {
      if (ctx->cfg.nranks > 1) { ctx->err = "particlization_enabled: freeze-out runs on single-context systems (the snapshot rows do not migrate between slabs)"; return CCB_ERR_INVALID; }
      const size_t rows = (size_t)cap * (6 * FZ_ROW + FR_ROW);
      CUDA_OK(cudaMalloc(&ctx->frz_slab, sizeof(double) * rows));
      CUDA_OK(cudaMemsetAsync(ctx->frz_slab, 0, sizeof(double) * rows, ctx->stream));
      double* q = ctx->frz_slab;
      double** six[6] = {&ctx->frz.frz1, &ctx->frz.frz2, &ctx->frz.fback, &ctx->frz.fback2, &ctx->frz.fback3, &ctx->frz.fback4};
      for (int k = 0; k < 6; ++k) { *six[k] = q; q += (size_t)cap * FZ_ROW; }
      ctx->frz.results = q;
      CUDA_OK(cudaMalloc(&ctx->frz.print, sizeof(int32_t) * cap));
      CUDA_OK(cudaMemsetAsync(ctx->frz.print, 0, sizeof(int32_t) * cap, ctx->stream));
    }
  }
  ctx->cells_valid = false;
  ctx->cells_built = false;
  ctx->have_state = true;
  return CCB_OK;
}

int exclusive_scan(ccb_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n, int write_total) {
  if (n <= 1024 * kScanSmallPer) {
    LAUNCH(scan_small, 1, 1024, in, out, n, write_total);
    return CCB_OK;
  }
  const int64_t nb = cdiv(n, kScanTile);
  LAUNCH(scan_block_sums, nb, kScanThreads, in, ctx->scan_sums, n);
  LAUNCH(scan_sums_inplace, 1, 1024, ctx->scan_sums, nb);
  LAUNCH(scan_apply, nb, kScanThreads, in, out, ctx->scan_sums, n, write_total);
  return CCB_OK;
}

int check_ready(ccb_ctx* ctx, bool need_cells) {
  if (!ctx) return CCB_ERR_INVALID;
  if (!ctx->have_state || ctx->n <= 0) { ctx->err = "no particles: call ccb_allocate or ccb_upload_ic first"; return CCB_ERR_STATE; }
  if (need_cells && !ctx->cells_valid) { ctx->err = "cell list not built: call ccb_build_cells first"; return CCB_ERR_STATE; }
  return CCB_OK;
}

StageParams make_stage(ccb_ctx* ctx, double h_stage, int do_update) {
  return make_stage_params(ctx->cfg, ctx->t, h_stage, do_update);
}

GatherParams make_gather(ccb_ctx* ctx) {
  GatherParams gp;
  gp.kD = ctx->kD; gp.k2 = ctx->k2; gp.k1 = ctx->k1;
  const double radius = 2 * ctx->cfg.hT;
  gp.rsqr = radius * radius;
  gp.rsqr_f = (float)(gp.rsqr * (1.0 + 1e-4));  // fp32 pre-filter: never misses a true neighbour (gather.cuh)
  gp.task_sel = 0; gp.task_info = ctx->d_taskinfo; gp.task_counter = ctx->d_taskinfo + 4;

  gp.t = ctx->t;
  gp.milne = ctx->milne ? 1 : 0;
  gp.nbr_off = nullptr; gp.nbr_idx = nullptr;
  gp.hitmask = nullptr; gp.hit_kmax = 0; gp.hit_stride = 0;
  return gp;
}

// Room for the hit masks of one stage: hit_kmax tiles of 64 candidates per target (tiles beyond that fall back to the pre-filter in
// the gradient gather). A quarter of the free device memory at most; both gathers must cut their tiles the same way.
int ensure_hitmask(ccb_ctx* ctx) {
  if (ctx->hit_off) return CCB_OK;
  const ccb::GatherShape &a = ctx->shape_fields, &b = ctx->shape_grad;
  if (a.nt != b.nt || a.tile != b.tile) { ctx->hit_off = true; return CCB_OK; }
  if (ctx->hitmask && ctx->hit_cap == ctx->cap) return CCB_OK;
  if (ctx->hitmask) { cudaFree(ctx->hitmask); ctx->hitmask = nullptr; }
  ctx->hit_valid = false;
  size_t free_b = 0, total_b = 0;
  CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
  const size_t per_tile = sizeof(unsigned long long) * (size_t)ctx->cap;
  // depth of the mask buffer: 8 GiB worth of tiles, but at least 48 tiles (a 3-D stencil at 64 .. 100 particles per cell needs 27 .. 43)
  // when that is more, and never more than a quarter of the free memory or 1024 tiles
  const int64_t k_quarter = (int64_t)((free_b / 4) / per_tile), k_8g = (int64_t)(((size_t)8 << 30) / per_tile);
  int64_t kmax = std::min<int64_t>(std::min<int64_t>(1024, k_quarter), std::max<int64_t>(k_8g, 48));
  if (kmax < 4) { ctx->hit_off = true; return CCB_OK; }
  // a stage needs ceil(candidates / tile) tiles per task: 27 for a 3-D lattice at h = 2 d, ~140 for the 2-D h = 0.4 Gubser run
  if (cudaMalloc(&ctx->hitmask, per_tile * (size_t)kmax) != cudaSuccess) { cudaGetLastError(); ctx->hitmask = nullptr; ctx->hit_off = true; return CCB_OK; }
  ctx->hit_kmax = (int)kmax;
  ctx->hit_cap = ctx->cap;
  return CCB_OK;
}

constexpr int kReduceBlocks = 296;
// ---- phase implementations ------------------------------------------------------
// (key, index) radix sort of the first n particles + permutation of every persistent column.
// drop_outside: keys outside the slab become the sentinel `ncells` (sorted last, not counted as particles).
int sort_and_permute(ccb_ctx* ctx, int64_t n, int drop_outside) {
  DeviceView& dv = ctx->dv;
  const CellGrid& g = ctx->grid;
  if (g.ncells + 3 > ctx->ncells_alloc) {
    cudaFree(ctx->cell_count);
    CUDA_OK(cudaMalloc(&ctx->cell_count, sizeof(uint32_t) * (g.ncells + 3)));
    ctx->ncells_alloc = g.ncells + 3;
  }
  dv.cell_start = ctx->cell_count;
  const int64_t P = (int64_t)g.nc[1] * g.nc[2];
  const uint32_t key_lo = (uint32_t)(ctx->slab_lo * P), key_hi = (uint32_t)(ctx->slab_hi * P), sentinel = (uint32_t)g.ncells;
  CUDA_OK(cudaMemsetAsync(ctx->cell_count, 0, sizeof(uint32_t) * (g.ncells + 2), ctx->stream));
  LAUNCH(cell_key_kernel, cdiv(n, 256), 256, dv, g, ctx->key_a, ctx->idx_a, ctx->cell_count, n, drop_outside, key_lo, key_hi, sentinel);
  // cell_start[c] for c <= ncells; cell_start[ncells] = number of particles that carry a real key
  int rc = exclusive_scan(ctx, ctx->cell_count, ctx->cell_count, g.ncells + 1, 1);
  if (rc) return rc;
  // LSD radix sort over the significant bits of the key (the sentinel included)
  int bits = 1;
  while (((int64_t)1 << bits) < g.ncells + 1) ++bits;
  const int nblocks = (int)cdiv(n, kSortTile);
  uint32_t *kin = ctx->key_a, *kout = ctx->key_b, *iin = ctx->idx_a, *iout = ctx->idx_b;
  for (int shift = 0; shift < bits; shift += 8) {
    LAUNCH(radix_hist, nblocks, kSortThreads, kin, ctx->hist, n, shift, nblocks);
    rc = exclusive_scan(ctx, ctx->hist, ctx->hist, (int64_t)256 * nblocks, 0);
    if (rc) return rc;
    LAUNCH(radix_scatter, nblocks, kSortThreads, kin, iin, kout, iout, ctx->hist, n, shift, nblocks);
    std::swap(kin, kout); std::swap(iin, iout);
  }
  // gather every persistent column (and the RK snapshot) through the permutation
  PermuteList pl;
  pl.ncols = 0; pl.nints = 0;
  const bool take_y0 = ctx->y0_pending;   // first sort of a step: y0 := the sorted evolved columns, its old content is dead
  ctx->y0_pending = false;
  for (int f = 0; f < kNumPersistentFields; ++f) {
    if (f == CCB_F_EXTENSIVE_DIFFUSION && !ctx->cfg.using_diffusion) continue;
    bool evolved = false;
    for (int k = 0; k < kNumEvolvedFields; ++k) evolved = evolved || kEvolvedFields[k] == f;
    for (int c = 0; c < kFieldWidth[f]; ++c) {
      pl.src[pl.ncols] = dv.f[f] + (int64_t)c * dv.cap; pl.dst[pl.ncols] = ctx->alt[f] + (int64_t)c * dv.cap;
      pl.dst2[pl.ncols] = (take_y0 && evolved) ? ctx->alt_y0[f] + (int64_t)c * dv.cap : nullptr;
      pl.ncols++;
    }
  }
  for (int k = 0; k < kNumEvolvedFields; ++k) {
    const int f = kEvolvedFields[k];
    if (f == CCB_F_EXTENSIVE_DIFFUSION && !ctx->cfg.using_diffusion) continue;
    if (!take_y0)
      for (int c = 0; c < kFieldWidth[f]; ++c) { pl.src[pl.ncols] = dv.y0[f] + (int64_t)c * dv.cap; pl.dst[pl.ncols] = ctx->alt_y0[f] + (int64_t)c * dv.cap; pl.dst2[pl.ncols] = nullptr; pl.ncols++; }
    if (ctx->k_slab)
      for (int c = 0; c < kFieldWidth[f]; ++c) { pl.src[pl.ncols] = ctx->kacc[f] + (int64_t)c * dv.cap; pl.dst[pl.ncols] = ctx->alt_k[f] + (int64_t)c * dv.cap; pl.dst2[pl.ncols] = nullptr; pl.ncols++; }
  }
  int32_t* icur[4] = {dv.id, dv.btrack, dv.freeze, dv.eos_id};
  for (int k = 0; k < 4; ++k) { pl.isrc[k] = icur[k]; pl.idst[k] = ctx->alt_int[k]; pl.nints++; }
  LAUNCH(permute_columns, cdiv(n, 256), 256, pl, iin, n);
  for (int f = 0; f < kNumPersistentFields; ++f) {
    if (f == CCB_F_EXTENSIVE_DIFFUSION && !ctx->cfg.using_diffusion) continue;
    std::swap(dv.f[f], ctx->alt[f]);
  }
  for (int k = 0; k < kNumEvolvedFields; ++k) {
    const int f = kEvolvedFields[k];
    if (f == CCB_F_EXTENSIVE_DIFFUSION && !ctx->cfg.using_diffusion) continue;
    std::swap(dv.y0[f], ctx->alt_y0[f]);
    if (ctx->k_slab) std::swap(ctx->kacc[f], ctx->alt_k[f]);
  }
  std::swap(dv.id, ctx->alt_int[0]); std::swap(dv.btrack, ctx->alt_int[1]); std::swap(dv.freeze, ctx->alt_int[2]);
  std::swap(dv.eos_id, ctx->alt_int[3]);
  dv.key = kin;  // sorted keys
  if (kin != ctx->key_a) { std::swap(ctx->key_a, ctx->key_b); std::swap(ctx->idx_a, ctx->idx_b); }
  return CCB_OK;
}

#define COMM_OK(call)                                                                 \
  do {                                                                                \
    if ((call)) { ctx->err = std::string(ctx->comm->name()) + " transport: " + ctx->comm->err; return CCB_ERR_COMM; } \
  } while (0)

// every column that travels with a particle (the permuted set), in its CURRENT buffer
void migrate_list(ccb_ctx* ctx, MigrateList& ml) {
  DeviceView& dv = ctx->dv;
  ml.ncols = 0;
  for (int f = 0; f < kNumPersistentFields; ++f) {
    if (f == CCB_F_EXTENSIVE_DIFFUSION && !ctx->cfg.using_diffusion) continue;
    for (int c = 0; c < kFieldWidth[f]; ++c) ml.col[ml.ncols++] = dv.f[f] + (int64_t)c * dv.cap;
  }
  for (int k = 0; k < kNumEvolvedFields; ++k) {
    const int f = kEvolvedFields[k];
    if (f == CCB_F_EXTENSIVE_DIFFUSION && !ctx->cfg.using_diffusion) continue;
    for (int c = 0; c < kFieldWidth[f]; ++c) ml.col[ml.ncols++] = dv.y0[f] + (int64_t)c * dv.cap;
    if (ctx->k_slab)
      for (int c = 0; c < kFieldWidth[f]; ++c) ml.col[ml.ncols++] = ctx->kacc[f] + (int64_t)c * dv.cap;
  }
  ml.icol[0] = dv.id; ml.icol[1] = dv.btrack; ml.icol[2] = dv.freeze; ml.icol[3] = dv.eos_id;
}

// Particles whose cell column left the slab go to the neighbouring rank. They are found and shipped BEFORE the stage's
// sort: the arrivals are appended behind the own particles and the one sort that follows (drop_outside) drops the leavers
// and files the arrivals. Returns the number of rows to sort in *n_sort and the leavers among them in *n_drop.
int migrate(ccb_ctx* ctx, int64_t* n_sort, int64_t* n_drop) {
  const CellGrid& g = ctx->grid;
  const int rank = ctx->cfg.rank, nranks = ctx->cfg.nranks;
  const int left = rank - 1, right = rank + 1 < nranks ? rank + 1 : -1;
  const int64_t n = ctx->n;
  *n_sort = n; *n_drop = 0;
  int32_t* info = ctx->d_info;  // [0] out left, [1] out right, [2] in from left, [3] in from right
  uint32_t *list_l = ctx->key_b, *list_r = ctx->idx_b;  // free until the sort
  CUDA_OK(cudaMemsetAsync(info, 0, sizeof(int32_t) * 4, ctx->stream));
  LAUNCH(find_leavers, cdiv(n, 256), 256, ctx->dv.f[CCB_F_POSITION] + (int64_t)g.axis[0] * ctx->cap, g.min[0], g.rdx[0], g.nc[0], ctx->slab_lo, ctx->slab_hi,
         left >= 0 ? 1 : 0, right >= 0 ? 1 : 0, n, list_l, list_r, info);
  COMM_OK(ctx->comm->group_begin());
  if (left >= 0) COMM_OK(ctx->comm->send(info + 0, sizeof(int32_t), left, ctx->stream));
  if (right >= 0) COMM_OK(ctx->comm->send(info + 1, sizeof(int32_t), right, ctx->stream));
  if (left >= 0) COMM_OK(ctx->comm->recv(info + 2, sizeof(int32_t), left, ctx->stream));
  if (right >= 0) COMM_OK(ctx->comm->recv(info + 3, sizeof(int32_t), right, ctx->stream));
  COMM_OK(ctx->comm->group_end());
  int32_t h[4];
  CUDA_OK(cudaMemcpyAsync(h, info, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  const int64_t outL = h[0], outR = h[1], inL = h[2], inR = h[3];
  if (outL + outR + inL + inR == 0) return CCB_OK;
  // A slab that cannot hold its arrivals still takes part in the payload exchange below (the rows land in the migration buffers,
  // not in the particle columns), so that no neighbour is left blocked in a receive; the error is returned after it.
  const bool overflow = n + inL + inR > ctx->cap;
  MigrateList ml;
  migrate_list(ctx, ml);
  const int64_t need = std::max<int64_t>(std::max(outL + outR, inL + inR), 1);
  if (need > ctx->mig_cap) {
    for (int k = 0; k < 2; ++k) { cudaFree(ctx->mig_d[k]); cudaFree(ctx->mig_i[k]); }
    cudaFree(ctx->mig_list);
    ctx->mig_cap = need * 2 + 1024;
    for (int k = 0; k < 2; ++k) {
      CUDA_OK(cudaMalloc(&ctx->mig_d[k], sizeof(double) * ml.ncols * ctx->mig_cap));
      CUDA_OK(cudaMalloc(&ctx->mig_i[k], sizeof(int32_t) * 4 * ctx->mig_cap));
    }
    CUDA_OK(cudaMalloc(&ctx->mig_list, sizeof(uint32_t) * ctx->mig_cap));
  }
  double *sd = ctx->mig_d[0], *rd = ctx->mig_d[1];
  int32_t *si = ctx->mig_i[0], *ri = ctx->mig_i[1];
  uint32_t* sorted = ctx->mig_list;
  if (outL) {
    LAUNCH(rank_sort_u32, cdiv(outL, 256), 256, list_l, sorted, (int)outL);
    LAUNCH(pack_rows_indexed, cdiv(outL, 128), 128, ml, sorted, outL, sd, si);
  }
  if (outR) {
    LAUNCH(rank_sort_u32, cdiv(outR, 256), 256, list_r, sorted + outL, (int)outR);
    LAUNCH(pack_rows_indexed, cdiv(outR, 128), 128, ml, sorted + outL, outR, sd + ml.ncols * outL, si + 4 * outL);
  }
  COMM_OK(ctx->comm->group_begin());
  if (left >= 0 && outL) { COMM_OK(ctx->comm->send(sd, sizeof(double) * ml.ncols * outL, left, ctx->stream)); COMM_OK(ctx->comm->send(si, sizeof(int32_t) * 4 * outL, left, ctx->stream)); }
  if (right >= 0 && outR) { COMM_OK(ctx->comm->send(sd + ml.ncols * outL, sizeof(double) * ml.ncols * outR, right, ctx->stream)); COMM_OK(ctx->comm->send(si + 4 * outL, sizeof(int32_t) * 4 * outR, right, ctx->stream)); }
  if (left >= 0 && inL) { COMM_OK(ctx->comm->recv(rd, sizeof(double) * ml.ncols * inL, left, ctx->stream)); COMM_OK(ctx->comm->recv(ri, sizeof(int32_t) * 4 * inL, left, ctx->stream)); }
  if (right >= 0 && inR) { COMM_OK(ctx->comm->recv(rd + ml.ncols * inL, sizeof(double) * ml.ncols * inR, right, ctx->stream)); COMM_OK(ctx->comm->recv(ri + 4 * inL, sizeof(int32_t) * 4 * inR, right, ctx->stream)); }
  COMM_OK(ctx->comm->group_end());
  if (overflow) {
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    ctx->err = "slab overflow: " + std::to_string(n + inL + inR) + " particles after migration, capacity " + std::to_string(ctx->cap) +
               " (the slab decomposition is static: cut the slabs again, or upload with fewer particles per rank)";
    return CCB_ERR_STATE;
  }
  if (inL) LAUNCH(unpack_rows, cdiv(inL, 128), 128, ml, n, inL, rd, ri);
  if (inR) LAUNCH(unpack_rows, cdiv(inR, 128), 128, ml, n + inL, inR, rd + ml.ncols * inL, ri + 4 * inL);
  *n_sort = n + inL + inR;
  *n_drop = outL + outR;
  ctx->migrated_out += outL + outR;
  ctx->migrated_in += inL + inR;
  return CCB_OK;
}

// Exchange the cell tables of the edge columns with the two neighbours: afterwards this rank knows which of its
// own records each neighbour needs (send_start / send_n) and where the ghost records it receives will sit.
int exchange_column_tables(ccb_ctx* ctx) {
  const CellGrid& g = ctx->grid;
  DeviceView& dv = ctx->dv;
  const int rank = ctx->cfg.rank, nranks = ctx->cfg.nranks;
  const int64_t P = (int64_t)g.nc[1] * g.nc[2], T = P + 1;
  const int left = rank - 1, right = rank + 1 < nranks ? rank + 1 : -1;
  uint32_t* tab = ctx->ghost_tab;  // [0] ghost table left, [1] ghost table right, [2] recv scratch left, [3] recv scratch right
  const uint32_t* own_first = dv.cell_start + (int64_t)ctx->slab_lo * P;
  const uint32_t* own_last = dv.cell_start + (int64_t)(ctx->slab_hi - 1) * P;
  COMM_OK(ctx->comm->group_begin());
  if (left >= 0) COMM_OK(ctx->comm->send(own_first, sizeof(uint32_t) * T, left, ctx->stream));
  if (right >= 0) COMM_OK(ctx->comm->send(own_last, sizeof(uint32_t) * T, right, ctx->stream));
  if (left >= 0) COMM_OK(ctx->comm->recv(tab + 2 * T, sizeof(uint32_t) * T, left, ctx->stream));
  if (right >= 0) COMM_OK(ctx->comm->recv(tab + 3 * T, sizeof(uint32_t) * T, right, ctx->stream));
  COMM_OK(ctx->comm->group_end());
  uint32_t h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // own first column [start, end], own last column, left table ends, right table ends, rows with a real key
  LAUNCH(table_ends, 1, 32, own_first, own_last, tab + 2 * T, tab + 3 * T, P, left >= 0 ? 1 : 0, right >= 0 ? 1 : 0, dv.cell_start + g.ncells, ctx->d_ends);
  CUDA_OK(cudaMemcpyAsync(h, ctx->d_ends, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  // cell_start[ncells] = rows that carry a key inside this slab. It must equal the count the migration arithmetic arrived at: an
  // arrival whose column lies outside the receiving slab (a particle that crossed a whole slab in one stage, or a slab changed
  // after the upload) got the sentinel key too and would otherwise be lost silently.
  if ((int64_t)h[8] != ctx->n) {
    ctx->err = "slab migration: " + std::to_string(ctx->n - (int64_t)h[8]) + " particle(s) arrived outside the receiving slab (moved more than one slab per stage?)";
    return CCB_ERR_STATE;
  }
  ctx->send_start[0] = h[0]; ctx->send_n[0] = left >= 0 ? (int64_t)h[1] - h[0] : 0;
  ctx->send_start[1] = h[2]; ctx->send_n[1] = right >= 0 ? (int64_t)h[3] - h[2] : 0;
  ctx->ghost_n[0] = left >= 0 ? (int64_t)h[5] - h[4] : 0;
  ctx->ghost_n[1] = right >= 0 ? (int64_t)h[7] - h[6] : 0;
  dv.ghost_col[0] = left >= 0 ? ctx->slab_lo - 1 : -1;
  dv.ghost_col[1] = right >= 0 ? ctx->slab_hi : -1;
  if (ctx->p2p) {
    // the ghost records stay where they are: the received tables already hold offsets into the neighbour's own record array
    dv.ghost_start[0] = tab + 2 * T; dv.ghost_start[1] = tab + 3 * T;
    dv.ghost_rec[0] = static_cast<const double*>(ctx->peer_rec[0]); dv.ghost_rec[1] = static_cast<const double*>(ctx->peer_rec[1]);
    return CCB_OK;
  }
  if (left >= 0) LAUNCH(rebase_ghost_table, cdiv(T, 256), 256, tab + 2 * T, tab, T, (uint32_t)ctx->n);
  if (right >= 0) LAUNCH(rebase_ghost_table, cdiv(T, 256), 256, tab + 3 * T, tab + T, T, (uint32_t)(ctx->n + ctx->ghost_n[0]));
  dv.ghost_start[0] = tab; dv.ghost_start[1] = tab + T;
  // room for the ghost records behind the own ones
  const int64_t need = ctx->n + ctx->ghost_n[0] + ctx->ghost_n[1];
  if (need > ctx->rec_cap) {
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->rec);
    ctx->rec_cap = need + need / 8 + 1024;
    CUDA_OK(cudaMalloc(&ctx->rec, sizeof(double) * REC_G_MAX * ctx->rec_cap));
    dv.rec_fields = ctx->rec; dv.rec_grad = ctx->rec;
  }
  dv.ghost_rec[0] = dv.ghost_rec[1] = ctx->rec;
  return CCB_OK;
}

int ensure_comm_stream(ccb_ctx* ctx);
int p2p_wait_readers(ccb_ctx* ctx);

// Ghost records of one gather (rec_doubles per record): own edge columns -> neighbours, theirs -> behind our own.
// The exchange runs on the context's communication stream: it starts when the records are packed (ev_pack) and the
// boundary tasks of the gather follow it on that stream, while the interior tasks run next to both on the compute stream.
int exchange_halo(ccb_ctx* ctx, int rec_doubles) {
  const int rank = ctx->cfg.rank, nranks = ctx->cfg.nranks;
  const int left = rank - 1, right = rank + 1 < nranks ? rank + 1 : -1;
  { const int rc0 = ensure_comm_stream(ctx); if (rc0) return rc0; }
  cudaStream_t cs = ctx->comm_stream;
  CUDA_OK(cudaEventRecord(ctx->ev_pack, ctx->stream));
  CUDA_OK(cudaStreamWaitEvent(cs, ctx->ev_pack, 0));
  double* rec = ctx->rec;
  const size_t rb = sizeof(double) * rec_doubles;
  COMM_OK(ctx->comm->group_begin());
  if (left >= 0 && ctx->send_n[0]) COMM_OK(ctx->comm->send(rec + ctx->send_start[0] * rec_doubles, rb * ctx->send_n[0], left, cs));
  if (right >= 0 && ctx->send_n[1]) COMM_OK(ctx->comm->send(rec + ctx->send_start[1] * rec_doubles, rb * ctx->send_n[1], right, cs));
  if (left >= 0 && ctx->ghost_n[0]) COMM_OK(ctx->comm->recv(rec + ctx->n * rec_doubles, rb * ctx->ghost_n[0], left, cs));
  if (right >= 0 && ctx->ghost_n[1]) COMM_OK(ctx->comm->recv(rec + (ctx->n + ctx->ghost_n[0]) * rec_doubles, rb * ctx->ghost_n[1], right, cs));
  COMM_OK(ctx->comm->group_end());
  return CCB_OK;
}

int grad_record_doubles(ccb_ctx* ctx) {
  const bool df = ctx->cfg.using_diffusion;
  if (ctx->D == 2) return df ? GradRec<2, true>::REC : GradRec<2, false>::REC;
  return df ? GradRec<3, true>::REC : GradRec<3, false>::REC;
}

// The task list of the gathers (cells.cuh): chunks of <= task_nt particles of one cell, in cell order.
int build_task_list(ccb_ctx* ctx) {
  const CellGrid& g = ctx->grid;
  DeviceView& dv = ctx->dv;
  const int64_t n = ctx->n, P = (int64_t)g.nc[1] * g.nc[2];
  LAUNCH(task_flag, cdiv(n, 256), 256, dv.key, dv.cell_start, ctx->task_pos, n, (uint32_t)ctx->shape_grad.nt);
  int rc = exclusive_scan(ctx, ctx->task_pos, ctx->task_pos, n, 1);
  if (rc) return rc;
  const bool has_l = ctx->comm && ctx->cfg.rank > 0, has_r = ctx->comm && ctx->cfg.rank + 1 < ctx->cfg.nranks;
  const uint32_t* lo = has_l ? dv.cell_start + (int64_t)(ctx->slab_lo + 1) * P : nullptr;   // end of the first own column
  const uint32_t* hi = has_r ? dv.cell_start + (int64_t)(ctx->slab_hi - 1) * P : nullptr;   // start of the last own column
  LAUNCH(task_scatter, cdiv(n, 256), 256, ctx->task_pos, ctx->tasks, n, lo, hi, ctx->d_taskinfo);
  return CCB_OK;
}

int do_build_cells(ccb_ctx* ctx) {
  PhaseTimer pt(ctx, PH_NEIGHBOUR);
  int rc;
  ctx->hit_valid = false;
  if (ctx->comm) {
    int64_t n_sort = ctx->n, n_drop = 0;
    { PhaseTimer pm(ctx, PH_SLAB_MIGRATE); rc = migrate(ctx, &n_sort, &n_drop); }
    if (rc) return rc;
    ctx->dv.n = n_sort;
    if ((rc = sort_and_permute(ctx, n_sort, 1))) return rc;   // leavers -> sentinel key -> sorted last -> dropped
    ctx->n = n_sort - n_drop;
    ctx->dv.n = ctx->n;
    { PhaseTimer pc(ctx, PH_SLAB_TABLES); rc = exchange_column_tables(ctx); }
    if (rc) return rc;
  } else {
    ctx->dv.n = ctx->n;
    if ((rc = sort_and_permute(ctx, ctx->n, 0))) return rc;
  }
  if ((rc = build_task_list(ctx))) return rc;
  ctx->cells_valid = true;
  ctx->cells_built = true;
  return CCB_OK;
}

// Dispatch over (dimension, metric) to the per-instantiation object files (launchers.h).
#define LAUNCHER(what, call)                                                          \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    ctx->launches++;                                                                  \
    if (e__ != cudaSuccess) {                                                         \
      ctx->err = std::string(what) + ": " + cudaGetErrorString(e__);                  \
      return CCB_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)
#define DISPATCH_DM(what, fn, ...)                                                    \
  do {                                                                                \
    if (ctx->D == 2) { if (ctx->milne) LAUNCHER(what, (fn<2, true>(__VA_ARGS__))); else LAUNCHER(what, (fn<2, false>(__VA_ARGS__))); } \
    else { if (ctx->milne) LAUNCHER(what, (fn<3, true>(__VA_ARGS__))); else LAUNCHER(what, (fn<3, false>(__VA_ARGS__))); }             \
  } while (0)

int do_prologue(ccb_ctx* ctx) {
  PhaseTimer pt(ctx, PH_GAMMA);
  { const int rc = p2p_wait_readers(ctx); if (rc) return rc; }
  CUDA_OK(cudaMemsetAsync(ctx->d_qflag + 2, 0, sizeof(int), ctx->stream));
  DISPATCH_DM("prologue_kernel", launch_prologue, ctx->stream, ctx->dv, ctx->t);
  return CCB_OK;
}

int ensure_comm_stream(ccb_ctx* ctx) {
  if (ctx->comm_stream) return CCB_OK;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  CUDA_OK(cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, hi));
  CUDA_OK(cudaEventCreateWithFlags(&ctx->ev_pack, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&ctx->ev_halo, cudaEventDisableTiming));
  return CCB_OK;
}

// Before this rank overwrites its record array (the next pack), both neighbours must have finished reading it.
int p2p_wait_readers(ccb_ctx* ctx) {
  if (!ctx->comm || !ctx->p2p || ctx->halo_epoch == 0) return CCB_OK;
  const int rank = ctx->cfg.rank, nranks = ctx->cfg.nranks;
  LAUNCH(p2p_wait, 1, 1, rank > 0 ? ctx->d_flags + 2 : nullptr, rank + 1 < nranks ? ctx->d_flags + 3 : nullptr, ctx->halo_epoch, (int*)nullptr, ctx->dv.err);
  return CCB_OK;
}

// Slab gather WITHOUT a halo message: the boundary tasks stage the ghost column straight out of the neighbour's record array
// (NVLink peer memory). Per gather (one epoch): this rank tells its neighbours that its records are packed (ready flag, with its
// sparsity bit), waits on the communication stream for theirs, runs the boundary tasks there, and tells them it has finished
// reading (done flag); the interior tasks run meanwhile on the compute stream.
template <class F>
int slab_gather_p2p(ccb_ctx* ctx, int rec_doubles, int nt, GatherParams gp, F&& launch) {
  int rc = ensure_comm_stream(ctx);
  if (rc) return rc;
  const int rank = ctx->cfg.rank, nranks = ctx->cfg.nranks;
  const bool has_l = rank > 0, has_r = rank + 1 < nranks;
  const bool fields = rec_doubles == REC_F;
  int* own_flag = ctx->d_qflag + (fields ? 2 : 0);
  int* gflag = own_flag + 1;
  const int e = ++ctx->halo_epoch;
  int* pl = static_cast<int*>(ctx->peer_flags[0]);
  int* pr = static_cast<int*>(ctx->peer_flags[1]);
  cudaStream_t cs = ctx->comm_stream;
  CUDA_OK(cudaMemsetAsync(gflag, 0, sizeof(int), ctx->stream));
  LAUNCH(p2p_signal, 1, 1, has_l ? pl + 1 : nullptr, has_r ? pr + 0 : nullptr, e, (const int*)own_flag);   // I am left's right neighbour, right's left one
  CUDA_OK(cudaEventRecord(ctx->ev_pack, ctx->stream));
  CUDA_OK(cudaStreamWaitEvent(cs, ctx->ev_pack, 0));
  p2p_wait<<<1, 1, 0, cs>>>(has_l ? ctx->d_flags + 0 : nullptr, has_r ? ctx->d_flags + 1 : nullptr, e, gflag, ctx->dv.err);
  ctx->launches++;
  {  // boundary tasks: the slab's first and last cell column (task counts and ranges live on the device)
    GatherParams gb = gp;
    gb.task_sel = 2; gb.task_counter = ctx->d_taskinfo + 5;
    if ((rc = launch(gb, cs))) return rc;
  }
  p2p_signal<<<1, 1, 0, cs>>>(has_l ? pl + 3 : nullptr, has_r ? pr + 2 : nullptr, e, (const int*)nullptr);
  ctx->launches++;
  CUDA_OK(cudaEventRecord(ctx->ev_halo, cs));
  gp.task_sel = 1;
  if ((rc = launch(gp, ctx->stream))) return rc;
  CUDA_OK(cudaStreamWaitEvent(ctx->stream, ctx->ev_halo, 0));
  return CCB_OK;
}

// One gather over all tasks, or (slab run): the halo exchange followed by the boundary tasks on the communication stream,
// the interior tasks next to them on the compute stream, which then waits for the boundary tasks.
template <class F>
int slab_gather(ccb_ctx* ctx, int rec_doubles, int nt, GatherParams gp, F&& launch) {
  CUDA_OK(cudaMemsetAsync(ctx->d_taskinfo + 4, 0, 2 * sizeof(int32_t), ctx->stream));  // task counters of this gather's launches
  if (!ctx->comm) return launch(gp, ctx->stream);
  if (ctx->p2p) return slab_gather_p2p(ctx, rec_doubles, nt, gp, launch);
  // do the ghost records carry what the own records may lack (q^mu_a != 0 for the gradient gather, charges for the fields)?
  const bool fields = rec_doubles == REC_F;
  const bool scan = fields || ctx->cfg.using_diffusion;
  int* gflag = ctx->d_qflag + (fields ? 3 : 1);
  if (scan) CUDA_OK(cudaMemsetAsync(gflag, 0, sizeof(int), ctx->stream));
  int rc = exchange_halo(ctx, rec_doubles);
  if (rc) return rc;
  if (scan) {
    const int64_t ng = ctx->ghost_n[0] + ctx->ghost_n[1];
    const int lo = fields ? 5 : (ctx->D == 2 ? GradRec<2, true>::Q0 : GradRec<3, true>::Q0);
    const int hi = fields ? 8 : (ctx->D == 2 ? GradRec<2, true>::FULL : GradRec<3, true>::FULL);
    if (ng > 0) {
      ghost_q_flag<<<(unsigned)cdiv(ng, 256), 256, 0, ctx->comm_stream>>>(ctx->rec + ctx->n * rec_doubles, ng, rec_doubles, lo, hi, gflag);
      ctx->launches++;
    }
  }
  {
    GatherParams gb = gp;
    gb.task_sel = 2; gb.task_counter = ctx->d_taskinfo + 5;
    if ((rc = launch(gb, ctx->comm_stream))) return rc;
  }
  CUDA_OK(cudaEventRecord(ctx->ev_halo, ctx->comm_stream));
  gp.task_sel = 1;
  if ((rc = launch(gp, ctx->stream))) return rc;
  CUDA_OK(cudaStreamWaitEvent(ctx->stream, ctx->ev_halo, 0));
  return CCB_OK;
}

int do_gather_fields(ccb_ctx* ctx) {
  PhaseTimer pt(ctx, PH_SMOOTH_FIELDS);
  const bool fact = ctx->D == 3 && ctx->cfg.factorized_kernel;
  int rc = ensure_hitmask(ctx);
  if (rc) return rc;
  GatherParams g0 = make_gather(ctx);
  ctx->hit_valid = false;
  if (ctx->hitmask) { g0.hitmask = ctx->hitmask; g0.hit_kmax = ctx->hit_kmax; g0.hit_stride = ctx->hit_cap; }
  rc = slab_gather(ctx, REC_F, ctx->shape_fields.nt, g0, [&](const GatherParams& gp, cudaStream_t st) -> int {
    if (ctx->D == 2) LAUNCHER("gather_kernel<fields>", launch_gather_fields<2>(false, ctx->shape_fields, st, ctx->dv, ctx->grid, gp));
    else LAUNCHER("gather_kernel<fields>", launch_gather_fields<3>(fact, ctx->shape_fields, st, ctx->dv, ctx->grid, gp));
    return CCB_OK;
  });
  if (!rc && ctx->hitmask) ctx->hit_valid = true;   // valid until the cells are rebuilt (do_build_cells)
  return rc;
}

int do_eos_local(ccb_ctx* ctx, int do_eos, int do_local) {
  StageParams sp = make_stage(ctx, 0.0, 0);
  if (!ctx->cfg.online_inverter_enabled) {
    // pre-inverted EoS: phase 04 is a table lookup (EoS_Interpolator::fill_thermodynamics), then the streaming local update
    if (!ctx->have_pre) { ctx->err = "online_inverter_enabled = 0: no pre-inverted table (ccb_set_preinverted_table / ccb_set_preinverted_flat)"; return CCB_ERR_STATE; }
    if (do_eos) {
      PhaseTimer pt(ctx, PH_THERMO);
      LAUNCHER("pre_fill_kernel", launch_pre_fill(ctx->stream, ctx->dv, ctx->pre_table));
    }
    if (do_local) {
      PhaseTimer pt(ctx, PH_LOCAL);
      { const int rc = p2p_wait_readers(ctx); if (rc) return rc; }
      sp.do_eos = 0; sp.do_local = 1;
      CUDA_OK(cudaMemsetAsync(ctx->d_qflag, 0, sizeof(int), ctx->stream));
      DISPATCH_DM("eos_local_kernel", launch_eos_local, ctx->stream, ctx->dv, sp);
    }
    return CCB_OK;
  }
  if (ctx->use_chain) {
    // tabulated / chain EoS: the root search has its own kernel (register-heavy, divergent); the streaming
    // local update then reads the thermodynamics it stored
    if (!ctx->have_table) { ctx->err = "eos_type = table: no table loaded (ccb_set_eos_table / ccb_load_eos_table_dat)"; return CCB_ERR_STATE; }
    if (do_eos) {
      PhaseTimer pt(ctx, PH_THERMO);
      LAUNCHER("eos_chain_stage_kernel", launch_eos_chain_stage(ctx->stream, ctx->dv, ctx->chain, ctx->task_pos, ctx->d_stage));   // task_pos, d_stage: free during a stage
      ctx->launches += 2;
    }
    if (do_local) {
      PhaseTimer pt(ctx, PH_LOCAL);
      { const int rc = p2p_wait_readers(ctx); if (rc) return rc; }
      sp.do_eos = 0; sp.do_local = 1;
      CUDA_OK(cudaMemsetAsync(ctx->d_qflag, 0, sizeof(int), ctx->stream));
      DISPATCH_DM("eos_local_kernel", launch_eos_local, ctx->stream, ctx->dv, sp);
    }
    return CCB_OK;
  }
  PhaseTimer pt(ctx, do_eos ? PH_THERMO : PH_LOCAL);
  sp.do_eos = do_eos; sp.do_local = do_local;
  if (do_local) { const int rc = p2p_wait_readers(ctx); if (rc) return rc; }
  if (do_local) CUDA_OK(cudaMemsetAsync(ctx->d_qflag, 0, sizeof(int), ctx->stream));
  DISPATCH_DM("eos_local_kernel", launch_eos_local, ctx->stream, ctx->dv, sp);
  return CCB_OK;
}

int do_gather_gradients(ccb_ctx* ctx) {
  PhaseTimer pt(ctx, PH_SMOOTH_GRAD);
  const bool fact = ctx->D == 3 && ctx->cfg.factorized_kernel;
  const bool sh = ctx->cfg.using_shear, df = ctx->cfg.using_diffusion;
  GatherParams g0 = make_gather(ctx);
  if (ctx->hit_valid && ctx->hitmask && ctx->hit_cap == ctx->cap) { g0.hitmask = ctx->hitmask; g0.hit_kmax = ctx->hit_kmax; g0.hit_stride = ctx->hit_cap; }
  return slab_gather(ctx, grad_record_doubles(ctx), ctx->shape_grad.nt, g0, [&](const GatherParams& gp, cudaStream_t st) -> int {
    if (ctx->D == 2) LAUNCHER("gather_kernel<gradients>", launch_gather_gradients<2>(sh, df, false, ctx->shape_grad, st, ctx->dv, ctx->grid, gp));
    else LAUNCHER("gather_kernel<gradients>", launch_gather_gradients<3>(sh, df, fact, ctx->shape_grad, st, ctx->dv, ctx->grid, gp));
    if (df) ctx->launches++;  // the two sparsity twins of a diffusion run
    return CCB_OK;
  });
}

int do_eom(ccb_ctx* ctx, double h_stage, int do_update) {
  PhaseTimer pt(ctx, PH_TIME_DERIV);
  const StageParams sp = make_stage(ctx, h_stage, do_update);
  DISPATCH_DM("eom_rk_kernel", launch_eom_rk, ctx->stream, ctx->dv, sp);
  if (do_update) ctx->cells_valid = false;   // positions moved: the sort order and cell table are stale
  return CCB_OK;
}

int derivative_evaluation(ccb_ctx* ctx, double h_stage, int do_update) {
  int rc;
  if ((rc = do_build_cells(ctx))) return rc;
  if ((rc = do_prologue(ctx))) return rc;
  if ((rc = do_gather_fields(ctx))) return rc;
  if ((rc = do_eos_local(ctx, 1, 1))) return rc;
  if ((rc = do_gather_gradients(ctx))) return rc;
  return do_eom(ctx, h_stage, do_update);
}

// ---- rk_order 4 (Evolver::advance_timestep_rk4, src/evolver.cpp:508-561) ----------------------------------------------
struct RkCols {
  int n, s_col, nzero;
  double* y[40];          // evolved column
  const double* y0[40];   // its snapshot
  double* k[40];          // its accumulated weighted derivative
  const double* f[40];    // the stage derivative that feeds it
  double* zero[4];        // components >= D of position and u: kept at zero
};
// update_k (src/evolver.cpp:172-211): k += w * f
__global__ void rk4_accumulate_kernel(RkCols c, double w, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int q = 0; q < c.n; ++q) c.k[q][i] += w * c.f[q][i];
}
// update_rk4 (src/evolver.cpp:216-339): y = y0 + dt * k, with the negative-entropy guard of every RK update
__global__ void rk4_final_kernel(RkCols c, double dt, int64_t n, const int32_t* __restrict__ freeze, const int32_t* __restrict__ id,
                                 int32_t* __restrict__ err) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int q = 0; q < c.n; ++q) {
    double v = c.y0[q][i] + dt * c.k[q][i];
    if (q == c.s_col && v < 0.0) {
      if (freeze[i] > 3) v = 1.e-3;
      else if (atomicCAS(&err[0], 0, (int)CCB_ERR_NEG_ENTROPY) == 0) err[1] = id[i];
    }
    c.y[q][i] = v;
  }
  for (int q = 0; q < c.nzero; ++q) c.zero[q][i] = 0.0;
}
RkCols rk_columns(ccb_ctx* ctx) {
  RkCols c;
  c.n = 0; c.nzero = 0; c.s_col = -1;
  const DeviceView& dv = ctx->dv;
  const int64_t cap = dv.cap;
  auto add = [&](int fy, int cy, int ff, int cf) {
    c.y[c.n] = dv.f[fy] + (int64_t)cy * cap; c.y0[c.n] = dv.y0[fy] + (int64_t)cy * cap; c.k[c.n] = ctx->kacc[fy] + (int64_t)cy * cap;
    c.f[c.n] = dv.f[ff] + (int64_t)cf * cap; c.n++;
  };
  for (int d = 0; d < ctx->D; ++d) { add(CCB_F_POSITION, d, CCB_F_V, d); add(CCB_F_U, d, CCB_F_DU_DT, d); }
  for (int d = ctx->D; d < 3; ++d) { c.zero[c.nzero++] = dv.f[CCB_F_POSITION] + (int64_t)d * cap; c.zero[c.nzero++] = dv.f[CCB_F_U] + (int64_t)d * cap; }
  c.s_col = c.n;
  for (int q = 0; q < 4; ++q) add(CCB_F_EXTENSIVE, q, CCB_F_D_EXTENSIVE_DT, q);
  add(CCB_F_EXTENSIVE_BULK, 0, CCB_F_D_BULK_DT, 0);
  for (int q = 0; q < 5; ++q) add(CCB_F_EXTENSIVE_SHV, q, CCB_F_D_SHV_DT, q);
  if (ctx->cfg.using_diffusion) for (int q = 0; q < 9; ++q) add(CCB_F_EXTENSIVE_DIFFUSION, q, CCB_F_D_DIFFUSION_DT, q);
  add(CCB_F_EZ, 0, CCB_F_DEZ, 0);
  return c;
}

// FreezeOut::check_freeze_out_status (include/freeze_out.h:462-541); *count = particles that are freezing out (status 3) now
int do_freeze_check_status(ccb_ctx* ctx, int64_t* count) {
  CUDA_OK(cudaMemsetAsync(ctx->frz.count, 0, sizeof(unsigned long long), ctx->stream));
  LAUNCHER("freeze_check_status_kernel", launch_freeze_check_status(ctx->stream, ctx->dv, ctx->frz, ctx->t, ctx->efcheck));
  unsigned long long h = 0;
  CUDA_OK(cudaMemcpyAsync(&h, ctx->frz.count, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  ctx->number_part_fo += (int64_t)h;
  if (count) *count = (int64_t)h;
  return CCB_OK;
}
// SPHWorkstation::freeze_out_particles (src/sph_workstation.cpp:1475-1485) = check_freeze_out_status + FreezeOut::bsqsvfreezeout (:550-641)
int do_freeze_out(ccb_ctx* ctx, int64_t* n_freezing_out) {
  if (!ctx->frz.frz1) { ctx->err = "ccb_freeze_out: the context was created with particlization_enabled = 0"; return CCB_ERR_STATE; }
  int64_t curfrz = 0;
  int rc = do_freeze_check_status(ctx, &curfrz);
  if (rc) return rc;
  if (n_freezing_out) *n_freezing_out = curfrz;
  if (ctx->frzc == 0) {          // first time step: current -> frz2
    ctx->taupp = ctx->t; ctx->frzc = 1;
    DISPATCH_DM("freeze_snapshot_kernel", launch_freeze_snapshot, ctx->stream, ctx->dv, ctx->frz, 2);
  } else if (ctx->frzc == 1) {   // second time step: current -> frz1
    ctx->taup = ctx->t; ctx->frzc = 2;
    DISPATCH_DM("freeze_snapshot_kernel", launch_freeze_snapshot, ctx->stream, ctx->dv, ctx->frz, 1);
    if (curfrz > 0) DISPATCH_DM("freeze_interpolate_kernel", launch_freeze_interpolate, ctx->stream, ctx->dv, ctx->frz, ctx->efcheck, ctx->taup, ctx->taupp, ctx->cfg.dt);
  } else {
    DISPATCH_DM("freeze_evo_kernel", launch_freeze_evo, ctx->stream, ctx->dv, ctx->frz);
    const double tau = ctx->t;
    if (curfrz > 0) DISPATCH_DM("freeze_interpolate_kernel", launch_freeze_interpolate, ctx->stream, ctx->dv, ctx->frz, ctx->efcheck, ctx->taup, ctx->taupp, ctx->cfg.dt);
    DISPATCH_DM("freeze_advance_kernel", launch_freeze_advance, ctx->stream, ctx->dv, ctx->frz);
    ctx->taupp = ctx->taup; ctx->taup = tau;
  }
  return CCB_OK;
}

int surface_device_error(ccb_ctx* ctx) {
  int32_t e[2] = {0, 0};
  CUDA_OK(cudaMemcpyAsync(e, ctx->d_small + 1, sizeof(e), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  if (e[0] != 0) {
    const char* what = e[0] == CCB_ERR_NEG_ENTROPY ? "negative entropy density" : e[0] == CCB_ERR_EOS_NO_ROOT ? "EoS: no solution found"
                       : e[0] == CCB_ERR_OUT_OF_GRID ? "particle left the neighbour grid" : "device error";
    ctx->err = std::string(what) + " (first offending particle id " + std::to_string(e[1]) + ")";
    CUDA_OK(cudaMemsetAsync(ctx->d_small + 1, 0, 2 * sizeof(int32_t), ctx->stream));
    return e[0];
  }
  return CCB_OK;
}

int do_init_particles(ccb_ctx* ctx, const StageParams& sp) {
  int eos_done = 0;
  if (!ctx->cfg.online_inverter_enabled) {
    // SPHWorkstation::initialize_entropy_and_charge_densities with use_preinverted (src/sph_workstation.cpp:260-320): flat table only
    if (!ctx->have_pre || !ctx->pre_table.flat) {
      ctx->err = "online_inverter_enabled = 0: the initialisation needs the flat (s, nB) pre-inverted table (ccb_set_preinverted_flat), as in the reference";
      return CCB_ERR_STATE;
    }
    LAUNCHER("pre_init_kernel", launch_pre_init(ctx->stream, ctx->dv, ctx->pre_flat_dev, ctx->init_tmp, ctx->cfg.input_as_entropy));
    eos_done = 1;
  } else
  if (ctx->use_chain) {
    if (!ctx->have_table) { ctx->err = "eos_type = table: no table loaded (ccb_set_eos_table / ccb_load_eos_table_dat)"; return CCB_ERR_STATE; }
    LAUNCHER("eos_chain_init_kernel", launch_eos_chain_init(ctx->stream, ctx->dv, ctx->chain, ctx->init_tmp, ctx->cfg.input_as_entropy));
    eos_done = 1;
  }
  DISPATCH_DM("init_particles_kernel", launch_init_particles, ctx->stream, ctx->dv, sp, ctx->init_tmp, ctx->cfg.input_as_entropy,
              (ctx->cfg.input_initial_shear ? 1 : 0) | (ctx->init_diffusion ? 2 : 0), ctx->cfg.bulk_from_trace, ctx->efcheck, eos_done);
  return CCB_OK;
}

// host table -> device copies + fall-back chain + efcheck
int install_table(ccb_ctx* ctx, bool normalize_by_T) {
  HostTable& t = *ctx->table;
  if (!prepare_table(t, normalize_by_T, ctx->err)) return CCB_ERR_INVALID;
  std::memset(&ctx->chain, 0, sizeof(ctx->chain));
  if (!match_fallbacks(t, ctx->cfg.freeze_out_temperature_MeV, ctx->chain, ctx->err)) return CCB_ERR_INVALID;
  ctx->chain.prohibit_unstable_cs2 = ctx->cfg.prohibit_unstable_cs2;
  ctx->chain.prohibit_acausal_cs2 = ctx->cfg.prohibit_acausal_cs2;
  ctx->chain.restrict_mu_T_ratios = ctx->cfg.restrict_mu_T_ratios;
  const long long nodes = t.nodes();
  std::vector<double> hot((size_t)nodes * 4);
  cudaFree(ctx->d_hot_s); cudaFree(ctx->d_hot_e); cudaFree(ctx->d_full);
  ctx->d_hot_s = ctx->d_hot_e = ctx->d_full = nullptr;
  CUDA_OK(cudaMalloc(&ctx->d_hot_s, sizeof(double) * 4 * nodes));
  CUDA_OK(cudaMalloc(&ctx->d_hot_e, sizeof(double) * 4 * nodes));
  CUDA_OK(cudaMalloc(&ctx->d_full, sizeof(double) * TH_N * nodes));
  for (int pass = 0; pass < 2; ++pass) {
    for (long long i = 0; i < nodes; ++i) {
      const double* f = &t.fields[(size_t)i * TH_N];
      hot[(size_t)i * 4 + 0] = pass ? f[TH_E] : f[TH_S];
      hot[(size_t)i * 4 + 1] = f[TH_B]; hot[(size_t)i * 4 + 2] = f[TH_SS]; hot[(size_t)i * 4 + 3] = f[TH_Q];
    }
    CUDA_OK(cudaMemcpy(pass ? ctx->d_hot_e : ctx->d_hot_s, hot.data(), sizeof(double) * 4 * nodes, cudaMemcpyHostToDevice));
  }
  CUDA_OK(cudaMemcpy(ctx->d_full, t.fields.data(), sizeof(double) * TH_N * nodes, cudaMemcpyHostToDevice));
  // efcheck = e(T_FO, 0, 0, 0) of the default EoS: SPHWorkstation::initialize, src/sph_workstation.cpp:66-81
  {
    const TableView hv = host_view(t);
    const double x[4] = {ctx->cfg.freeze_out_temperature_MeV / kHbarcMeVfm, 0.0, 0.0, 0.0};
    double r[TH_N];
    if (!table_full_thermo(hv, x, r)) { ctx->err = "EoS table does not cover the freeze-out temperature"; return CCB_ERR_INVALID; }
    ctx->efcheck = r[TH_E];
  }
  TableView& v = ctx->chain.table;
  for (int d = 0; d < 4; ++d) { v.n[d] = t.n[d]; v.min[d] = t.min[d]; v.max[d] = t.max[d]; v.delta[d] = t.delta[d]; }
  v.hot_s = ctx->d_hot_s; v.hot_e = ctx->d_hot_e; v.full = ctx->d_full;
  ctx->have_table = true;
  // the host copy was only needed for the matching
  t.grid.clear(); t.grid.shrink_to_fit(); t.fields.clear(); t.fields.shrink_to_fit();
  return CCB_OK;
}

}  // namespace

namespace {
// scratch device buffer of a host-query call: freed on every exit path
template <class T> struct DevBuf {
  T* p = nullptr;
  ~DevBuf() { cudaFree(p); }
  cudaError_t alloc(int64_t count) { return cudaMalloc(&p, sizeof(T) * (size_t)count); }
};
}  // namespace

// =================================================================================
extern "C" {

int ccb_field_width(int field) { return (field >= 0 && field < CCB_F_COUNT) ? kFieldWidth[field] : -1; }

int ccb_create(ccb_ctx** out, const ccb_config* cfg) {
  if (!out || !cfg) return CCB_ERR_INVALID;
  *out = nullptr;
  ccb_ctx* ctx = new ccb_ctx();
  ctx->cfg = *cfg;
  *out = ctx;  // handed back even on failure so that ccb_last_error works
  if (cfg->dim != 2 && cfg->dim != 3) { ctx->err = "dim must be 2 or 3"; return CCB_ERR_INVALID; }
  if (cfg->rk_order != 2 && cfg->rk_order != 4) { ctx->err = "rk_order must be 2 or 4 (Evolver::execute_timestep, src/evolver.cpp:98-116)"; return CCB_ERR_INVALID; }
  if (cfg->eos_type < CCB_EOS_CONFORMAL || cfg->eos_type > CCB_EOS_CONFORMAL_DIAGONAL) { ctx->err = "eos_type must be one of CCB_EOS_*"; return CCB_ERR_INVALID; }
  ctx->use_chain = cfg->eos_type != CCB_EOS_CONFORMAL;
  if (cfg->eos_type == CCB_EOS_TANH_CONFORMAL || cfg->eos_type == CCB_EOS_CONFORMAL_DIAGONAL) {
    // EquationOfState::set_up_chosen_EOSs, src/eos_initialization.cpp:249-266: a chain of ONE analytic member
    std::memset(&ctx->chain, 0, sizeof(ctx->chain));
    EosChain& ch = ctx->chain;
    ch.n = 1;
    const AnalyticParams P{cfg->conformal_c, cfg->conformal_T0, cfg->conformal_muB0, cfg->conformal_muS0, cfg->conformal_muQ0,
                           220.0 / kHbarcMeVfm, 120.0 / kHbarcMeVfm};
    double r[TH_N];
    const double T_FO = cfg->freeze_out_temperature_MeV / kHbarcMeVfm;
    if (cfg->eos_type == CCB_EOS_TANH_CONFORMAL) { ch.kind[0] = EOS_KIND_TANH; ch.tanh_p = P; tanh_full_thermo(P, T_FO, 0.0, 0.0, 0.0, r); }
    else { ch.kind[0] = EOS_KIND_DIAGONAL; ch.diag_p = P; ch.diag_p.Tc = ch.diag_p.Ts = 0.0; diagonal_full_thermo(ch.diag_p, T_FO, 0.0, 0.0, 0.0, r); }
    ch.prohibit_unstable_cs2 = cfg->prohibit_unstable_cs2; ch.prohibit_acausal_cs2 = cfg->prohibit_acausal_cs2;
    ch.restrict_mu_T_ratios = cfg->restrict_mu_T_ratios;
    ctx->efcheck = r[TH_E];  // SPHWorkstation::initialize, src/sph_workstation.cpp:66-81
    ctx->have_table = true;
  }
  if (!(cfg->hT > 0.0)) { ctx->err = "hT must be positive"; return CCB_ERR_INVALID; }
  if (cfg->slab_axis < 0 || cfg->slab_axis >= cfg->dim) { ctx->err = "slab_axis must be a spatial axis of the run (0 .. dim - 1)"; return CCB_ERR_INVALID; }
  ctx->D = cfg->dim;
  ctx->milne = cfg->coordinate_system == CCB_COORD_HYPERBOLIC;
  ctx->t = cfg->t0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { ctx->err = "no CUDA device available (this library has no CPU path)"; return CCB_ERR_CUDA; }
  CUDA_OK(cudaSetDevice(cfg->device));
  CUDA_OK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CUDA_OK(cudaMalloc(&ctx->d_small, sizeof(int32_t) * 8));
  CUDA_OK(cudaMemsetAsync(ctx->d_small, 0, sizeof(int32_t) * 8, ctx->stream));
  CUDA_OK(cudaMalloc(&ctx->d_reduce, sizeof(double) * 7 * 296));
  CUDA_OK(cudaMalloc(&ctx->d_taskinfo, sizeof(int32_t) * 16));
  CUDA_OK(cudaMemsetAsync(ctx->d_taskinfo, 0, sizeof(int32_t) * 16, ctx->stream));
  {
    CUDA_OK(cudaMalloc(&ctx->d_qflag, sizeof(int) * 4));
    const int one_zero[4] = {1, 0, 1, 0};  // until the first local update has looked: assume non-zero currents
    CUDA_OK(cudaMemcpy(ctx->d_qflag, one_zero, sizeof(one_zero), cudaMemcpyHostToDevice));
    ctx->dv.qflag = ctx->d_qflag;
  }

  for (int p = 0; p < PH_COUNT; ++p) { ctx->phase_ms[p] = 0.; ctx->phase_calls[p] = 0; }
  setup_grid(ctx);
  ctx->slab_lo = 0; ctx->slab_hi = ctx->grid.nc[0];
  ctx->dv.ghost_col[0] = ctx->dv.ghost_col[1] = -1;
  ctx->dv.ghost_start[0] = ctx->dv.ghost_start[1] = nullptr;
  cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, cfg->device);
  {  // gather launch shape: 64 targets per task, 64 source records per shared-memory tile (measured best over the BASELINE
     // configs, profiles/README.md); environment overrides are for tuning runs only
    if (const char* e = std::getenv("CCB_GATHER_NT")) { const int v = std::atoi(e); if (v == 64 || v == 128) ctx->shape_grad.nt = ctx->shape_fields.nt = v; }
    if (const char* e = std::getenv("CCB_GATHER_TILE")) { const int v = std::atoi(e); if (v == 32 || v == 64) ctx->shape_grad.tile = v; }
    if (const char* e = std::getenv("CCB_FIELDS_TILE")) { const int v = std::atoi(e); if (v == 32 || v == 64) ctx->shape_fields.tile = v; }
    if (const char* e = std::getenv("CCB_HITMASK")) ctx->hit_off = std::atoi(e) == 0;
  }
  if (ctx->grid.ncells >= ((int64_t)1 << 31)) { ctx->err = "neighbour grid has more than 2^31 cells"; return CCB_ERR_INVALID; }
  ctx->kD = make_kernel_norm(ctx->D, cfg->hT);
  ctx->k2 = make_kernel_norm(2, cfg->hT);
  ctx->k1 = make_kernel_norm(1, cfg->hEta > 0 ? cfg->hEta : cfg->hT);
  // efcheck = e(T_FO, mu = 0): SPHWorkstation::initialize, src/sph_workstation.cpp:66-81
  if (cfg->eos_type == CCB_EOS_CONFORMAL || cfg->eos_type == CCB_EOS_TABLE) {  // (the table's own value replaces it on load)
    const double T = cfg->freeze_out_temperature_MeV / kHbarcMeVfm;
    ConformalParams cp{cfg->conformal_c, cfg->conformal_T0, cfg->conformal_muB0, cfg->conformal_muS0, cfg->conformal_muQ0};
    double r[TH_N];
    conformal_full_thermo(cp, T, 0., 0., 0., r);
    ctx->efcheck = r[TH_E];
  }
  return CCB_OK;
}

int ccb_destroy(ccb_ctx* ctx) {
  if (!ctx) return CCB_ERR_INVALID;
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  free_particles(ctx);
  cudaFree(ctx->cell_count); cudaFree(ctx->scan_sums); cudaFree(ctx->d_small); cudaFree(ctx->d_reduce); cudaFree(ctx->d_taskinfo);
  cudaFree(ctx->ghost_tab); cudaFree(ctx->d_info); cudaFree(ctx->d_ends); cudaFree(ctx->d_qflag); cudaFree(ctx->d_flags);
  if (ctx->comm_stream) { cudaStreamDestroy(ctx->comm_stream); cudaEventDestroy(ctx->ev_pack); cudaEventDestroy(ctx->ev_halo); }
  cudaFree(ctx->d_hot_s); cudaFree(ctx->d_hot_e); cudaFree(ctx->d_full);
  for (double* q : ctx->pre_alloc) cudaFree(q);
  cudaFree(ctx->d_pre_tiles);
  ctx->comm.reset();
  for (auto& e : ctx->ev_pool) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return CCB_OK;
}

const char* ccb_last_error(ccb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int ccb_set_eos_table(ccb_ctx* ctx, const ccb_eos_table* tb) {
  if (!ctx || !tb || !tb->grid || !tb->fields) return CCB_ERR_INVALID;
  if (ctx->cfg.eos_type != CCB_EOS_TABLE) { ctx->err = "ccb_set_eos_table: the context was created with another eos_type"; return CCB_ERR_STATE; }
  if (!ctx->stream) { ctx->err = "no CUDA device available (this library has no CPU path)"; return CCB_ERR_CUDA; }
  ctx->table.reset(new HostTable());
  long long nodes = 1;
  for (int d = 0; d < 4; ++d) { if (tb->n[d] < 1) { ctx->err = "table grid sizes must be positive"; return CCB_ERR_INVALID; } ctx->table->n[d] = tb->n[d]; nodes *= tb->n[d]; }
  ctx->table->grid.assign(tb->grid, tb->grid + nodes * 4);
  ctx->table->fields.assign(tb->fields, tb->fields + nodes * TH_N);
  return install_table(ctx, tb->normalize_by_T != 0);
}
int ccb_load_eos_table_dat(ccb_ctx* ctx, const char* path, int normalize_by_T) {
  if (!ctx || !path) return CCB_ERR_INVALID;
  if (ctx->cfg.eos_type != CCB_EOS_TABLE) { ctx->err = "ccb_load_eos_table_dat: the context was created with another eos_type"; return CCB_ERR_STATE; }
  if (!ctx->stream) { ctx->err = "no CUDA device available (this library has no CPU path)"; return CCB_ERR_CUDA; }
  ctx->table.reset(new HostTable());
  if (!read_thermo_dat(path, *ctx->table, ctx->err)) return CCB_ERR_INVALID;
  return install_table(ctx, normalize_by_T != 0);
}
// ---- pre-inverted equation of state ------------------------------------------------------------------------------------
static int pre_reset(ccb_ctx* ctx) {
  for (double* q : ctx->pre_alloc) cudaFree(q);
  ctx->pre_alloc.clear();
  ctx->have_pre = false;
  if (!ctx->d_pre_tiles) CUDA_OK(cudaMalloc(&ctx->d_pre_tiles, sizeof(PreTile) * kPreTileCount));
  return CCB_OK;
}
static int pre_upload(ccb_ctx* ctx, const double* host, size_t count, double** dev) {
  CUDA_OK(cudaMalloc(dev, sizeof(double) * count));
  ctx->pre_alloc.push_back(*dev);
  CUDA_OK(cudaMemcpy(*dev, host, sizeof(double) * count, cudaMemcpyHostToDevice));
  return CCB_OK;
}
int ccb_set_preinverted_table(ccb_ctx* ctx, int32_t ntiles, const ccb_preinverted_tile* tiles) {
  if (!ctx || ntiles < 1 || !tiles) return CCB_ERR_INVALID;
  if (!ctx->stream) { ctx->err = "no CUDA device available (this library has no CPU path)"; return CCB_ERR_CUDA; }
  int rc = pre_reset(ctx);
  if (rc) return rc;
  std::vector<PreTile> h(kPreTileCount);
  for (auto& t : h) std::memset(&t, 0, sizeof(t));
  for (int k = 0; k < ntiles; ++k) {
    const ccb_preinverted_tile& in = tiles[k];
    const int lim[4] = {5, 4, 4, 4};
    size_t count = PRE_NVAR;
    for (int i = 0; i < 4; ++i) {
      if (in.slot[i] < 0 || in.slot[i] >= lim[i] || in.N[i] < 1) { ctx->err = "ccb_set_preinverted_table: tile slot or axis size out of range"; return CCB_ERR_INVALID; }
      count *= (size_t)in.N[i];
    }
    if (!in.vars) return CCB_ERR_INVALID;
    PreTile& t = h[((in.slot[0] * 4 + in.slot[1]) * 4 + in.slot[2]) * 4 + in.slot[3]];
    for (int i = 0; i < 4; ++i) { t.min[i] = in.min[i]; t.step[i] = in.step[i]; t.N[i] = in.N[i]; t.log_spacing[i] = in.log_spacing[i]; }
    double* d = nullptr;
    if ((rc = pre_upload(ctx, in.vars, count, &d))) return rc;
    t.vars = d;
  }
  CUDA_OK(cudaMemcpy(ctx->d_pre_tiles, h.data(), sizeof(PreTile) * kPreTileCount, cudaMemcpyHostToDevice));
  ctx->pre_table.tiles = ctx->d_pre_tiles;
  ctx->pre_table.flat = 0;
  ctx->have_pre = true;
  return CCB_OK;
}
// EoS_Interpolator::load_flat_table (src/eos_interpolator.cpp:262-521) from arrays: unit conversion, log-uniformity check of the
// axes, dw/ds and dw/dB from the chi tensor (baryon-only Schur complement), host copies for the initialisation queries
int ccb_set_preinverted_flat(ccb_ctx* ctx, const ccb_preinverted_flat* in) {
  if (!ctx || !in || in->Ns < 2 || in->NB < 2) return CCB_ERR_INVALID;
  const double* req[] = {in->s, in->nB, in->T, in->muB, in->muS, in->muQ, in->e, in->p, in->cs2, in->chiTT, in->chiTB, in->chiBB};
  for (const double* q : req) if (!q) return CCB_ERR_INVALID;
  if (!ctx->stream) { ctx->err = "no CUDA device available (this library has no CPU path)"; return CCB_ERR_CUDA; }
  int rc = pre_reset(ctx);
  if (rc) return rc;
  const int Ns = in->Ns, NB = in->NB;
  const size_t Ntot = (size_t)Ns * NB;
  const double hc1 = 1.0 / kHbarcMeVfm, hc2 = hc1 * hc1, hc3 = hc2 * hc1, hc4 = hc2 * hc2;
  std::vector<std::vector<double>>& H = ctx->pre_flat_host;
  H.assign(11, std::vector<double>());
  enum { HS = 0, HNB, HT, HMUB, HMUS, HMUQ, HE, HP, HCS2, HDWDS, HDWDB };
  H[HS].resize(Ns); H[HNB].resize(NB);
  for (int is = 0; is < Ns; ++is) H[HS][is] = in->s[(size_t)is * NB] * hc3;
  for (int iB = 0; iB < NB; ++iB) H[HNB][iB] = in->nB[iB] * hc3;
  double lmin[2], lstep[2];
  for (int a = 0; a < 2; ++a) {
    const std::vector<double>& ax = H[a];
    const double l0 = std::log10(ax.front()), l1 = std::log10(ax.back()), step = (l1 - l0) / (ax.size() - 1);
    double dev = 0.0;
    for (size_t i = 1; i + 1 < ax.size(); ++i) dev = std::max(dev, std::fabs(std::log10(ax[i + 1]) - std::log10(ax[i]) - step));
    if (!(dev <= 1e-6)) { ctx->err = std::string("ccb_set_preinverted_flat: axis ") + (a ? "nB" : "s") + " is not uniformly log-spaced"; return CCB_ERR_INVALID; }
    lmin[a] = l0; lstep[a] = step;
  }
  for (int k = HT; k <= HDWDB; ++k) H[k].resize(Ntot);
  const double EPS = 1e-30;
  for (int is = 0; is < Ns; ++is)
    for (int iB = 0; iB < NB; ++iB) {
      const size_t i = (size_t)is * NB + iB;
      const double T = in->T[i] * hc1, muB = in->muB[i] * hc1;
      const double chiTT = in->chiTT[i] * hc2, chiTB = in->chiTB[i] * hc2, chiBB = in->chiBB[i] * hc2;
      const double s_c = H[HS][is], nB_c = H[HNB][iB];
      const double dentr_dt = chiTT - (chiTB * chiTB) / (chiBB + EPS), dentr_dmub = chiTB - (chiTT * chiBB) / (chiTB + EPS);
      const double db_dt = chiTB - (chiBB * chiTT) / (chiTB + EPS), db_dmub = chiBB - (chiTB * chiTB) / (chiTT + EPS);
      double dwds = T + s_c / dentr_dt;
      if (std::fabs(nB_c) > kTiny) dwds += nB_c / dentr_dmub;
      double dwdB = muB;
      if (std::fabs(nB_c) > kTiny) dwdB += s_c / db_dt + nB_c / db_dmub;
      H[HT][i] = T; H[HMUB][i] = muB; H[HMUS][i] = in->muS[i] * hc1; H[HMUQ][i] = in->muQ[i] * hc1;
      H[HE][i] = in->e[i] * hc4; H[HP][i] = in->p[i] * hc4; H[HCS2][i] = in->cs2[i];
      H[HDWDS][i] = dwds; H[HDWDB][i] = dwdB;
    }
  // device: the tile in slot [0][0][0][0] (11 fields, eos_variables order; dwdS = dwdQ = 0) and the flat arrays of the init queries
  std::vector<double> vars((size_t)PRE_NVAR * Ntot, 0.0);
  const int src[PRE_NVAR] = {HT, HMUB, HMUQ, HMUS, HE, HP, HCS2, HDWDS, HDWDB, -1, -1};
  for (int v = 0; v < PRE_NVAR; ++v) if (src[v] >= 0) std::copy(H[src[v]].begin(), H[src[v]].end(), vars.begin() + (size_t)v * Ntot);
  std::vector<PreTile> h(kPreTileCount);
  for (auto& t : h) std::memset(&t, 0, sizeof(t));
  PreTile& t = h[0];
  t.min[0] = lmin[0]; t.step[0] = lstep[0]; t.N[0] = Ns; t.log_spacing[0] = 1;
  t.min[1] = lmin[1]; t.step[1] = lstep[1]; t.N[1] = NB; t.log_spacing[1] = 1;
  t.N[2] = t.N[3] = 1;
  double* d = nullptr;
  if ((rc = pre_upload(ctx, vars.data(), vars.size(), &d))) return rc;
  t.vars = d;
  CUDA_OK(cudaMemcpy(ctx->d_pre_tiles, h.data(), sizeof(PreTile) * kPreTileCount, cudaMemcpyHostToDevice));
  ctx->pre_table.tiles = ctx->d_pre_tiles;
  ctx->pre_table.flat = 1;
  PreFlat& f = ctx->pre_flat_dev;
  f.Ns = Ns; f.NB = NB; f.log_s_min = lmin[0]; f.dlog_s = lstep[0]; f.log_nB_min = lmin[1]; f.dlog_nB = lstep[1];
  const double** dst[11] = {&f.s_axis, &f.nB_axis, &f.T, &f.muB, &f.muS, &f.muQ, &f.e, &f.p, &f.cs2, &f.dwds, &f.dwdB};
  for (int k = 0; k < 11; ++k) {
    double* q = nullptr;
    if ((rc = pre_upload(ctx, H[k].data(), H[k].size(), &q))) return rc;
    *dst[k] = q;
  }
  // efcheck = EoS_Interpolator::efreeze_at_T(T_FO) (SPHWorkstation::initialize, src/sph_workstation.cpp:75-79), on the host copy
  PreFlat hf = f;
  hf.s_axis = H[HS].data(); hf.nB_axis = H[HNB].data(); hf.T = H[HT].data(); hf.muB = H[HMUB].data(); hf.muS = H[HMUS].data();
  hf.muQ = H[HMUQ].data(); hf.e = H[HE].data(); hf.p = H[HP].data(); hf.cs2 = H[HCS2].data(); hf.dwds = H[HDWDS].data(); hf.dwdB = H[HDWDB].data();
  ctx->efcheck = pre_efreeze_at_T(hf, ctx->cfg.freeze_out_temperature_MeV / kHbarcMeVfm);
  ctx->have_pre = true;
  return CCB_OK;
}
// utilities/generate_eos_table.cpp on the device: one decade tile located with the run's own inversion kernel
int ccb_generate_preinverted_tile(ccb_ctx* ctx, const int32_t exponent[4], int32_t N, double* vars, int32_t* eos) {
  if (!ctx || !exponent || N < 2 || N > 64 || !vars || !eos) return CCB_ERR_INVALID;
  if (!ctx->stream) { ctx->err = "no CUDA device available (this library has no CPU path)"; return CCB_ERR_CUDA; }
  if (ctx->use_chain && !ctx->have_table) { ctx->err = "eos_type = table: no table loaded (ccb_set_eos_table / ccb_load_eos_table_dat)"; return CCB_ERR_STATE; }
  const int64_t total = (int64_t)N * N * N * N;
  DevBuf<double> d_vars;
  DevBuf<int32_t> d_eos;
  CUDA_OK(d_vars.alloc(PRE_NVAR * total));
  CUDA_OK(d_eos.alloc(total));
  const ConformalParams cp{ctx->cfg.conformal_c, ctx->cfg.conformal_T0, ctx->cfg.conformal_muB0, ctx->cfg.conformal_muS0, ctx->cfg.conformal_muQ0};
  LAUNCHER("pre_generate_kernel", launch_pre_generate(ctx->stream, ctx->chain, cp, ctx->use_chain ? 1 : 0, exponent, N, d_vars.p, d_eos.p));
  CUDA_OK(cudaMemcpyAsync(vars, d_vars.p, sizeof(double) * PRE_NVAR * total, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaMemcpyAsync(eos, d_eos.p, sizeof(int32_t) * total, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return CCB_OK;
}

int ccb_eos_full_thermo(ccb_ctx* ctx, int member, int64_t n, const double* points, double* out) {
  if (!ctx || n < 0 || (n && (!points || !out))) return CCB_ERR_INVALID;
  if (n == 0) return CCB_OK;
  if (ctx->use_chain && !ctx->have_table) { ctx->err = "eos_type = table: no table loaded (ccb_set_eos_table / ccb_load_eos_table_dat)"; return CCB_ERR_STATE; }
  if (member < 0 || member >= (ctx->use_chain ? ctx->chain.n : 1)) { ctx->err = "ccb_eos_full_thermo: no such member of chosen_EOSs"; return CCB_ERR_INVALID; }
  DevBuf<double> d_in, d_out;
  CUDA_OK(d_in.alloc(4 * n));
  CUDA_OK(d_out.alloc(TH_N * n));
  CUDA_OK(cudaMemcpyAsync(d_in.p, points, sizeof(double) * 4 * n, cudaMemcpyHostToDevice, ctx->stream));
  const ConformalParams cp{ctx->cfg.conformal_c, ctx->cfg.conformal_T0, ctx->cfg.conformal_muB0, ctx->cfg.conformal_muS0, ctx->cfg.conformal_muQ0};
  LAUNCHER("eos_query_kernel", launch_eos_query(ctx->stream, ctx->chain, ctx->use_chain ? nullptr : &cp, member, n, d_in.p, d_out.p));
  CUDA_OK(cudaMemcpyAsync(out, d_out.p, sizeof(double) * TH_N * n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return CCB_OK;
}

int ccb_eos_locate(ccb_ctx* ctx, int energy, int64_t n, const double* dens, const double* seed, const int32_t* start, double* thermo, int32_t* eos) {
  if (!ctx || n < 0 || (n && (!dens || !thermo || !eos))) return CCB_ERR_INVALID;
  if (n == 0) return CCB_OK;
  if (!energy && !seed && ctx->use_chain) { ctx->err = "ccb_eos_locate: update_s needs a seed point"; return CCB_ERR_INVALID; }
  if (ctx->use_chain && !ctx->have_table) { ctx->err = "eos_type = table: no table loaded (ccb_set_eos_table / ccb_load_eos_table_dat)"; return CCB_ERR_STATE; }
  DevBuf<double> d_dens, d_seed, d_th;
  DevBuf<int32_t> d_start, d_eos;
  CUDA_OK(d_dens.alloc(4 * n));
  CUDA_OK(d_th.alloc(16 * n));
  CUDA_OK(d_eos.alloc(n));
  CUDA_OK(cudaMemcpyAsync(d_dens.p, dens, sizeof(double) * 4 * n, cudaMemcpyHostToDevice, ctx->stream));
  if (seed) { CUDA_OK(d_seed.alloc(4 * n)); CUDA_OK(cudaMemcpyAsync(d_seed.p, seed, sizeof(double) * 4 * n, cudaMemcpyHostToDevice, ctx->stream)); }
  if (start) { CUDA_OK(d_start.alloc(n)); CUDA_OK(cudaMemcpyAsync(d_start.p, start, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream)); }
  const ConformalParams cp{ctx->cfg.conformal_c, ctx->cfg.conformal_T0, ctx->cfg.conformal_muB0, ctx->cfg.conformal_muS0, ctx->cfg.conformal_muQ0};
  LAUNCHER("eos_locate_kernel", launch_eos_locate(ctx->stream, ctx->chain, cp, ctx->use_chain ? 1 : 0, energy, n, d_dens.p, d_seed.p, d_start.p, d_th.p, d_eos.p));
  CUDA_OK(cudaMemcpyAsync(thermo, d_th.p, sizeof(double) * 16 * n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaMemcpyAsync(eos, d_eos.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return CCB_OK;
}

int ccb_eos_chain_parameters(ccb_ctx* ctx, double out[18]) {
  if (!ctx || !out) return CCB_ERR_INVALID;
  if (!ctx->have_table) { ctx->err = "no table loaded"; return CCB_ERR_STATE; }
  const AnalyticParams* P[3] = {&ctx->chain.tanh_p, &ctx->chain.conf_p, &ctx->chain.diag_p};
  int o = 0;
  for (int k = 0; k < 3; ++k) {
    out[o++] = P[k]->c; out[o++] = P[k]->T0; out[o++] = P[k]->muB0; out[o++] = P[k]->muS0; out[o++] = P[k]->muQ0;
    if (k == 0) { out[o++] = P[k]->Tc; out[o++] = P[k]->Ts; }
  }
  out[17] = ctx->efcheck;
  return CCB_OK;
}
int64_t ccb_n(ccb_ctx* ctx) { return ctx ? ctx->n : -1; }
double ccb_time(ccb_ctx* ctx) { return ctx ? ctx->t : 0.0; }
int ccb_set_time(ccb_ctx* ctx, double t) { if (!ctx) return CCB_ERR_INVALID; ctx->t = t; return CCB_OK; }
int64_t ccb_kernel_launches(ccb_ctx* ctx) { return ctx ? ctx->launches : -1; }
void* ccb_stream(ccb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int ccb_grid_shape(ccb_ctx* ctx, int32_t nc[3]) { if (!ctx) return CCB_ERR_INVALID; for (int k = 0; k < 3; ++k) nc[ctx->grid.axis[k]] = ctx->grid.nc[k]; return CCB_OK; }

int ccb_profile_enable(ccb_ctx* ctx, int on) {
  if (!ctx) return CCB_ERR_INVALID;
  resolve_phase_events(ctx);
  ctx->profiling = on != 0;
  for (int p = 0; p < PH_COUNT; ++p) { ctx->phase_ms[p] = 0.; ctx->phase_calls[p] = 0; }
  return CCB_OK;
}
int ccb_profile_get(ccb_ctx* ctx, int phase, double* ms, int64_t* calls) {
  if (!ctx || phase < 0 || phase >= PH_COUNT) return CCB_ERR_INVALID;
  resolve_phase_events(ctx);
  if (ms) *ms = ctx->phase_ms[phase];
  if (calls) *calls = ctx->phase_calls[phase];
  return CCB_OK;
}

int ccb_profile_phase_count(void) { return PH_COUNT; }
const char* ccb_profile_phase_name(int phase) {
  // the two slab entries are PARTS of 01_neighbour_list in a multi-GPU run (they include waiting for the neighbouring ranks)
  static const char* kName[PH_COUNT] = {"01_neighbour_list", "02_gamma_velocity", "03_smooth_fields", "04_thermodynamics", "05_06_08_local_update",
                                        "09_smooth_gradients", "10_time_derivatives", "other", "01a_slab_migration", "01b_slab_column_tables"};
  return phase >= 0 && phase < PH_COUNT ? kName[phase] : nullptr;
}

int ccb_allocate(ccb_ctx* ctx, int64_t n) {
  if (!ctx || n <= 0) return CCB_ERR_INVALID;
  int rc = allocate_particles(ctx, n);
  if (rc) return rc;
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return CCB_OK;
}

int ccb_set_field(ccb_ctx* ctx, int field, const double* host) {
  int rc = check_ready(ctx, false);
  if (rc) return rc;
  if (ctx->comm) { ctx->err = "ccb_set_field: particle-id ordered access is per GPU; with a slab decomposition use the *_device_order getters and CCB_I_ID"; return CCB_ERR_STATE; }
  if (field < 0 || field >= CCB_F_COUNT || !host) return CCB_ERR_INVALID;
  const int w = kFieldWidth[field];
  CUDA_OK(cudaMemcpyAsync(ctx->d_stage, host, sizeof(double) * (size_t)w * ctx->n, cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(gather_rows_by_id, cdiv(ctx->n, 256), 256, ctx->d_stage, w, 0, 1, w, ctx->dv.f[field], ctx->cap, ctx->dv.id, ctx->n);
  CUDA_OK(cudaStreamSynchronize(ctx->stream));   // the caller's buffer may be pageable and reused at once
  if (field == CCB_F_POSITION) ctx->cells_valid = false;
  return CCB_OK;
}

int ccb_get_field(ccb_ctx* ctx, int field, double* host) {
  int rc = check_ready(ctx, false);
  if (rc) return rc;
  if (ctx->comm) { ctx->err = "ccb_get_field: particle-id ordered access is per GPU; with a slab decomposition use the *_device_order getters and CCB_I_ID"; return CCB_ERR_STATE; }
  if (field < 0 || field >= CCB_F_COUNT || !host) return CCB_ERR_INVALID;
  const int w = kFieldWidth[field];
  LAUNCH(scatter_rows_by_id, cdiv(ctx->n, 256), 256, ctx->dv.f[field], ctx->cap, w, ctx->d_stage, ctx->dv.id, ctx->n);
  CUDA_OK(cudaMemcpyAsync(host, ctx->d_stage, sizeof(double) * (size_t)w * ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return CCB_OK;
}

int ccb_set_int_field(ccb_ctx* ctx, int field, const int32_t* host) {
  int rc = check_ready(ctx, false);
  if (rc) return rc;
  if (ctx->comm) { ctx->err = "ccb_set_int_field: particle-id ordered access is per GPU; with a slab decomposition use the *_device_order getters and CCB_I_ID"; return CCB_ERR_STATE; }
  if (!(field == CCB_I_BTRACK || field == CCB_I_FREEZE || field == CCB_I_EOS) || !host) return CCB_ERR_INVALID;
  int32_t* dst = field == CCB_I_BTRACK ? ctx->dv.btrack : field == CCB_I_FREEZE ? ctx->dv.freeze : ctx->dv.eos_id;
  int32_t* stage = reinterpret_cast<int32_t*>(ctx->d_stage);
  CUDA_OK(cudaMemcpyAsync(stage, host, sizeof(int32_t) * ctx->n, cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(gather_int_by_id, cdiv(ctx->n, 256), 256, stage, dst, ctx->dv.id, ctx->n);
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return CCB_OK;
}

int ccb_get_int_field(ccb_ctx* ctx, int field, int32_t* host) {
  int rc = check_ready(ctx, false);
  if (rc) return rc;
  if (ctx->comm) { ctx->err = "ccb_get_int_field: particle-id ordered access is per GPU; with a slab decomposition use the *_device_order getters and CCB_I_ID"; return CCB_ERR_STATE; }
  if (field < 0 || field >= CCB_I_COUNT || !host) return CCB_ERR_INVALID;
  const int32_t* src = field == CCB_I_ID ? ctx->dv.id : field == CCB_I_BTRACK ? ctx->dv.btrack : field == CCB_I_FREEZE ? ctx->dv.freeze
                       : field == CCB_I_EOS ? ctx->dv.eos_id : reinterpret_cast<const int32_t*>(ctx->dv.key);
  if (field == CCB_I_CELL && !ctx->cells_built) { ctx->err = "cell list not built"; return CCB_ERR_STATE; }
  int32_t* stage = reinterpret_cast<int32_t*>(ctx->d_stage);
  LAUNCH(scatter_int_by_id, cdiv(ctx->n, 256), 256, src, stage, ctx->dv.id, ctx->n);
  CUDA_OK(cudaMemcpyAsync(host, stage, sizeof(int32_t) * ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  if (field == CCB_I_CELL && ctx->grid.axis[0] != 0)   // the key's digits follow the slab axis: report the reference's numbering
    for (int64_t i = 0; i < ctx->n; ++i) host[i] = reference_cell_id(ctx->grid, (uint32_t)host[i]);
  return CCB_OK;
}

int ccb_device_field(ccb_ctx* ctx, int field, int component, double** dptr) {
  int rc = check_ready(ctx, false);
  if (rc) return rc;
  if (field < 0 || field >= CCB_F_COUNT || component < 0 || component >= kFieldWidth[field] || !dptr) return CCB_ERR_INVALID;
  *dptr = ctx->dv.f[field] + (int64_t)component * ctx->cap;
  return CCB_OK;
}

// Evolved state (the reference's EvolverCache member list, include/evolver.h:21-47) as one
// column-major block [column][particle id]: position 3, u 3, extensive 4, extensive_bulk 1,
// extensive_shv 5, (extensive_diffusion 9 when diffusion is on), Ez 1.
static int state_columns(ccb_ctx* ctx, ColumnList& cl) {
  cl.ncols = 0;
  for (int k = 0; k < kNumEvolvedFields; ++k) {
    const int f = kEvolvedFields[k];
    if (f == CCB_F_EXTENSIVE_DIFFUSION && !ctx->cfg.using_diffusion) continue;
    for (int c = 0; c < kFieldWidth[f]; ++c) cl.col[cl.ncols++] = ctx->dv.f[f] + (int64_t)c * ctx->cap;
  }
  return cl.ncols;
}
int ccb_state_width(ccb_ctx* ctx) {
  if (!ctx) return -1;
  int w = 0;
  for (int k = 0; k < kNumEvolvedFields; ++k) {
    const int f = kEvolvedFields[k];
    if (f == CCB_F_EXTENSIVE_DIFFUSION && !ctx->cfg.using_diffusion) continue;
    w += kFieldWidth[f];
  }
  return w;
}
int ccb_upload_state(ccb_ctx* ctx, const double* host) {
  int rc = check_ready(ctx, false);
  if (rc) return rc;
  if (!host) return CCB_ERR_INVALID;
  ColumnList cl;
  const int w = state_columns(ctx, cl);
  if (ctx->comm) {  // slab decomposition: device order, the layout ccb_download_state produced (column stride = ccb_n)
    for (int c = 0; c < w; ++c)
      CUDA_OK(cudaMemcpyAsync(cl.col[c], host + (size_t)c * ctx->n, sizeof(double) * ctx->n, cudaMemcpyHostToDevice, ctx->stream));
    ctx->cells_valid = false;
    return CCB_OK;
  }
  CUDA_OK(cudaMemcpyAsync(ctx->d_stage, host, sizeof(double) * (size_t)w * ctx->n, cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(gather_columns_by_id, dim3((unsigned)cdiv(ctx->n, 256), w), 256, ctx->d_stage, cl, ctx->dv.id, ctx->n);
  ctx->cells_valid = false;
  return CCB_OK;
}
int ccb_download_state(ccb_ctx* ctx, double* host) {
  int rc = check_ready(ctx, false);
  if (rc) return rc;
  if (!host) return CCB_ERR_INVALID;
  ColumnList cl;
  const int w = state_columns(ctx, cl);
  if (ctx->comm) {  // slab decomposition: this rank's particles in device order (ids: ccb_get_int_field_device_order)
    for (int c = 0; c < w; ++c)
      CUDA_OK(cudaMemcpyAsync(host + (size_t)c * ctx->n, cl.col[c], sizeof(double) * ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
    return surface_device_error(ctx);
  }
  LAUNCH(scatter_columns_by_id, dim3((unsigned)cdiv(ctx->n, 256), w), 256, cl, ctx->d_stage, ctx->dv.id, ctx->n);
  CUDA_OK(cudaMemcpyAsync(host, ctx->d_stage, sizeof(double) * (size_t)w * ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
  return surface_device_error(ctx);
}

// fp64 FMA peak of the device, for the roofline of the neighbour gathers (the driver measures HBM
// and bf16 only). 16 independent DFMA chains per thread, 8 resident blocks of 256 per SM.
int ccb_measure_dfma_peak(int device, double* tflops) {
  if (!tflops) return CCB_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return CCB_ERR_CUDA;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return CCB_ERR_CUDA;
  double* d = nullptr;
  if (cudaMalloc(&d, sizeof(double) * 256) != cudaSuccess) return CCB_ERR_CUDA;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 4096, blocks = sms * 8;
  dfma_peak_kernel<<<blocks, 256>>>(d, iters, 1.0000001);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a);
    dfma_peak_kernel<<<blocks, 256>>>(d, iters, 1.0000001);
    cudaEventRecord(b);
    if (cudaEventSynchronize(b) != cudaSuccess) { cudaFree(d); return CCB_ERR_CUDA; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    const double fl = 2.0 * 16.0 * (double)iters * 256.0 * blocks;
    const double tf = fl / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  cudaFree(d);
  *tflops = best;
  return CCB_OK;
}

// ---- slab decomposition --------------------------------------------------------------------
static int comm_buffers(ccb_ctx* ctx) {
  const int64_t T = (int64_t)ctx->grid.nc[1] * ctx->grid.nc[2] + 1;
  CUDA_OK(cudaMalloc(&ctx->ghost_tab, sizeof(uint32_t) * 4 * T));
  CUDA_OK(cudaMemset(ctx->ghost_tab, 0, sizeof(uint32_t) * 4 * T));
  CUDA_OK(cudaMalloc(&ctx->d_info, sizeof(int32_t) * 8 + sizeof(double) * 8));
  CUDA_OK(cudaMalloc(&ctx->d_ends, sizeof(uint32_t) * 12));
  CUDA_OK(cudaMalloc(&ctx->d_flags, sizeof(int) * 8));
  CUDA_OK(cudaMemset(ctx->d_flags, 0, sizeof(int) * 8));
  return CCB_OK;
}
int ccb_comm_unique_id(char id[128]) {
  if (!id) return CCB_ERR_INVALID;
  std::string err;
  NcclApi& api = NcclApi::get();
  if (!api.load(err)) return CCB_ERR_COMM;
  NcclApi::unique_id uid;
  if (api.GetUniqueId(&uid)) return CCB_ERR_COMM;
  std::memcpy(id, uid.internal, 128);
  return CCB_OK;
}
int ccb_comm_init(ccb_ctx* ctx, const char id[128]) {
  if (!ctx || !id) return CCB_ERR_INVALID;
  if (ctx->cfg.nranks < 2) { ctx->err = "ccb_comm_init: nranks < 2"; return CCB_ERR_INVALID; }
  if (ctx->have_state) { ctx->err = "ccb_comm_init must precede the particle upload"; return CCB_ERR_STATE; }
  CUDA_OK(cudaSetDevice(ctx->cfg.device));
  auto t = std::make_unique<NcclLink>();
  if (t->init(id, ctx->cfg.rank, ctx->cfg.nranks)) { ctx->err = "ccb_comm_init: " + t->err; return CCB_ERR_COMM; }
  ctx->comm = std::move(t);
  // one process per GPU: the boundary tasks read the neighbours' records over NVLink peer memory (CCB_HALO_P2P=0: halo messages)
  ctx->p2p = true;
  if (const char* e = std::getenv("CCB_HALO_P2P")) ctx->p2p = std::atoi(e) != 0;
  return comm_buffers(ctx);
}
int ccb_comm_init_local(ccb_ctx* ctx, int group) {
  if (!ctx) return CCB_ERR_INVALID;
  if (ctx->cfg.nranks < 2) { ctx->err = "ccb_comm_init_local: nranks < 2"; return CCB_ERR_INVALID; }
  if (ctx->have_state) { ctx->err = "ccb_comm_init_local must precede the particle upload"; return CCB_ERR_STATE; }
  ctx->comm = std::make_unique<LocalLink>(group, ctx->cfg.rank, ctx->cfg.nranks);
  // several contexts of ONE process on one GPU: halo messages by default. The flag-ordered peer path works here too
  // (CCB_HALO_P2P=1) but its spinning wait kernels need every stream on its own hardware queue (CUDA_DEVICE_MAX_CONNECTIONS).
  ctx->p2p = false;
  if (const char* e = std::getenv("CCB_HALO_P2P")) ctx->p2p = std::atoi(e) != 0;
  return comm_buffers(ctx);
}
int ccb_set_slab(ccb_ctx* ctx, int32_t col_lo, int32_t col_hi) {
  if (!ctx) return CCB_ERR_INVALID;
  const int nx = ctx->grid.nc[0];
  if (ctx->cfg.rank == 0) col_lo = 0;                          // the outermost slabs extend to the grid edge
  if (ctx->cfg.rank == ctx->cfg.nranks - 1) col_hi = nx;
  if (col_lo < 0 || col_hi > nx || col_lo >= col_hi) { ctx->err = "ccb_set_slab: empty or out-of-grid column range"; return CCB_ERR_INVALID; }
  ctx->slab_lo = col_lo; ctx->slab_hi = col_hi;
  ctx->cells_valid = false;
  return CCB_OK;
}
int ccb_grid_geometry(ccb_ctx* ctx, double min[3], double edge[3], int32_t nc[3]) {
  if (!ctx || !min || !edge || !nc) return CCB_ERR_INVALID;
  for (int k = 0; k < 3; ++k) { const int d = ctx->grid.axis[k]; min[d] = ctx->grid.min[k]; edge[d] = 1.0 / ctx->grid.rdx[k]; nc[d] = ctx->grid.nc[k]; }
  return CCB_OK;
}
int ccb_migration_counts(ccb_ctx* ctx, int64_t* out, int64_t* in) {
  if (!ctx) return CCB_ERR_INVALID;
  if (out) *out = ctx->migrated_out;
  if (in) *in = ctx->migrated_in;
  return CCB_OK;
}
int ccb_get_field_device_order(ccb_ctx* ctx, int field, double* host) {
  int rc = check_ready(ctx, false);
  if (rc) return rc;
  if (field < 0 || field >= CCB_F_COUNT || !host) return CCB_ERR_INVALID;
  const int w = kFieldWidth[field];
  LAUNCH(scatter_rows_by_id, cdiv(ctx->n, 256), 256, ctx->dv.f[field], ctx->cap, w, ctx->d_stage, (const int32_t*)nullptr, ctx->n);
  CUDA_OK(cudaMemcpyAsync(host, ctx->d_stage, sizeof(double) * (size_t)w * ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return CCB_OK;
}
int ccb_get_int_field_device_order(ccb_ctx* ctx, int field, int32_t* host) {
  int rc = check_ready(ctx, false);
  if (rc) return rc;
  if (field < 0 || field >= CCB_I_COUNT || !host) return CCB_ERR_INVALID;
  if (field == CCB_I_CELL && !ctx->cells_built) { ctx->err = "cell list not built"; return CCB_ERR_STATE; }
  const int32_t* src = field == CCB_I_ID ? ctx->dv.id : field == CCB_I_BTRACK ? ctx->dv.btrack : field == CCB_I_FREEZE ? ctx->dv.freeze
                       : field == CCB_I_EOS ? ctx->dv.eos_id : reinterpret_cast<const int32_t*>(ctx->dv.key);
  CUDA_OK(cudaMemcpyAsync(host, src, sizeof(int32_t) * ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  if (field == CCB_I_CELL && ctx->grid.axis[0] != 0)
    for (int64_t i = 0; i < ctx->n; ++i) host[i] = reference_cell_id(ctx->grid, (uint32_t)host[i]);
  return CCB_OK;
}

// Debug / test entry point: the neighbour SETS the field gather visits, as a CSR list in particle-id order with ascending
// neighbour ids -- what SystemState::reset_neighbour_list stores as its Verlet list (src/system_state.cpp:651-719). The product
// path never materialises this list (2 GB for config 1); a twin of the field gather (FieldsPolicy EMIT) writes the id of every
// neighbour it visits. Call after ccb_smooth_fields; offsets has ccb_n() + 1 entries, idx room for `cap` entries.
int ccb_debug_neighbour_list(ccb_ctx* ctx, int64_t* offsets, int32_t* idx, int64_t cap) {
  int rc = check_ready(ctx, true);
  if (rc) return rc;
  if (!offsets || !idx) return CCB_ERR_INVALID;
  if (ctx->comm) { ctx->err = "ccb_debug_neighbour_list: single-context runs only"; return CCB_ERR_STATE; }
  const int64_t n = ctx->n;
  std::vector<int32_t> bt(n), ids(n);
  CUDA_OK(cudaMemcpyAsync(bt.data(), ctx->dv.btrack, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaMemcpyAsync(ids.data(), ctx->dv.id, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  std::vector<int64_t> off(n + 1, 0);
  for (int64_t i = 0; i < n; ++i) {
    if (bt[i] < 0) { ctx->err = "ccb_debug_neighbour_list: btrack = -1 (sticky) hides a neighbour count"; return CCB_ERR_STATE; }
    off[i + 1] = off[i] + bt[i];
  }
  const int64_t total = off[n];
  if (total > cap) { ctx->err = "ccb_debug_neighbour_list: idx buffer too small"; return CCB_ERR_INVALID; }
  DevBuf<int64_t> d_off;
  DevBuf<int32_t> d_idx;
  CUDA_OK(d_off.alloc(n + 1));
  CUDA_OK(d_idx.alloc(std::max<int64_t>(total, 1)));
  CUDA_OK(cudaMemcpyAsync(d_off.p, off.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_OK(cudaMemsetAsync(d_idx.p, 0xff, sizeof(int32_t) * std::max<int64_t>(total, 1), ctx->stream));
  // the field records share their buffer with the gradient records: pack them again (phase 02 is idempotent)
  if ((rc = do_prologue(ctx))) return rc;
  GatherParams gp = make_gather(ctx);
  gp.nbr_off = d_off.p; gp.nbr_idx = d_idx.p;
  CUDA_OK(cudaMemsetAsync(ctx->d_taskinfo + 4, 0, 2 * sizeof(int32_t), ctx->stream));
  if (ctx->D == 2) LAUNCHER("gather_kernel<fields, emit>", launch_gather_fields<2>(false, ctx->shape_fields, ctx->stream, ctx->dv, ctx->grid, gp));
  else LAUNCHER("gather_kernel<fields, emit>", launch_gather_fields<3>(false, ctx->shape_fields, ctx->stream, ctx->dv, ctx->grid, gp));
  std::vector<int32_t> raw(std::max<int64_t>(total, 1));
  CUDA_OK(cudaMemcpyAsync(raw.data(), d_idx.p, sizeof(int32_t) * std::max<int64_t>(total, 1), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  std::vector<int64_t> where(n);  // particle id -> sorted index
  for (int64_t i = 0; i < n; ++i) where[ids[i]] = i;
  offsets[0] = 0;
  for (int64_t p = 0; p < n; ++p) offsets[p + 1] = offsets[p] + bt[where[p]];
  for (int64_t p = 0; p < n; ++p) {
    const int64_t i = where[p];
    int32_t* dst = idx + offsets[p];
    std::copy(raw.begin() + off[i], raw.begin() + off[i + 1], dst);
    std::sort(dst, dst + bt[i]);
  }
  return CCB_OK;
}

int ccb_efcheck(ccb_ctx* ctx, double* efcheck) { if (!ctx || !efcheck) return CCB_ERR_INVALID; *efcheck = ctx->efcheck; return CCB_OK; }
int ccb_freeze_out(ccb_ctx* ctx, int64_t* n_freezing_out) { int rc = check_ready(ctx, false); return rc ? rc : do_freeze_out(ctx, n_freezing_out); }
int ccb_freeze_out_results(ccb_ctx* ctx, double* results, int32_t* print) {
  int rc = check_ready(ctx, false);
  if (rc) return rc;
  if (!results || !print) return CCB_ERR_INVALID;
  if (!ctx->frz.results) { ctx->err = "ccb_freeze_out_results: the context was created with particlization_enabled = 0"; return CCB_ERR_STATE; }
  CUDA_OK(cudaMemcpyAsync(results, ctx->frz.results, sizeof(double) * FR_ROW * ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaMemcpyAsync(print, ctx->frz.print, sizeof(int32_t) * ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return CCB_OK;
}
int ccb_build_cells(ccb_ctx* ctx) { int rc = check_ready(ctx, false); return rc ? rc : do_build_cells(ctx); }
int ccb_smooth_fields(ccb_ctx* ctx) {
  int rc = check_ready(ctx, true);
  if (rc) return rc;
  if ((rc = do_prologue(ctx))) return rc;
  return do_gather_fields(ctx);
}
int ccb_update_thermodynamics(ccb_ctx* ctx) { int rc = check_ready(ctx, false); return rc ? rc : do_eos_local(ctx, 1, 0); }
int ccb_local_update(ccb_ctx* ctx) { int rc = check_ready(ctx, false); return rc ? rc : do_eos_local(ctx, 0, 1); }
int ccb_smooth_gradients(ccb_ctx* ctx) { int rc = check_ready(ctx, true); return rc ? rc : do_gather_gradients(ctx); }
int ccb_time_derivatives(ccb_ctx* ctx) { int rc = check_ready(ctx, false); return rc ? rc : do_eom(ctx, 0.0, 0); }
int ccb_get_time_derivatives(ccb_ctx* ctx) { int rc = check_ready(ctx, false); return rc ? rc : derivative_evaluation(ctx, 0.0, 0); }

int ccb_step(ccb_ctx* ctx, double dt) {
  int rc = check_ready(ctx, false);
  if (rc) return rc;
  ctx->y0_pending = true;   // the snapshot y0 = y(t_n) is written by the first sort of the step (sort_and_permute)
  const double t0 = ctx->t;
  if (ctx->cfg.rk_order == 4) {
    // Evolver::advance_timestep_rk4 (src/evolver.cpp:508-561), stage for stage AS THE REFERENCE RUNS IT: five derivative
    // evaluations, the first of which only moves the state to the mid-point; the other four enter k with the weights
    // 1/6, 2/6, 2/6, 2/6 (update_k is called with n = 3 twice, :548-554). Kept for drop-in parity; note the weights sum to 7/6.
    int wy0 = 0;
    for (int k = 0; k < kNumEvolvedFields; ++k) wy0 += kFieldWidth[kEvolvedFields[k]];
    CUDA_OK(cudaMemsetAsync(ctx->k_slab, 0, sizeof(double) * wy0 * ctx->cap, ctx->stream));
    CUDA_OK(cudaMemsetAsync(ctx->alt_k_slab, 0, sizeof(double) * wy0 * ctx->cap, ctx->stream));
    const double h[5] = {0.5 * dt, 0.5 * dt, 0.5 * dt, dt, 0.0};
    const double w[5] = {0.0, 1.0 / 6.0, 2.0 / 6.0, 2.0 / 6.0, 2.0 / 6.0};
    for (int stage = 0; stage < 5; ++stage) {
      if ((rc = derivative_evaluation(ctx, h[stage], stage < 4 ? 1 : 0))) return rc;
      if (stage < 4) ctx->t = t0 + h[stage];
      if (w[stage] != 0.0) {
        const RkCols c = rk_columns(ctx);
        LAUNCH(rk4_accumulate_kernel, cdiv(ctx->n, 256), 256, c, w[stage], ctx->n);
      }
    }
    const RkCols c = rk_columns(ctx);
    LAUNCH(rk4_final_kernel, cdiv(ctx->n, 256), 256, c, dt, ctx->n, ctx->dv.freeze, ctx->dv.id, ctx->dv.err);
    ctx->cells_valid = false;
    ctx->t = t0 + dt;
    if (ctx->cfg.particlization_enabled) return do_freeze_out(ctx, nullptr);
    return CCB_OK;
  }
  // Evolver::advance_timestep_rk2 (src/evolver.cpp:494-507): midpoint rule
  if ((rc = derivative_evaluation(ctx, 0.5 * dt, 1))) return rc;
  ctx->t = t0 + 0.5 * dt;
  if ((rc = derivative_evaluation(ctx, dt, 1))) return rc;
  ctx->t = t0 + dt;
  if (ctx->cfg.particlization_enabled) return do_freeze_out(ctx, nullptr);   // SPHWorkstation::advance_timestep, src/sph_workstation.cpp:2063
  return CCB_OK;
}

int ccb_synchronize(ccb_ctx* ctx) {
  if (!ctx) return CCB_ERR_INVALID;
  const int rc = surface_device_error(ctx);
  resolve_phase_events(ctx);
  return rc;
}

int ccb_reduce(ccb_ctx* ctx, ccb_totals* out) {
  int rc = check_ready(ctx, false);
  if (rc) return rc;
  if (!out) return CCB_ERR_INVALID;
  const int nb = kReduceBlocks;
  double* d_out = ctx->d_reduce;
  unsigned long long* d_fr = reinterpret_cast<unsigned long long*>(ctx->d_reduce + 6 * nb);
  if (ctx->milne) LAUNCHER("reduce_totals_kernel", launch_reduce_totals<true>(nb, ctx->stream, ctx->dv, ctx->t, d_out, d_fr));
  else LAUNCHER("reduce_totals_kernel", launch_reduce_totals<false>(nb, ctx->stream, ctx->dv, ctx->t, d_out, d_fr));
  std::vector<double> h(7 * nb);
  CUDA_OK(cudaMemcpyAsync(h.data(), d_out, sizeof(double) * 7 * nb, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  double acc[6] = {0., 0., 0., 0., 0., 0.};
  unsigned long long fr = 0;
  for (int b = 0; b < nb; ++b) {
    for (int k = 0; k < 6; ++k) acc[k] += h[b * 6 + k];
    unsigned long long f;
    std::memcpy(&f, &h[6 * nb + b], sizeof(f));
    fr += f;
  }
  if (ctx->comm) {  // totals over all slabs
    double* d = reinterpret_cast<double*>(ctx->d_info + 8);
    double hbuf[7] = {acc[0], acc[1], acc[2], acc[3], acc[4], acc[5], (double)fr};
    CUDA_OK(cudaMemcpyAsync(d, hbuf, sizeof(hbuf), cudaMemcpyHostToDevice, ctx->stream));
    COMM_OK(ctx->comm->allreduce_sum(d, 7, ctx->stream));
    CUDA_OK(cudaMemcpyAsync(hbuf, d, sizeof(hbuf), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 6; ++k) acc[k] = hbuf[k];
    fr = (unsigned long long)(hbuf[6] + 0.5);
  }
  out->S = acc[0]; out->B = acc[1]; out->Sq = acc[2]; out->Q = acc[3]; out->E = acc[4]; out->Ez = acc[5];
  out->n_frozen = (int64_t)fr;
  return CCB_OK;
}

int ccb_upload_ic(ccb_ctx* ctx, const ccb_ic* ic) {
  if (!ctx || !ic || ic->n <= 0 || !ic->r || !ic->e_or_s) return CCB_ERR_INVALID;
  int rc = allocate_particles(ctx, ic->n);
  if (rc) return rc;
  const int64_t n = ctx->n, cap = ctx->cap;
  ctx->t = ctx->cfg.t0;
  if (ctx->comm && ctx->p2p && !std::getenv("CCB_HALO_P2P")) {
    // Peer reads pay off while the halo is latency-bound. A large ghost column is re-read over NVLink by every boundary task
    // that stages it, where one message lands it in local memory (and L2) once: measured 7 % slower on config 4 cut into 8
    // slabs (56 MB per column) and 3 % faster on config 2 (0.9 MB). All ranks agree on the mean column size.
    double est = (double)ctx->n / std::max(1, ctx->slab_hi - ctx->slab_lo) * sizeof(double) * REC_G_MAX;
    double* d = reinterpret_cast<double*>(ctx->d_info + 8);
    CUDA_OK(cudaMemcpyAsync(d, &est, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    COMM_OK(ctx->comm->allreduce_sum(d, 1, ctx->stream));
    CUDA_OK(cudaMemcpyAsync(&est, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    if (est / ctx->cfg.nranks > 4.0e6) ctx->p2p = false;
  }
  if (ctx->comm && ctx->p2p) {  // peer-memory halo: every rank shows its (new) record array and flag words to its neighbours
    CUDA_OK(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int) * 8, ctx->stream));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    ctx->halo_epoch = 0;
    const int rank = ctx->cfg.rank, nranks = ctx->cfg.nranks;
    // ... and all ranks agree on whether that worked everywhere; if CUDA IPC is not available (container without a shared IPC
    // namespace, no peer access) every rank falls back to halo messages together
    double failed = ctx->comm->exchange_pointers(ctx->rec, ctx->d_flags, rank - 1, rank + 1 < nranks ? rank + 1 : -1, ctx->peer_rec, ctx->peer_flags,
                                                 ctx->stream) ? 1.0 : 0.0;
    double* d = reinterpret_cast<double*>(ctx->d_info + 8);
    CUDA_OK(cudaMemcpyAsync(d, &failed, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    COMM_OK(ctx->comm->allreduce_sum(d, 1, ctx->stream));
    CUDA_OK(cudaMemcpyAsync(&failed, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    if (failed > 0.0) {
      if (std::getenv("CCB_HALO_P2P")) {
        ctx->err = std::string(ctx->comm->name()) + " transport: peer-memory halo unavailable (" + ctx->comm->err + ")";
        return CCB_ERR_COMM;
      }
      ctx->p2p = false;
    }
  }
  const int64_t mig0 = ctx->migrated_out + ctx->migrated_in;
  DeviceView& dv = ctx->dv;
  // row-major host arrays go to the device as they are (one copy each) and are split into columns there
  auto stage_rows = [&](const double* host, int w, int c0, int cstep, int ncols, double* dst, const int32_t* id) -> int {
    CUDA_OK(cudaMemcpyAsync(ctx->d_stage, host, sizeof(double) * (size_t)w * n, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(gather_rows_by_id, cdiv(n, 256), 256, ctx->d_stage, w, c0, cstep, ncols, dst, cap, id, n);
    return CCB_OK;
  };
  // positions, velocities; sph_mass = extensive = 1 (process_initial_conditions, src/sph_workstation.cpp:1229-1340)
  if ((rc = stage_rows(ic->r, 3, 0, 1, ctx->D, dv.f[CCB_F_POSITION], nullptr))) return rc;
  if (ic->u) { if ((rc = stage_rows(ic->u, 3, 0, 1, ctx->D, dv.f[CCB_F_U], nullptr))) return rc; }
  for (int c = 0; c < 4; ++c) {
    LAUNCH(fill_kernel, cdiv(n, 256), 256, dv.f[CCB_F_SPH_MASS] + (int64_t)c * cap, 1.0, n);
    LAUNCH(fill_kernel, cdiv(n, 256), 256, dv.f[CCB_F_EXTENSIVE] + (int64_t)c * cap, 1.0, n);
  }
  // neighbour grid + initial sigma (calculate_intial_sigma, src/sph_workstation.cpp:583-640)
  if ((rc = do_build_cells(ctx))) return rc;
  if ((rc = do_prologue(ctx))) return rc;
  if ((rc = do_gather_fields(ctx))) return rc;
  // IC densities into temporary columns, in the current (cell-sorted) order
  CUDA_OK(cudaMalloc(&ctx->init_tmp, sizeof(double) * 19 * cap));
  CUDA_OK(cudaMemsetAsync(ctx->init_tmp, 0, sizeof(double) * 19 * cap, ctx->stream));
  const double* src1[5] = {ic->e_or_s, ic->rhoB, ic->rhoS, ic->rhoQ, ic->bulk};
  for (int k = 0; k < 5; ++k)
    if (src1[k] && (rc = stage_rows(src1[k], 1, 0, 1, 1, ctx->init_tmp + (int64_t)k * cap, dv.id))) return rc;
  // xx, xy, x-eta, yy, y-eta of the six-component row
  if (ic->shv6 && (rc = stage_rows(ic->shv6, 6, 0, 1, 5, ctx->init_tmp + (int64_t)5 * cap, dv.id))) return rc;
  ctx->init_diffusion = ic->diffusion != nullptr && ctx->cfg.using_diffusion;
  if (ctx->init_diffusion) {   // q^i_a, i = 1..3 of the [charge][mu] rows
    for (int c = 0; c < 3; ++c)
      if ((rc = stage_rows(ic->diffusion, 12, c * 4 + 1, 1, 3, ctx->init_tmp + (int64_t)(10 + c * 3) * cap, dv.id))) return rc;
  }
  const StageParams sp = make_stage(ctx, 0.0, 0);
  if ((rc = do_init_particles(ctx, sp))) return rc;
  // initial_smoothing (src/sph_workstation.cpp:395-457): smooth with the true extensive densities,
  // rebuild pi^{mu nu}, update thermodynamics
  if ((rc = do_prologue(ctx))) return rc;
  if ((rc = do_gather_fields(ctx))) return rc;
  if ((rc = do_eos_local(ctx, 1, 1))) return rc;
  // the initial freeze-out check of SPHWorkstation::initial_smoothing (src/sph_workstation.cpp:450-452), unconditional there
  if (!rc) rc = do_freeze_check_status(ctx, nullptr);
  if (!rc) rc = surface_device_error(ctx);
  cudaFree(ctx->init_tmp);
  ctx->init_tmp = nullptr;
  if (rc) return rc;
  if (ctx->migrated_out + ctx->migrated_in != mig0) { ctx->err = "ccb_upload_ic: rows outside this rank's slab (ccb_set_slab) were uploaded"; return CCB_ERR_INVALID; }
  if (ic->id) {  // local upload index -> global particle id (slab decomposition)
    int32_t* stage = reinterpret_cast<int32_t*>(ctx->d_stage);
    CUDA_OK(cudaMemcpyAsync(stage, ic->id, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(gather_int_by_id, cdiv(n, 256), 256, stage, ctx->dv.id, ctx->dv.id, n);
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
  }
  return CCB_OK;
}

}  // extern "C"
