// Launch entry points of the templated kernels. Each (D, metric) instantiation lives in
// its own object file (inst_gather.cu / inst_local.cu compiled once per -DCCB_INST_*
// combination by __graft_entry__.build()), so the library builds in parallel.
#pragma once

#include "gather_params.cuh"
#include "stage_params.h"

namespace ccb {

// launch shape of a gather: nt = targets (= threads) per task (64 | 128), tile = source records per shared-memory tile (32 | 64)
struct GatherShape { int nt, tile; };
template <int D> cudaError_t launch_gather_fields(bool fact, const GatherShape& sh, cudaStream_t s, const DeviceView& dv, const CellGrid& g,
                                                  const GatherParams& gp);
template <int D> cudaError_t launch_gather_gradients(bool shear, bool diff, bool fact, const GatherShape& sh, cudaStream_t s, const DeviceView& dv,
                                                     const CellGrid& g, const GatherParams& gp);
template <int D, bool MILNE> cudaError_t launch_prologue(cudaStream_t s, const DeviceView& dv, double t);
template <int D, bool MILNE> cudaError_t launch_eos_local(cudaStream_t s, const DeviceView& dv, const StageParams& sp);
template <int D, bool MILNE> cudaError_t launch_eom_rk(cudaStream_t s, const DeviceView& dv, const StageParams& sp);
template <int D, bool MILNE>
cudaError_t launch_init_particles(cudaStream_t s, const DeviceView& dv, const StageParams& sp, const double* tmp,
                                  int input_as_entropy, int input_initial_shear, int bulk_from_trace, double efcheck, int eos_done);
struct EosChain;
// tabulated EoS + fall-back chain (eos_kernels.cuh): phase 04 of a derivative evaluation, and the initial e|s -> (T, mu) location
cudaError_t launch_eos_chain_stage(cudaStream_t s, const DeviceView& dv, const EosChain& ch, uint32_t* todo, double* xs);
cudaError_t launch_eos_chain_init(cudaStream_t s, const DeviceView& dv, const EosChain& ch, const double* tmp, int input_as_entropy);
// host-callable EoS queries (ccb_eos_full_thermo, ccb_eos_locate)
cudaError_t launch_eos_query(cudaStream_t s, const EosChain& ch, const ConformalParams* conf, int member, int64_t n, const double* pts, double* out);
cudaError_t launch_eos_locate(cudaStream_t s, const EosChain& ch, const ConformalParams& conf, int use_chain, int energy, int64_t n,
                              const double* dens, const double* seed, const int32_t* start, double* thermo, int32_t* eos_out);
// pre-inverted EoS (eos_preinverted.cuh): phase 04 lookup, initial location on the flat table, tile generator
struct PreTable;
struct PreFlat;
cudaError_t launch_pre_fill(cudaStream_t s, const DeviceView& dv, const PreTable& tb);
cudaError_t launch_pre_init(cudaStream_t s, const DeviceView& dv, const PreFlat& fl, const double* tmp, int input_as_entropy);
cudaError_t launch_pre_generate(cudaStream_t s, const EosChain& ch, const ConformalParams& conf, int use_chain, const int32_t exponent[4], int N,
                                double* vars, int32_t* eos);
// freeze-out (freeze_out.cuh)
struct FreezeArrays;
cudaError_t launch_freeze_check_status(cudaStream_t s, const DeviceView& dv, const FreezeArrays& fa, double time_in, double efcheck);
template <int D, bool MILNE> cudaError_t launch_freeze_snapshot(cudaStream_t s, const DeviceView& dv, const FreezeArrays& fa, int which);
template <int D, bool MILNE> cudaError_t launch_freeze_evo(cudaStream_t s, const DeviceView& dv, const FreezeArrays& fa);
template <int D, bool MILNE> cudaError_t launch_freeze_advance(cudaStream_t s, const DeviceView& dv, const FreezeArrays& fa);
template <int D, bool MILNE>
cudaError_t launch_freeze_interpolate(cudaStream_t s, const DeviceView& dv, const FreezeArrays& fa, double efcheck, double taup, double taupp, double dt);
template <bool MILNE>
cudaError_t launch_reduce_totals(int nblocks, cudaStream_t s, const DeviceView& dv, double t, double* out,
                                 unsigned long long* nfrozen);

}  // namespace ccb
