// Instantiations of the per-particle kernels for one (dimension, metric) pair
// (-DCCB_INST_D=2|3 -DCCB_INST_MILNE=0|1).
#include "freeze_out.cuh"
#include "local.cuh"
#include "launchers.h"

#if !defined(CCB_INST_D) || !defined(CCB_INST_MILNE)
#error "compile with -DCCB_INST_D=2|3 -DCCB_INST_MILNE=0|1"
#endif

namespace ccb {

namespace { inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; } }
constexpr int D_ = CCB_INST_D;
constexpr bool M_ = CCB_INST_MILNE != 0;

template <> cudaError_t launch_prologue<D_, M_>(cudaStream_t s, const DeviceView& dv, double t) {
  prologue_kernel<D_, M_><<<(unsigned)cdiv(dv.n, 256), 256, 0, s>>>(dv, t);
  return cudaGetLastError();
}
template <> cudaError_t launch_eos_local<D_, M_>(cudaStream_t s, const DeviceView& dv, const StageParams& sp) {
  eos_local_kernel<D_, M_><<<(unsigned)cdiv(dv.n, 128), 128, 0, s>>>(dv, sp);
  return cudaGetLastError();
}
template <> cudaError_t launch_eom_rk<D_, M_>(cudaStream_t s, const DeviceView& dv, const StageParams& sp) {
  eom_rk_kernel<D_, M_><<<(unsigned)cdiv(dv.n, 128), 128, 0, s>>>(dv, sp);
  return cudaGetLastError();
}
template <>
cudaError_t launch_init_particles<D_, M_>(cudaStream_t s, const DeviceView& dv, const StageParams& sp, const double* tmp,
                                          int input_as_entropy, int input_initial_shear, int bulk_from_trace, double efcheck, int eos_done) {
  init_particles_kernel<D_, M_><<<(unsigned)cdiv(dv.n, 128), 128, 0, s>>>(dv, sp, tmp, input_as_entropy, input_initial_shear,
                                                                          bulk_from_trace, efcheck, eos_done);
  return cudaGetLastError();
}
// freeze-out (freeze_out.cuh)
template <> cudaError_t launch_freeze_snapshot<D_, M_>(cudaStream_t s, const DeviceView& dv, const FreezeArrays& fa, int which) {
  freeze_snapshot_kernel<D_><<<(unsigned)cdiv(dv.n, 128), 128, 0, s>>>(dv, fa, which);
  return cudaGetLastError();
}
template <> cudaError_t launch_freeze_evo<D_, M_>(cudaStream_t s, const DeviceView& dv, const FreezeArrays& fa) {
  freeze_evo_kernel<D_><<<(unsigned)cdiv(dv.n, 128), 128, 0, s>>>(dv, fa);
  return cudaGetLastError();
}
template <> cudaError_t launch_freeze_advance<D_, M_>(cudaStream_t s, const DeviceView& dv, const FreezeArrays& fa) {
  freeze_advance_kernel<D_><<<(unsigned)cdiv(dv.n, 128), 128, 0, s>>>(dv, fa);
  return cudaGetLastError();
}
template <> cudaError_t launch_freeze_interpolate<D_, M_>(cudaStream_t s, const DeviceView& dv, const FreezeArrays& fa, double efcheck, double taup,
                                                          double taupp, double dt) {
  freeze_interpolate_kernel<D_, M_><<<(unsigned)cdiv(dv.n, 128), 128, 0, s>>>(dv, fa, efcheck, taup, taupp, dt);
  return cudaGetLastError();
}
#if CCB_INST_D == 2 && CCB_INST_MILNE == 1
cudaError_t launch_freeze_check_status(cudaStream_t s, const DeviceView& dv, const FreezeArrays& fa, double time_in, double efcheck) {
  freeze_check_status_kernel<0><<<(unsigned)cdiv(dv.n, 128), 128, 0, s>>>(dv, fa, time_in, efcheck);
  return cudaGetLastError();
}
#endif
#if CCB_INST_D == 2
template <>
cudaError_t launch_reduce_totals<M_>(int nblocks, cudaStream_t s, const DeviceView& dv, double t, double* out,
                                     unsigned long long* nfrozen) {
  reduce_totals_kernel<M_><<<nblocks, 256, 0, s>>>(dv, t, out, nfrozen);
  return cudaGetLastError();
}
#endif

}  // namespace ccb
