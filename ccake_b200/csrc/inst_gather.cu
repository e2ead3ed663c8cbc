// Instantiations of the two neighbour gathers for one spatial dimension (-DCCB_INST_D=2|3).
#include "gather.cuh"
#include "launchers.h"

#ifndef CCB_INST_D
#error "compile with -DCCB_INST_D=2 or 3"
#endif

namespace ccb {

namespace {
// persistent grid: as many CTAs as fit on the device at once; tasks come from the atomic counter
template <class P, int NT>
cudaError_t launch(int sms, cudaStream_t s, const DeviceView& dv, const CellGrid& g, const GatherParams& gp, int* task_counter) {
  static int per_sm = 0;
  if (per_sm == 0) {
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gather_kernel<P, NT>, NT, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
  }
  gather_kernel<P, NT><<<sms * per_sm, NT, 0, s>>>(dv, g, gp, task_counter);
  return cudaGetLastError();
}
template <class P>
cudaError_t launch_nt(int nt, int sms, cudaStream_t s, const DeviceView& dv, const CellGrid& g, const GatherParams& gp, int* tc) {
  return nt == 128 ? launch<P, 128>(sms, s, dv, g, gp, tc) : launch<P, 64>(sms, s, dv, g, gp, tc);
}
}  // namespace

template <>
cudaError_t launch_gather_fields<CCB_INST_D>(bool fact, int nt, int sms, cudaStream_t s, const DeviceView& dv, const CellGrid& g,
                                             const GatherParams& gp, int* tc) {
  constexpr int D = CCB_INST_D;
  if (D == 3 && fact) return launch_nt<FieldsPolicy<D, (D == 3)>>(nt, sms, s, dv, g, gp, tc);
  return launch_nt<FieldsPolicy<D, false>>(nt, sms, s, dv, g, gp, tc);
}

template <>
cudaError_t launch_gather_gradients<CCB_INST_D>(bool shear, bool diff, bool fact, int nt, int sms, cudaStream_t s,
                                                const DeviceView& dv, const CellGrid& g, const GatherParams& gp, int* tc) {
  constexpr int D = CCB_INST_D;
#define CCB_GG(S, F, X) return launch_nt<GradPolicy<D, S, F, X>>(nt, sms, s, dv, g, gp, tc)
  if (D == 3 && fact) {
    if (shear) { if (diff) CCB_GG(true, true, (D == 3)); else CCB_GG(true, false, (D == 3)); }
    else { if (diff) CCB_GG(false, true, (D == 3)); else CCB_GG(false, false, (D == 3)); }
  } else {
    if (shear) { if (diff) CCB_GG(true, true, false); else CCB_GG(true, false, false); }
    else { if (diff) CCB_GG(false, true, false); else CCB_GG(false, false, false); }
  }
#undef CCB_GG
}

}  // namespace ccb
