// Instantiations of the two neighbour gathers for one spatial dimension (-DCCB_INST_D=2|3).
#include "gather.cuh"
#include "launchers.h"

#ifndef CCB_INST_D
#error "compile with -DCCB_INST_D=2 or 3"
#endif

namespace ccb {

template <>
cudaError_t launch_gather_fields<CCB_INST_D>(bool fact, int grid, int block, cudaStream_t s, const DeviceView& dv,
                                             const CellGrid& g, const GatherParams& gp) {
  constexpr int D = CCB_INST_D;
  if (D == 3 && fact) gather_fields_kernel<D, (D == 3)><<<grid, block, 0, s>>>(dv, g, gp);
  else gather_fields_kernel<D, false><<<grid, block, 0, s>>>(dv, g, gp);
  return cudaGetLastError();
}

template <>
cudaError_t launch_gather_gradients<CCB_INST_D>(bool shear, bool diff, bool fact, int grid, int block, cudaStream_t s,
                                                const DeviceView& dv, const CellGrid& g, const GatherParams& gp) {
  constexpr int D = CCB_INST_D;
#define CCB_GG(S, F, X) gather_gradients_kernel<D, S, F, X><<<grid, block, 0, s>>>(dv, g, gp)
  if (D == 3 && fact) {
    if (shear) { if (diff) CCB_GG(true, true, (D == 3)); else CCB_GG(true, false, (D == 3)); }
    else { if (diff) CCB_GG(false, true, (D == 3)); else CCB_GG(false, false, (D == 3)); }
  } else {
    if (shear) { if (diff) CCB_GG(true, true, false); else CCB_GG(true, false, false); }
    else { if (diff) CCB_GG(false, true, false); else CCB_GG(false, false, false); }
  }
#undef CCB_GG
  return cudaGetLastError();
}

}  // namespace ccb
