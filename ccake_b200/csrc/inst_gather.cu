// Instantiations of the two neighbour gathers for one spatial dimension (-DCCB_INST_D=2|3).
#include "gather.cuh"
#include "launchers.h"

#ifndef CCB_INST_D
#error "compile with -DCCB_INST_D=2 or 3"
#endif

namespace ccb {

namespace {
// one CTA per task of NT consecutive sorted particles; the hardware block scheduler balances the SMs
template <class P, int NT, int MINB, int TILE>
cudaError_t launch(cudaStream_t s, const DeviceView& dv, const CellGrid& g, const GatherParams& gp) {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(gather_kernel<P, NT, MINB, TILE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  GatherParams q = gp;
  unsigned ntasks = (unsigned)q.task_count;
  if (q.task_count <= 0) { ntasks = (unsigned)((dv.n + NT - 1) / NT); q.task_n0 = (int)ntasks; q.task_off0 = q.task_off1 = 0; }
  gather_kernel<P, NT, MINB, TILE><<<ntasks, NT, 0, s>>>(dv, g, q);
  return cudaGetLastError();
}
// nt = targets per task (64 | 128) + 1000 x (source records per shared-memory tile: 32 | 64; default 64). Tiles are
// double-buffered, so 128-record tiles of the largest record would not fit the 48 KB of static shared memory.
template <class P, int MINB64, int MINB128>
cudaError_t launch_nt(int nt, cudaStream_t s, const DeviceView& dv, const CellGrid& g, const GatherParams& gp) {
  const int tile = nt / 1000;
  nt %= 1000;
  if (tile == 32) return nt == 128 ? launch<P, 128, MINB128, 32>(s, dv, g, gp) : launch<P, 64, MINB64, 32>(s, dv, g, gp);
  return nt == 128 ? launch<P, 128, MINB128, 64>(s, dv, g, gp) : launch<P, 64, MINB64, 64>(s, dv, g, gp);
}
}  // namespace

template <>
cudaError_t launch_gather_fields<CCB_INST_D>(bool fact, int nt, cudaStream_t s, const DeviceView& dv, const CellGrid& g,
                                             const GatherParams& gp) {
  constexpr int D = CCB_INST_D;
  if (D == 3 && fact) return launch_nt<FieldsPolicy<D, (D == 3)>, 12, 6>(nt, s, dv, g, gp);
  return launch_nt<FieldsPolicy<D, false>, 12, 6>(nt, s, dv, g, gp);
}

template <>
cudaError_t launch_gather_gradients<CCB_INST_D>(bool shear, bool diff, bool fact, int nt, cudaStream_t s,
                                                const DeviceView& dv, const CellGrid& g, const GatherParams& gp) {
  constexpr int D = CCB_INST_D;
  // register budget: 128 per thread (16 warps per SM) for every variant. The largest one (D = 3, shear + diffusion: 43
  // accumulators) then spills ~200 bytes inside the neighbour loop, which still measured 6 % faster than 168 registers
  // without spills at 12 warps per SM (profiles/README.md, r01c).
#define CCB_GG(S, F, X) return launch_nt<GradPolicy<D, S, F, X>, 8, 4>(nt, s, dv, g, gp)
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
