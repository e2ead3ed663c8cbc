// Instantiations of the two neighbour gathers for one spatial dimension (-DCCB_INST_D=2|3).
#include "gather.cuh"
#include "launchers.h"

#ifndef CCB_INST_D
#error "compile with -DCCB_INST_D=2 or 3"
#endif

namespace ccb {

namespace {
// persistent CTAs, as many as fit the device at once; they pull tasks (cells.cuh: task list) from GatherParams::task_counter
template <class P, int NT, int MINB, int TILE>
cudaError_t launch(cudaStream_t s, const DeviceView& dv, const CellGrid& g, const GatherParams& gp) {
  static int grid = 0;
  if (!grid) {
    cudaFuncSetAttribute(gather_kernel<P, NT, MINB, TILE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gather_kernel<P, NT, MINB, TILE>, NT, 0);
    grid = sms * (per_sm > 0 ? per_sm : 1);
  }
  // no more CTAs than there can be tasks (small systems): every particle starts at most one task
  const int64_t most = dv.n > 0 ? dv.n : 1;
  gather_kernel<P, NT, MINB, TILE><<<(unsigned)(most < grid ? most : grid), NT, 0, s>>>(dv, g, gp);
  return cudaGetLastError();
}
// GatherShape: targets per task (64 | 128) and source records per shared-memory tile (32 | 64; tiles are double-buffered).
template <class P, int MINB64, int MINB128>
cudaError_t launch_shape(const GatherShape& sh, cudaStream_t s, const DeviceView& dv, const CellGrid& g, const GatherParams& gp) {
  if (sh.tile == 32) return sh.nt == 128 ? launch<P, 128, MINB128, 32>(s, dv, g, gp) : launch<P, 64, MINB64, 32>(s, dv, g, gp);
  return sh.nt == 128 ? launch<P, 128, MINB128, 64>(s, dv, g, gp) : launch<P, 64, MINB64, 64>(s, dv, g, gp);
}
// Register budget of a gradient variant, by its accumulator count (measured on the five BASELINE configs, profiles/r02aa_*): the
// variants with up to 30 fp64 accumulators (2-D all, 3-D without diffusion) run fastest at 102 registers = 10 CTAs of 64 threads per
// SM (-1 .. -9 % against 128 registers), the 3-D shear + diffusion variants (37 | 43 accumulators) need their 128 registers (102: +60 %,
// 85: +190 %, 168: +6 %).
template <class P>
cudaError_t launch_regs(const GatherShape& sh, cudaStream_t s, const DeviceView& dv, const CellGrid& g, const GatherParams& gp) {
  constexpr int M = P::NACC <= 30 ? 10 : 8;
  return launch_shape<P, M, M / 2>(sh, s, dv, g, gp);
}
// a diffusion run launches both compile-time sparsity twins; the one that does not match DeviceView::qflag exits at once
template <int D, bool S, bool F, bool X>
cudaError_t launch_grad(const GatherShape& sh, cudaStream_t s, const DeviceView& dv, const CellGrid& g, const GatherParams& gp) {
  if (!F) return launch_regs<GradPolicy<D, S, false, X, false>>(sh, s, dv, g, gp);
  cudaError_t e = launch_regs<GradPolicy<D, S, F, X, false>>(sh, s, dv, g, gp);
  if (e != cudaSuccess) return e;
  return launch_regs<GradPolicy<D, S, F, X, true>>(sh, s, dv, g, gp);
}
}  // namespace

template <>
cudaError_t launch_gather_fields<CCB_INST_D>(bool fact, const GatherShape& sh, cudaStream_t s, const DeviceView& dv, const CellGrid& g,
                                             const GatherParams& gp) {
  constexpr int D = CCB_INST_D;
  if (gp.nbr_idx) return launch_shape<FieldsPolicy<D, false, true>, 8, 4>(sh, s, dv, g, gp);  // debug twin: neighbour ids (plain kernel)
  if (D == 3 && fact) return launch_shape<FieldsPolicy<D, (D == 3)>, 16, 8>(sh, s, dv, g, gp);
  return launch_shape<FieldsPolicy<D, false>, 16, 8>(sh, s, dv, g, gp);
}

template <>
cudaError_t launch_gather_gradients<CCB_INST_D>(bool shear, bool diff, bool fact, const GatherShape& sh, cudaStream_t s,
                                                const DeviceView& dv, const CellGrid& g, const GatherParams& gp) {
  constexpr int D = CCB_INST_D;
  if (D == 3 && fact) {
    if (shear) return diff ? launch_grad<D, true, true, (D == 3)>(sh, s, dv, g, gp) : launch_grad<D, true, false, (D == 3)>(sh, s, dv, g, gp);
    return diff ? launch_grad<D, false, true, (D == 3)>(sh, s, dv, g, gp) : launch_grad<D, false, false, (D == 3)>(sh, s, dv, g, gp);
  }
  if (shear) return diff ? launch_grad<D, true, true, false>(sh, s, dv, g, gp) : launch_grad<D, true, false, false>(sh, s, dv, g, gp);
  return diff ? launch_grad<D, false, true, false>(sh, s, dv, g, gp) : launch_grad<D, false, false, false>(sh, s, dv, g, gp);
}

}  // namespace ccb
