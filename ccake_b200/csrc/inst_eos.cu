// Tabulated-EoS kernels (not templated on dimension or metric), built as four objects: -DCCB_EOS_PART=1..4 (eos_kernels.cuh).
#include <cstdlib>

#include "eos_kernels.cuh"
#include "launchers.h"

namespace ccb {

#if CCB_EOS_HAS(1)
// phase 04 of a table run: the two lean table-only kernels for the particles that stay on the table (todo, xs != nullptr), then the
// full chain for the particles they flagged. CCB_EOS_FAST=0 runs the full chain alone.
cudaError_t launch_eos_chain_stage(cudaStream_t s, const DeviceView& dv, const EosChain& ch, uint32_t* todo, double* xs) {
  static int mode = -1;
  if (mode < 0) { const char* e = std::getenv("CCB_EOS_FAST"); mode = e ? std::atoi(e) : 1; }
  const unsigned grid = (unsigned)((dv.n + 127) / 128);
  if (todo && xs && mode != 0) {
    // launch bounds measured at 2e6 particles (profiles/r02ah): search at 168 registers / 3 CTAs per SM 1.44 ms, at 128 / 4 1.61 ms,
    // at 96 / 5 1.84 ms (the search spills below 168); epilogue at 128 / 4 1.61 ms, at 80 / 6 1.67 ms
    eos_table_root_kernel<3><<<grid, 128, 0, s>>>(dv, ch, todo, xs);
    eos_table_epilogue_kernel<4><<<grid, 128, 0, s>>>(dv, ch, todo, xs);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    eos_chain_stage_kernel<<<grid, 128, 0, s>>>(dv, ch, todo);
  } else {
    eos_chain_stage_kernel<<<grid, 128, 0, s>>>(dv, ch, nullptr);
  }
  return cudaGetLastError();
}
#endif
#if CCB_EOS_HAS(2)
cudaError_t launch_eos_chain_init(cudaStream_t s, const DeviceView& dv, const EosChain& ch, const double* tmp, int input_as_entropy) {
  eos_chain_init_kernel<<<(unsigned)((dv.n + 127) / 128), 128, 0, s>>>(dv, ch, tmp, input_as_entropy);
  return cudaGetLastError();
}

#endif
#if CCB_EOS_HAS(3)
cudaError_t launch_eos_query(cudaStream_t s, const EosChain& ch, const ConformalParams* conf, int member, int64_t n, const double* pts, double* out) {
  ConformalParams c{};
  if (conf) c = *conf;
  eos_query_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(ch, conf, c, member, n, pts, out);
  return cudaGetLastError();
}
cudaError_t launch_eos_locate(cudaStream_t s, const EosChain& ch, const ConformalParams& conf, int use_chain, int energy, int64_t n,
                              const double* dens, const double* seed, const int32_t* start, double* thermo, int32_t* eos_out) {
  eos_locate_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(ch, conf, use_chain, energy, n, dens, seed, start, thermo, eos_out);
  return cudaGetLastError();
}

#endif
#if CCB_EOS_HAS(4)
cudaError_t launch_pre_fill(cudaStream_t s, const DeviceView& dv, const PreTable& tb) {
  pre_fill_kernel<<<(unsigned)((dv.n + 127) / 128), 128, 0, s>>>(dv, tb);
  return cudaGetLastError();
}
cudaError_t launch_pre_init(cudaStream_t s, const DeviceView& dv, const PreFlat& fl, const double* tmp, int input_as_entropy) {
  pre_init_kernel<<<(unsigned)((dv.n + 127) / 128), 128, 0, s>>>(dv, fl, tmp, input_as_entropy);
  return cudaGetLastError();
}
cudaError_t launch_pre_generate(cudaStream_t s, const EosChain& ch, const ConformalParams& conf, int use_chain, const int32_t exponent[4], int N,
                                double* vars, int32_t* eos) {
  const int64_t total = (int64_t)N * N * N * N;
  pre_generate_kernel<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(ch, conf, use_chain, exponent[0], exponent[1], exponent[2], exponent[3], N, vars, eos);
  return cudaGetLastError();
}

#endif
}  // namespace ccb
