// Tabulated-EoS kernels (not templated on dimension or metric): one object file.
#include "eos_kernels.cuh"
#include "launchers.h"

namespace ccb {

cudaError_t launch_eos_chain_stage(cudaStream_t s, const DeviceView& dv, const EosChain& ch) {
  eos_chain_stage_kernel<<<(unsigned)((dv.n + 127) / 128), 128, 0, s>>>(dv, ch);
  return cudaGetLastError();
}
cudaError_t launch_eos_chain_init(cudaStream_t s, const DeviceView& dv, const EosChain& ch, const double* tmp, int input_as_entropy) {
  eos_chain_init_kernel<<<(unsigned)((dv.n + 127) / 128), 128, 0, s>>>(dv, ch, tmp, input_as_entropy);
  return cudaGetLastError();
}

}  // namespace ccb
