// Phase 04 for a tabulated equation of state: one thread locates one particle in the phase diagram
// (physics_eos_chain.cuh) and writes its thermodynamics. The reference does this as a serial host loop between
// two deep copies of the whole particle array (src/sph_workstation.cpp:1714-1772).
// Bound: fp64 pipe + L2 latency (16 scattered 32-byte vertex reads per Newton iterate from the 4-field "hot"
// table, which is sized to stay L2-resident); the 17-field table is touched once per particle, after convergence.
#pragma once

#include "local.cuh"
#include "physics_eos_chain.cuh"

namespace ccb {

// (s, rhoB, rhoS, rhoQ) of the local rest frame (thermo columns 5..8, written by the field gather) -> thermo.
// SPHWorkstation::locate_phase_diagram_point_sBSQ, src/sph_workstation.cpp:659-741
__global__ void __launch_bounds__(128) eos_chain_stage_kernel(DeviceView dv, EosChain ch) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dv.n) return;
  const double s = col(dv.f[CCB_F_THERMO], dv.cap, 5, i), B = col(dv.f[CCB_F_THERMO], dv.cap, 6, i),
               S = col(dv.f[CCB_F_THERMO], dv.cap, 7, i), Q = col(dv.f[CCB_F_THERMO], dv.cap, 8, i);
  // previous point: TSEED holds (T, muB, muS, muQ); the solver's order is (T, muB, muQ, muS)
  const double pos[4] = {col(dv.f[CCB_F_TSEED], dv.cap, 0, i), col(dv.f[CCB_F_TSEED], dv.cap, 1, i),
                         col(dv.f[CCB_F_TSEED], dv.cap, 3, i), col(dv.f[CCB_F_TSEED], dv.cap, 2, i)};
  ChainResult res;
  if (!chain_update_s(ch, dv.eos_id[i], pos, s, B, S, Q, res)) { raise_error(dv, CCB_ERR_EOS_NO_ROOT, dv.id[i]); return; }
  Thermo th;
  chain_thermo(ch, res, th);
  store_thermo(dv, i, th);
  dv.eos_id[i] = res.eos;
}

// Initial location from the IC densities (tmp columns: 0 e [GeV/fm^3] or s, 1 rhoB, 2 rhoS, 3 rhoQ).
// SPHWorkstation::initialize_entropy_and_charge_densities -> locate_phase_diagram_point_{eBSQ,sBSQ}
// (src/sph_workstation.cpp:253-320, 1791-1848): a particle without a solution is frozen out and located again
// with its charge densities set to zero.
__global__ void __launch_bounds__(128) eos_chain_init_kernel(DeviceView dv, EosChain ch, const double* __restrict__ tmp, int input_as_entropy) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dv.n) return;
  const int64_t cap = dv.cap;
  const double rhoB = tmp[1 * cap + i], rhoS = tmp[2 * cap + i], rhoQ = tmp[3 * cap + i];
  ChainResult res;
  bool ok;
  if (input_as_entropy) {
    const double pos[4] = {580.0 / kHbarcMeVfm, 0.0, 0.0, 0.0};  // process_initial_conditions, src/sph_workstation.cpp:1291-1298
    ok = chain_update_s(ch, 0, pos, tmp[i], rhoB, rhoS, rhoQ, res);
    if (!ok) { raise_error(dv, CCB_ERR_EOS_NO_ROOT, dv.id[i]); return; }
  } else {
    const double e_in = tmp[i] / kHbarcGeVfm;
    ok = chain_s_out(ch, e_in, rhoB, rhoS, rhoQ, res);
    if (!ok) {
      dv.freeze[i] = 4;
      ok = chain_s_out(ch, e_in, 0.0, 0.0, 0.0, res);
      if (!ok) { raise_error(dv, CCB_ERR_EOS_NO_ROOT, dv.id[i]); return; }
    }
  }
  Thermo th;
  chain_thermo(ch, res, th);
  store_thermo(dv, i, th);
  dv.eos_id[i] = res.eos;
}

}  // namespace ccb
