// Per-stage parameter block handed to the per-particle kernels, and its
// construction from the public ccb_config (include/ccake_b200.h). Kept free of
// CUDA syntax so that tests/host_harness.cpp maps a config exactly as the library does.
#pragma once

#include <cstring>

#include "../../include/ccake_b200.h"
#include "physics_eom.cuh"

namespace ccb {

struct StageParams {
  double t;          // time of this derivative evaluation
  double h_stage;    // RK stage step (y = y0 + h_stage * f)
  int do_update;     // apply the RK update after evaluating f
  int do_eos, do_local;
  EomFlags fl;
  TransportParams tp;
  ConformalParams conf;
  int eos_type;
  int prohibit_unstable_cs2, prohibit_acausal_cs2;
  // SPHWorkstation::regulate_diffusion (src/sph_workstation.cpp:133-218): e0, xi0 already in fm^-4
  int reg_enabled;
  double reg_chi0, reg_e0, reg_xi0, reg_rq_max;
};

inline StageParams make_stage_params(const ccb_config& c, double t, double h_stage, int do_update) {
  StageParams sp;
  std::memset(&sp, 0, sizeof(sp));
  sp.t = t; sp.h_stage = h_stage; sp.do_update = do_update; sp.do_eos = 1; sp.do_local = 1;
  sp.fl.shear = c.using_shear != 0; sp.fl.diffusion = c.using_diffusion != 0; sp.fl.vorticity = c.using_vorticity != 0;
  TransportParams& tp = sp.tp;
  tp.shear_mode = c.shear_mode; tp.shear_relax_mode = c.shear_relax_mode; tp.zeta_mode = c.zeta_mode;
  tp.bulk_relax_mode = c.bulk_relax_mode; tp.diffusion_mode = c.diffusion_mode; tp.diffusion_relax_mode = c.diffusion_relax_mode;
  tp.delta_pipi_mode = c.delta_pipi_mode; tp.tau_pipi_mode = c.tau_pipi_mode; tp.lambda_piPi_mode = c.lambda_piPi_mode;
  tp.phi6_mode = c.phi6_mode; tp.phi7_mode = c.phi7_mode; tp.delta_PiPi_mode = c.delta_PiPi_mode;
  tp.lambda_Pipi_mode = c.lambda_Pipi_mode; tp.phi1_mode = c.phi1_mode; tp.phi3_mode = c.phi3_mode;
  tp.modulate_zeta_with_tanh = c.modulate_zeta_with_tanh; tp.critical_scaling_bulk = c.critical_scaling_bulk;
  tp.eta_over_s = c.constant_eta_over_s; tp.zeta_over_s = c.constant_zeta_over_s; tp.zeta_A = c.cs2_dependent_zeta_A;
  tp.zeta_p = c.cs2_dependent_zeta_p; tp.C_B = c.C_B;
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) tp.kappa_over_T2[a][b] = c.constant_kappa_over_T2[a * 3 + b];
  tp.critical_scaling_diffusion = c.critical_scaling_diffusion;   // MeV -> fm^-1 (hbarc_MeVfm, include/constants.h:16)
  tp.crit_T = c.critical_point_T / 197.3269804; tp.crit_muB = c.critical_point_muB / 197.3269804;
  tp.crit_wT = c.critical_gaussian_width_T / 197.3269804; tp.crit_wmuB = c.critical_gaussian_width_muB / 197.3269804;
  sp.conf.c = c.conformal_c; sp.conf.T0 = c.conformal_T0; sp.conf.muB0 = c.conformal_muB0; sp.conf.muS0 = c.conformal_muS0;
  sp.conf.muQ0 = c.conformal_muQ0;
  sp.reg_enabled = (c.using_diffusion != 0 && c.diffusion_regulator_enabled != 0) ? 1 : 0;
  sp.reg_chi0 = c.diffusion_regulator_chi0; sp.reg_rq_max = c.diffusion_regulator_rq_max;
  sp.reg_e0 = c.diffusion_regulator_e0 / 0.1973269804; sp.reg_xi0 = c.diffusion_regulator_xi0 / 0.1973269804;  // hbarc_GeVfm, include/constants.h:15
  sp.eos_type = c.eos_type; sp.prohibit_unstable_cs2 = c.prohibit_unstable_cs2; sp.prohibit_acausal_cs2 = c.prohibit_acausal_cs2;
  return sp;
}

}  // namespace ccb
