// Drop-in layer, part 1 of 8: the device context behind the replacement classes, and Settings -> ccb_config.
//
// The headers of include/ccake_dropin/ carry the NAMES AND SIGNATURES of the reference classes they replace
// (system_state.h, sph_workstation.h, evolver.h, eom_default.h, eom_cartesian.h, eos.h, eos_interpolator.h, freeze_out.h), so
// that src/ccake.cpp and BSQHydro<D,TEOM> (src/BSQHydro.cpp) compile against them unchanged. They are thin: every method
// forwards to the C ABI of include/ccake_b200.h. The reference headers that stay the maintainer's own -- settings.h,
// particle.h, hydrodynamic_info.h, thermodynamic_info.h, densities.h, vector.h, matrix.h, constants.h, formatted_output.h,
// phase_profiler.h -- are included by their reference names and found on the maintainer's include path.
//
// Settings -> ccb_config follows Input::decode_settings (src/input.cpp:96-912) for the derived switches and
// transport_coefficients::setup_parameters (src/transport_coefficients.cpp:9-279) for the mode strings.
#ifndef CCAKE_DROPIN_DEVICE_CONTEXT_H
#define CCAKE_DROPIN_DEVICE_CONTEXT_H

#include <array>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>

#include "ccake_b200.h"
#include "settings.h"

namespace ccake {
namespace b200 {

inline int mode_of(const std::map<std::string, int>& table, const std::string& v) {
  auto it = table.find(v);
  return it == table.end() ? 0 : it->second;  // unknown strings select the reference's default branch
}

inline ccb_config to_config(const Settings& s, int device = 0, int rank = 0, int nranks = 1) {
  ccb_config c;
  std::memset(&c, 0, sizeof(c));
  c.dim = 0;  // filled in by the caller (the template parameter D)
  c.coordinate_system = s.coordinate_system == "cartesian" ? CCB_COORD_CARTESIAN : CCB_COORD_HYPERBOLIC;
  c.rk_order = s.rk_order;
  c.factorized_kernel = s.factorized_kernel ? 1 : 0;
  c.t0 = s.t0; c.hT = s.hT; c.hEta = s.hEta;
  c.xmin = s.xmin; c.ymin = s.ymin; c.etamin = s.etamin;
  c.stepx = s.stepx; c.stepy = s.stepy; c.stepEta = s.stepEta;
  c.using_shear = s.using_shear ? 1 : 0; c.using_diffusion = s.using_diffusion ? 1 : 0; c.using_vorticity = s.using_vorticity ? 1 : 0;
  c.input_as_entropy = s.input_as_entropy ? 1 : 0; c.input_initial_shear = s.input_initial_shear ? 1 : 0; c.bulk_from_trace = s.bulk_from_trace ? 1 : 0;
  static const std::map<std::string, int> SHEAR{{"default", 0}, {"constant", 1}, {"JakiParam", 2}, {"LinearMus", 3}, {"interpolate", 4}};
  static const std::map<std::string, int> TAU_PI{{"default", 0}, {"minVal", 1}, {"Gubser", 2}};
  static const std::map<std::string, int> ZETA{{"default", 0}, {"constant", 1}, {"DNMR", 2}, {"interpolate", 3}, {"cs2_dependent", 4}};
  static const std::map<std::string, int> TAU_BULK{{"default", 0}, {"DNMR", 1}};
  static const std::map<std::string, int> KAPPA{{"constant_over_T2", 0}, {"DNMR", 1}, {"kappaB_coth", 2}};
  static const std::map<std::string, int> TAU_Q{{"constant_over_T", 0}, {"cb_over_T", 1}};
  static const std::map<std::string, int> DELTA{{"default", 0}, {"israel-stewart", 1}, {"disabled", 2}};
  static const std::map<std::string, int> ONOFF{{"default", 0}, {"disabled", 1}};
  c.shear_mode = mode_of(SHEAR, s.etaMode); c.shear_relax_mode = mode_of(TAU_PI, s.shearRelaxMode);
  c.zeta_mode = mode_of(ZETA, s.zetaMode); c.bulk_relax_mode = mode_of(TAU_BULK, s.bulkRelaxMode);
  c.diffusion_mode = mode_of(KAPPA, s.diffusionMode); c.diffusion_relax_mode = mode_of(TAU_Q, s.diffusionRelaxMode);
  c.delta_pipi_mode = mode_of(DELTA, s.delta_pipi_mode); c.tau_pipi_mode = mode_of(ONOFF, s.tau_pipi_mode);
  c.lambda_piPi_mode = mode_of(ONOFF, s.lambda_piPi_mode); c.phi6_mode = mode_of(ONOFF, s.phi6_mode); c.phi7_mode = mode_of(ONOFF, s.phi7_mode);
  c.delta_PiPi_mode = mode_of(DELTA, s.delta_PiPi_mode); c.lambda_Pipi_mode = mode_of(ONOFF, s.lambda_Pipi_mode);
  c.phi1_mode = mode_of(ONOFF, s.phi1_mode); c.phi3_mode = mode_of(ONOFF, s.phi3_mode);
  c.modulate_zeta_with_tanh = s.modulate_zeta_with_tanh ? 1 : 0; c.critical_scaling_bulk = s.critical_scaling_bulk ? 1 : 0;
  c.constant_eta_over_s = s.constant_eta_over_s; c.constant_zeta_over_s = s.constant_zeta_over_s;
  c.cs2_dependent_zeta_A = s.cs2_dependent_zeta_A; c.cs2_dependent_zeta_p = s.cs2_dependent_zeta_p; c.C_B = s.C_B;
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) c.constant_kappa_over_T2[a * 3 + b] = s.constant_kappa_over_T2[a][b];
  c.critical_scaling_diffusion = s.critical_scaling_diffusion ? 1 : 0;
  c.critical_point_T = s.critical_point_T; c.critical_point_muB = s.critical_point_muB;
  c.critical_gaussian_width_T = s.critical_gaussian_width_T; c.critical_gaussian_width_muB = s.critical_gaussian_width_muB;
  c.diffusion_regulator_enabled = s.diffusion_regulator_enabled ? 1 : 0;
  c.diffusion_regulator_chi0 = s.diffusion_regulator_chi0; c.diffusion_regulator_e0 = s.diffusion_regulator_e0;
  c.diffusion_regulator_xi0 = s.diffusion_regulator_xi0; c.diffusion_regulator_rq_max = s.diffusion_regulator_rq_max;
  c.eos_type = s.eos_type == "table" ? CCB_EOS_TABLE : s.eos_type == "tanh_conformal" ? CCB_EOS_TANH_CONFORMAL
             : s.eos_type == "conformal_diagonal" ? CCB_EOS_CONFORMAL_DIAGONAL : CCB_EOS_CONFORMAL;
  // EquationOfState::set_up_chosen_EOSs, src/eos_initialization.cpp:216-231
  const double Nc = 3.0, Nf = 2.5;
  c.conformal_c = M_PI * M_PI * (2.0 * (Nc * Nc - 1.0) + (7.0 / 2.0) * Nc * Nf) / 90.0;
  c.conformal_T0 = c.conformal_muB0 = c.conformal_muS0 = c.conformal_muQ0 = 1.0;
  c.freeze_out_temperature_MeV = s.Freeze_Out_Temperature;
  c.prohibit_unstable_cs2 = s.prohibit_unstable_cs2 ? 1 : 0; c.prohibit_acausal_cs2 = s.prohibit_acausal_cs2 ? 1 : 0;
  c.restrict_mu_T_ratios = s.restrict_mu_T_ratios ? 1 : 0;
  c.device = device; c.rank = rank; c.nranks = nranks;
  c.particlization_enabled = s.particlization_enabled ? 1 : 0;
  c.dt = s.dt;
  c.online_inverter_enabled = s.online_inverter_enabled ? 1 : 0;
  c.slab_axis = 0;
  return c;
}

// One ccb_ctx shared by the SystemState / SPHWorkstation / Evolver / FreezeOut of a run. Errors of the C ABI surface as
// exceptions carrying ccb_last_error (the reference exit()s in the same places).
class Device {
 public:
  ccb_ctx* ctx = nullptr;
  Device() = default;
  Device(const Device&) = delete;
  Device& operator=(const Device&) = delete;
  ~Device() { if (ctx) ccb_destroy(ctx); }
  void create(const ccb_config& cfg) {
    if (ctx) { ccb_destroy(ctx); ctx = nullptr; }
    check(ccb_create(&ctx, &cfg));
  }
  void check(int rc) const {
    if (rc) throw std::runtime_error(std::string("ccake_b200 error ") + std::to_string(rc) + ": " + (ctx ? ccb_last_error(ctx) : "no context"));
  }
};

}  // namespace b200
}  // namespace ccake
#endif
