// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// key=value lines -> the reference's Settings (include/settings.h), standing in for Input::decode_settings (src/input.cpp:96-912;
// yaml-cpp is absent from this image). Shared by oracle/ref_driver.cpp (the reference build) and tests/cpp/dropin_bsqhydro.cpp
// (the reference's unmodified BSQHydro.cpp compiled against the drop-in headers), so both sides run on identical Settings.
#ifndef CCAKE_ORACLE_SETTINGS_KV_H
#define CCAKE_ORACLE_SETTINGS_KV_H

#include <cstdio>
#include <memory>
#include <sstream>
#include <string>

#include "settings.h"

namespace settings_kv {


inline bool to_bool(const std::string& v) { return v == "1" || v == "true" || v == "True"; }

// key=value lines -> Settings, named after the Settings members (include/settings.h).
inline std::shared_ptr<Settings> parse_settings(const char* kv) {
  auto s = std::make_shared<Settings>();
  // defaults the YAML reader would apply (include/defaults.h), restricted to the hot path
  s->IC_type = "ccake"; s->coordinate_system = "hyperbolic"; s->kernel_type = "cubic_spline";
  s->eos_type = "conformal"; s->online_inverter_enabled = true; s->particlization_enabled = false;
  s->Freeze_Out_Temperature = 150.0; s->etaMode = "constant"; s->shearRelaxMode = "default";
  s->zetaMode = "constant"; s->bulkRelaxMode = "default"; s->diffusionMode = "disabled";
  s->diffusionRelaxMode = "constant_over_T"; s->delta_pipi_mode = "default"; s->tau_pipi_mode = "default";
  s->lambda_piPi_mode = "default"; s->phi6_mode = "default"; s->phi7_mode = "default";
  s->delta_PiPi_mode = "default"; s->lambda_Pipi_mode = "default"; s->phi1_mode = "default"; s->phi3_mode = "default";
  s->source_type = "disabled"; s->jets_type = "disabled"; s->print_conservation_status = false;
  s->buffer_event = false; s->modulate_zeta_with_tanh = false; s->rk_order = 2;
  s->constant_eta_over_s = 0.08; s->constant_zeta_over_s = 0.005;
  s->cs2_dependent_zeta_A = 1.67552; s->cs2_dependent_zeta_p = 2.0;
  std::istringstream in(kv);
  std::string line;
  while (std::getline(in, line)) {
    const auto eq = line.find('=');
    if (eq == std::string::npos) continue;
    const std::string k = line.substr(0, eq), v = line.substr(eq + 1);
    if (k == "dim") s->dim = std::stoi(v);
    else if (k == "coordinate_system") s->coordinate_system = v;
    else if (k == "t0") s->t0 = std::stod(v);
    else if (k == "dt") s->dt = std::stod(v);
    else if (k == "hT") s->hT = std::stod(v);
    else if (k == "hEta") s->hEta = std::stod(v);
    else if (k == "rk_order") s->rk_order = std::stoi(v);
    else if (k == "factorized_kernel") s->factorized_kernel = to_bool(v);
    else if (k == "e_cutoff") s->e_cutoff = std::stod(v);
    else if (k == "input_as_entropy") s->input_as_entropy = to_bool(v);
    else if (k == "stepx") s->stepx = std::stod(v);
    else if (k == "stepy") s->stepy = std::stod(v);
    else if (k == "stepEta") s->stepEta = std::stod(v);
    else if (k == "xmin") s->xmin = std::stod(v);
    else if (k == "ymin") s->ymin = std::stod(v);
    else if (k == "etamin") s->etamin = std::stod(v);
    else if (k == "max_tau") s->max_tau = std::stod(v);
    else if (k == "eos_type") s->eos_type = v;
    else if (k == "eos_path") s->eos_path = v;
    else if (k == "normalize_by_T") s->normalize_by_T = to_bool(v);
    else if (k == "online_inverter_enabled") s->online_inverter_enabled = to_bool(v);
    else if (k == "particlization_enabled") s->particlization_enabled = to_bool(v);
    else if (k == "preinverted_eos_path") s->preinverted_eos_path = v;
    else if (k == "Freeze_Out_Temperature") s->Freeze_Out_Temperature = std::stod(v);
    else if (k == "baryon_charge_enabled") s->baryon_charge_enabled = to_bool(v);
    else if (k == "strange_charge_enabled") s->strange_charge_enabled = to_bool(v);
    else if (k == "electric_charge_enabled") s->electric_charge_enabled = to_bool(v);
    else if (k == "etaMode") s->etaMode = v;
    else if (k == "constant_eta_over_s") s->constant_eta_over_s = std::stod(v);
    else if (k == "shearRelaxMode") s->shearRelaxMode = v;
    else if (k == "use_vorticity") s->use_vorticity = to_bool(v);
    else if (k == "delta_pipi_mode") s->delta_pipi_mode = v;
    else if (k == "tau_pipi_mode") s->tau_pipi_mode = v;
    else if (k == "lambda_piPi_mode") s->lambda_piPi_mode = v;
    else if (k == "phi6_mode") s->phi6_mode = v;
    else if (k == "phi7_mode") s->phi7_mode = v;
    else if (k == "input_initial_shear") s->input_initial_shear = to_bool(v);
    else if (k == "zetaMode") s->zetaMode = v;
    else if (k == "constant_zeta_over_s") s->constant_zeta_over_s = std::stod(v);
    else if (k == "cs2_dependent_zeta_A") s->cs2_dependent_zeta_A = std::stod(v);
    else if (k == "cs2_dependent_zeta_p") s->cs2_dependent_zeta_p = std::stod(v);
    else if (k == "bulk_from_trace") s->bulk_from_trace = to_bool(v);
    else if (k == "bulkRelaxMode") s->bulkRelaxMode = v;
    else if (k == "modulate_zeta_with_tanh") s->modulate_zeta_with_tanh = to_bool(v);
    else if (k == "C_B") s->C_B = std::stod(v);
    else if (k == "critical_scaling_bulk") s->critical_scaling_bulk = to_bool(v);
    else if (k == "critical_scaling_diffusion") s->critical_scaling_diffusion = to_bool(v);
    else if (k == "critical_point_T") s->critical_point_T = std::stod(v);
    else if (k == "critical_point_muB") s->critical_point_muB = std::stod(v);
    else if (k == "critical_gaussian_width_T") s->critical_gaussian_width_T = std::stod(v);
    else if (k == "critical_gaussian_width_muB") s->critical_gaussian_width_muB = std::stod(v);
    else if (k == "diffusion_regulator_enabled") s->diffusion_regulator_enabled = to_bool(v);
    else if (k == "diffusion_regulator_chi0") s->diffusion_regulator_chi0 = std::stod(v);
    else if (k == "diffusion_regulator_e0") s->diffusion_regulator_e0 = std::stod(v);
    else if (k == "diffusion_regulator_xi0") s->diffusion_regulator_xi0 = std::stod(v);
    else if (k == "diffusion_regulator_rq_max") s->diffusion_regulator_rq_max = std::stod(v);
    else if (k == "delta_PiPi_mode") s->delta_PiPi_mode = v;
    else if (k == "lambda_Pipi_mode") s->lambda_Pipi_mode = v;
    else if (k == "phi1_mode") s->phi1_mode = v;
    else if (k == "phi3_mode") s->phi3_mode = v;
    else if (k == "diffusionMode") s->diffusionMode = v;
    else if (k == "diffusionRelaxMode") s->diffusionRelaxMode = v;
    else if (k == "IC_type") s->IC_type = v;
    else if (k == "IC_file") s->IC_file = v;
    else if (k == "freezein_input_file") s->freezein_input_file = v;
    else if (k == "freezein_tol") s->freezein_tol = std::stod(v);
    else if (k == "freezein_max_iter") s->freezein_max_iter = std::stoi(v);
    else if (k == "results_directory") s->results_directory = v;
    else if (k == "txt_evolution") s->txt_evolution = to_bool(v);
    else if (k == "evolution_stride") s->evolution_stride = std::stoi(v);
    else if (k == "get_neighbors") s->get_neighbors = to_bool(v);
    else if (k == "input_initial_diffusion") s->input_initial_diffusion = to_bool(v);
    else if (k == "print_conservation_status") s->print_conservation_status = to_bool(v);
    else if (k == "prohibit_unstable_cs2") s->prohibit_unstable_cs2 = to_bool(v);
    else if (k == "prohibit_acausal_cs2") s->prohibit_acausal_cs2 = to_bool(v);
    else if (k == "restrict_mu_T_ratios") s->restrict_mu_T_ratios = to_bool(v);
    else if (k.rfind("kappa_", 0) == 0 && k.size() == 8) s->constant_kappa_over_T2[k[6] - '0'][k[7] - '0'] = std::stod(v);
    else { std::fprintf(stderr, "ccake_ref: unknown setting '%s'\n", k.c_str()); return nullptr; }
  }
  // Derived switches exactly as Input::decode_settings sets them (src/input.cpp:390-408,
  // :419-421, :577-582). NB Settings::check_consistency() is never called by the reference.
  if (s->etaMode == "disabled") s->using_shear = false;
  else if (s->etaMode == "constant") s->using_shear = !(s->constant_eta_over_s <= 1.E-3);
  s->using_diffusion = (s->diffusionMode != "disabled");
  s->using_vorticity = s->use_vorticity;
  return s;
}

}  // namespace settings_kv
#endif
