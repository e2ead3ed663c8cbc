// Transport coefficients per particle (phase 08).
// Reference: src/transport_coefficients.cpp:284-1059 (functions) and :9-279
// (mode strings -> enums), called from SPHWorkstation::update_all_particle_viscosities
// (src/sph_workstation.cpp:1623-1699). The reference stores 13 scalars and twelve
// 3x3 matrices (nine of them identically zero) per particle; here they live in
// registers of the fused equations-of-motion kernel.
#pragma once

#include "physics_eos.cuh"

namespace ccb {

// Values mirror the reference enums (include/transport_coefficients.h:18-97).
enum { SHEAR_DEFAULT = 0, SHEAR_CONST, SHEAR_JAKI, SHEAR_LINEAR_MUS, SHEAR_INTERPOLATOR };
enum { TAU_PI_DEFAULT = 0, TAU_PI_MINVAL, TAU_PI_GUBSER };
enum { ZETA_DEFAULT = 0, ZETA_CONSTANT, ZETA_DNMR, ZETA_INTERPOLATE, ZETA_CS2_DEPENDENT };
enum { TAU_BULK_DEFAULT = 0, TAU_BULK_DNMR };
enum { KAPPA_DEFAULT = 0, KAPPA_DNMR, KAPPA_B_COTH };
enum { TAU_Q_DEFAULT = 0, TAU_Q_CB_OVER_T };
enum { MODE_DEFAULT = 0, MODE_IS = 1, MODE_DISABLED = 2 };  // delta_* : default / israel-stewart / disabled
enum { ONOFF_DEFAULT = 0, ONOFF_DISABLED = 1 };             // tau_pipi, lambda_*, phi*

struct TransportParams {
  int shear_mode, shear_relax_mode, zeta_mode, bulk_relax_mode, diffusion_mode, diffusion_relax_mode;
  int delta_pipi_mode, tau_pipi_mode, lambda_piPi_mode, phi6_mode, phi7_mode;
  int delta_PiPi_mode, lambda_Pipi_mode, phi1_mode, phi3_mode;
  int modulate_zeta_with_tanh, critical_scaling_bulk;
  double eta_over_s, zeta_over_s, zeta_A, zeta_p, C_B;
  double kappa_over_T2[3][3];
  // Gaussian suppression of kappa near the critical point (src/transport_coefficients.cpp:1047-1056), in fm^-1
  int critical_scaling_diffusion;
  double crit_T, crit_muB, crit_wT, crit_wmuB;
};

struct Transport {
  double eta_pi, tau_pi, zeta_Pi, tau_Pi, delta_PiPi, delta_pipi, lambda_piPi, lambda_Pipi, phi1, phi3, phi6, phi7, tau_pipi;
};

CCB_HD double tc_eta(const Thermo& th, const TransportParams& P) {
  switch (P.shear_mode) {
    case SHEAR_CONST: return (th.T <= 0.0) ? 0.0 : P.eta_over_s * (th.w / th.T);
    case SHEAR_JAKI: {
      const double TC = 155.0 / kHbarcMeVfm;
      const double z = pow(0.66 * th.T / TC, 2);
      const double alpha = 33. / (12. * kPi) * (z - 1) / (z * log(z));
      return th.s * (0.0416762 / pow(alpha, 1.6) + 0.0388977 / pow(th.T, 5.1));
    }
    case SHEAR_LINEAR_MUS:
      return (th.T <= 0.0) ? 0.0 : (P.eta_over_s + 0.0033 * (th.muB + th.muS + th.muQ)) * (th.w / th.T);
    case SHEAR_INTERPOLATOR: return 0.0;
    default: return 0.20 * (th.w / th.T);
  }
}
CCB_HD double tc_tau_pi(const Thermo& th, const TransportParams& P) {
  const double x = 5.0 * tc_eta(th, P) / th.w;
  switch (P.shear_relax_mode) {
    case TAU_PI_GUBSER: return x;
    case TAU_PI_MINVAL: return fmax(x, 0.001);
    default: return fmax(x, 0.005);
  }
}
CCB_HD double tc_zeta(const Thermo& th, const TransportParams& P) {
  double zeta;
  switch (P.zeta_mode) {
    case ZETA_CONSTANT: zeta = P.zeta_over_s * (th.w / th.T); break;
    case ZETA_DNMR: zeta = 0.0; break;
    case ZETA_CS2_DEPENDENT: {
      double factor = 1.0;
      if (P.modulate_zeta_with_tanh) {
        const double Ttr = 150.0 / kHbarcMeVfm, Tsc = 10.0 / kHbarcMeVfm;
        const double th_x = tanh((th.T - Ttr) / Tsc);
        const double x_p = pow(th.T / Ttr, P.zeta_p);
        factor = 0.5 * (1.0 + x_p) + 0.5 * (1.0 - x_p) * th_x;
      }
      const double zs = P.zeta_A * factor * pow((1.0 / 3.0) - fmin(th.cs2, 1.0), P.zeta_p);
      zeta = zs * th.w / th.T;
      break;
    }
    default: zeta = 0.005 * (th.w / th.T); break;
  }
  if (P.critical_scaling_bulk) {  // src/transport_coefficients.cpp:1027-1044
    const double xi = sqrt(th.chiBB), xi0 = 1.0176198446545879;
    if (!(fabs(xi0 - xi) < kTiny)) zeta *= (1. + pow(xi / xi0, 3));
  }
  return zeta;
}
CCB_HD double tc_tau_Pi(const Thermo& th, const TransportParams& P) {
  if (P.bulk_relax_mode == TAU_BULK_DNMR) return 0.0;
  const double ccs2 = th.cs2 > 1 ? 1 : th.cs2;
  const double f = pow((1.0 / 3.0 - ccs2), P.zeta_p);
  if (f < 1e-10) return 1e10;
  return fmax(5.0 * tc_zeta(th, P) / (f * th.w), 0.1);
}

CCB_HD void transport_coefficients(const Thermo& th, const TransportParams& P, Transport& tc) {
  tc.eta_pi = tc_eta(th, P);
  tc.tau_pi = tc_tau_pi(th, P);
  tc.zeta_Pi = tc_zeta(th, P);
  tc.tau_Pi = tc_tau_Pi(th, P);
  tc.delta_PiPi = P.delta_PiPi_mode == MODE_DISABLED ? 0.0 : (P.delta_PiPi_mode == MODE_IS ? tc.tau_Pi : 2. * tc.tau_Pi / 3.0);
  tc.delta_pipi = P.delta_pipi_mode == MODE_DISABLED ? 0.0 : (P.delta_pipi_mode == MODE_IS ? tc.tau_pi : 4. * tc.tau_pi / 3.0);
  tc.lambda_piPi = P.lambda_piPi_mode == ONOFF_DISABLED ? 0.0 : 6. / 5. * tc.tau_pi;
  {
    const double ccs2 = th.cs2 > 1 ? 1 : th.cs2;
    tc.lambda_Pipi = P.lambda_Pipi_mode == ONOFF_DISABLED ? 0.0 : 8. * (1. / 3. - ccs2) * tc.tau_Pi / 5.;
  }
  tc.phi1 = 0.0;
  tc.phi3 = 0.0;
  tc.phi6 = 0.0;
  tc.phi7 = P.phi7_mode == ONOFF_DISABLED ? 0.0 : 9. / (70. * th.p);
  tc.tau_pipi = P.tau_pipi_mode == ONOFF_DISABLED ? 0.0 : 10. * tc.tau_pi / 7.0;
}

// kappa_q and tau_q (3x3 over B,S,Q), src/transport_coefficients.cpp:580-623, :934-1022.
// is_table: the particle's thermodynamics came from the tabulated EoS (thermo_info::eos_type_is_table)
CCB_HD void diffusion_coefficients(const Thermo& th, const TransportParams& P, bool is_table, double kappa[3][3], double tauq[3][3]) {
  CCB_UNROLL for (int i = 0; i < 3; ++i) CCB_UNROLL for (int j = 0; j < 3; ++j) { kappa[i][j] = 0.0; tauq[i][j] = 0.0; }
  if (P.diffusion_mode == KAPPA_DNMR) return;
  if (P.diffusion_mode == KAPPA_B_COTH) {
    constexpr double floor_x = 1.0e-2;
    const double x = (th.T > kTiny) ? th.muB / th.T : 0.0;
    if (th.T > kTiny && fabs(th.w) > kTiny && fabs(x) > floor_x) {
      const double kB = (P.C_B / th.T) * th.rhoB * ((1.0 / 3.0) / tanh(x) - th.rhoB * th.T / th.w);
      if (kB == kB && fabs(kB) < 1e300 && kB > 0.0) kappa[0][0] = kB;  // isfinite(kB)
    }
    if (P.diffusion_relax_mode == TAU_Q_CB_OVER_T) {
      if (th.T > kTiny) CCB_UNROLL for (int i = 0; i < 3; ++i) tauq[i][i] = P.C_B / th.T;
    } else {
      CCB_UNROLL for (int i = 0; i < 3; ++i) tauq[i][i] += 0.2 / th.T;
    }
    return;
  }
  double suppression = 1.0;
  if (P.critical_scaling_diffusion && is_table) {
    const double dT = th.T - P.crit_T, dmuB = th.muB - P.crit_muB;
    const double exponent = 0.5 * (dT * dT / (P.crit_wT * P.crit_wT) + dmuB * dmuB / (P.crit_wmuB * P.crit_wmuB));
    suppression = 1.0 - exp(-exponent);
  }
  CCB_UNROLL for (int i = 0; i < 3; ++i) CCB_UNROLL for (int j = 0; j < 3; ++j) kappa[i][j] = suppression * P.kappa_over_T2[i][j] * (th.T * th.T);
  CCB_UNROLL for (int i = 0; i < 3; ++i) tauq[i][i] += 0.2 / th.T;
}

}  // namespace ccb
