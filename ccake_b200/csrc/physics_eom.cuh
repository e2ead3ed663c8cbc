// Israel-Stewart / DNMR equations of motion for one SPH particle in the
// reference's "M du/dt + R dN/dt = F" form, evaluated entirely in registers.
//
// Reference: EoM_default<D>::evaluate_time_derivatives (src/eom_default.cpp:212-869),
// calculate_MRF_shear (:883-1412), calculate_MRF_diffusion (:1416-1823) and the
// Cartesian twin src/eom_cartesian.cpp (same text with t^2 -> 1 in the metric and
// every 1/t, t*u_eta and delta_i_eta term removed). The reference
// runs these as 5-7 kernels that round-trip ~300 doubles of M/R/F scratch per
// particle through global memory; here one call produces the time derivatives.
//
// Index conventions kept from the reference (SURVEY.md section 7 "index traps"):
//   shv[4][4]: 0=tau 1=x 2=y 3=eta; charges 0=B 1=S 2=Q; gradV[i][j] = d_i v_j;
//   grad_alpha[idir][icharge]; ext_shv order (xx, xy, x-eta, yy, y-eta).
// D = 2 or 3 (D = 1 indexes out of bounds in the reference's own shear code,
// src/eom_default.cpp:1326-1329, and is not supported).
#pragma once

#include "physics_transport.cuh"

namespace ccb {

template <int D> struct EomIn {
  double u[D], v[D], gamma, sigma, sigma_lab, mass_s;
  double bulk;
  double shv[4][4];
  double q[3][4];
  double gradP[D], gradBulk[D], gradshear[D], divshear[D], gradV[D][D];
  double div_qa[3], grad_qa[3], grad_alpha[D][3];
};

template <int D> struct EomOut {
  double du_dt[D];
  double ds_dt;      // d(extensive s)/dt
  double dN_dt[3];   // d(extensive rho_B,S,Q)/dt
  double dbulk_dt;   // d(extensive bulk)/dt
  double dshv_dt[5]; // d(extensive shv)/dt in (xx, xy, x-eta, yy, y-eta) order
  double dq_dt[3][3];// d(extensive diffusion)/dt [icharge][idir]
  double theta, shv_nabla_u, dEz;
};

struct EomFlags { bool shear, diffusion, vorticity; };

// The evaluation is split in two passes so that the kernel never holds the M/R/F tensors of every
// sector at once (the reference's ~300 doubles of scratch per particle):
//   PASS 1: everything the linear solve needs -> out.du_dt, out.dN_dt;
//   PASS 2: (reads out.du_dt, out.dN_dt) every other time derivative, each M/R/F block contracted with
//           du/dt and dN/dt as soon as it is formed.
// Shared sub-expressions are simply evaluated in both passes; every formula keeps the reference's
// expression order, so the split does not change a single rounding.
template <int D, bool MILNE, int PASS>
CCB_HD void eom_pass(const Metric<MILNE>& m, const EomIn<D>& in, const Thermo& th, const Transport& tc,
                     const double kappa_q[3][3], const double tau_q[3][3], const EomFlags fl, EomOut<D>& out) {
  const double t = m.t;
  const double gamma = in.gamma, sigma = in.sigma, bulk = in.bulk;
  // Reciprocals are formed once and multiplied: the reference divides ~750 times per particle (an fp64 division is
  // ~35 instructions on the GPU); the results differ from the divided form by at most 1 ulp per operation.
  const double ig = 1.0 / gamma, it_ = MILNE ? 1.0 / t : 0.0, third = 1.0 / 3.0, iT = 1.0 / th.T, i_g33 = 1.0 / m.g33;
  const double i_gs = 1.0 / (gamma * sigma), i_sgT = 1.0 / (sigma * gamma * th.T), i_sgtP = 1.0 / (sigma * gamma * tc.tau_Pi);
  const double i_tpg = 1.0 / (tc.tau_pi * gamma), i_den = 1.0 / (sigma * tc.tau_pi * gamma);
#define CCB_OT(x) (MILNE ? (x) * it_ : 0.0)  /* terms carrying 1/t   */
#define CCB_TT(x) (MILNE ? (x) * t : 0.0)  /* terms carrying t     */
  constexpr double die_last = (D == 3) ? 1.0 : 0.0;  // delta_i_eta (reference milne.hpp:1093-1108)
  double die[D];
  CCB_UNROLL for (int i = 0; i < D; ++i) die[i] = (i == D - 1) ? die_last : 0.0;
  const double full_die[3] = {0., 0., 1.};
  const double* u = in.u;
  const double* v = in.v;
  double ucov[D], fu[3] = {0., 0., 0.}, fucov[3];
  lower_vec<D, MILNE>(m, u, ucov);
  CCB_UNROLL for (int i = 0; i < D; ++i) fu[i] = u[i];
  lower_vec3<MILNE>(m, fu, fucov);
  const double gf = (MILNE && D == 3) ? die_last * u[D - 1] * u[D - 1] * t * ig : 0.0;  // geometric_factor
  double cmd[4];
  contra_metric_diag<MILNE>(m, cmd);

  double divV = 0.;
  CCB_UNROLL for (int i = 0; i < D; ++i) divV += in.gradV[i][i];
  const double theta_geo = MILNE ? (gamma * divV + gamma * it_ + gf) : gamma * divV;

  // pi^mu_nu and pi_{mu nu} (reference milne.hpp:549-591)
  double shv[4][4], shv_h[4][4], shv_c[4][4];
  CCB_UNROLL for (int a = 0; a < 4; ++a) CCB_UNROLL for (int b = 0; b < 4; ++b) {
    shv[a][b] = in.shv[a][b];
    shv_h[a][b] = (b == 0) ? shv[a][b] : ((b == 3) ? shv[a][b] * (-m.g33) : -shv[a][b]);
  }
  CCB_UNROLL for (int a = 0; a < 4; ++a) CCB_UNROLL for (int b = 0; b < 4; ++b)
    shv_c[a][b] = (a == 0) ? shv_h[a][b] : ((a == 3) ? shv_h[a][b] * (-m.g33) : -shv_h[a][b]);

  // ------------------------------------------------------------------ shear M/R/F
  double M_shv_nabla_u[D], F_shv_nabla_u = 0.0, F_0i_shear[D], M_0i_shear[D][D];
  CCB_UNROLL for (int i = 0; i < D; ++i) {
    M_shv_nabla_u[i] = 0.; F_0i_shear[i] = 0.;
    CCB_UNROLL for (int j = 0; j < D; ++j) M_0i_shear[i][j] = 0.;
  }
  if (PASS == 2) { CCB_UNROLL for (int n = 0; n < 5; ++n) out.dshv_dt[n] = 0.0; }
  if (fl.shear) {
    const double use_vort = fl.vorticity ? 1. : 0.;
    const double tilde_delta = tc.delta_pipi - tc.tau_pi;
    double grad_uj[D][D], grad_u0[D], grad_u0_contra[D], cgu3[3][3], gu3[3][3];
    CCB_UNROLL for (int i = 0; i < D; ++i) CCB_UNROLL for (int j = 0; j < D; ++j) {
      double g = gamma * in.gradV[i][j];
      CCB_UNROLL for (int k = 0; k < D; ++k) g += -gamma * u[j] * ucov[k] * in.gradV[i][k];
      grad_uj[i][j] = g;
    }
    CCB_UNROLL for (int i = 0; i < D; ++i) {
      double c = 0.;
      CCB_UNROLL for (int j = 0; j < D; ++j) c += grad_uj[i][j] * ucov[j];
      grad_u0[i] = -1. * c * ig;
    }
    raise_vec<D, MILNE>(m, grad_u0, grad_u0_contra);
    CCB_UNROLL for (int i = 0; i < 3; ++i) CCB_UNROLL for (int j = 0; j < 3; ++j) { cgu3[i][j] = 0.; gu3[i][j] = 0.; }
    CCB_UNROLL for (int i = 0; i < D; ++i) CCB_UNROLL for (int j = 0; j < D; ++j) {
      gu3[i][j] = grad_uj[i][j];
      // first index raised: D==2 flips every entry (milne.hpp:488-496), D==3 scales the last row by -1/t^2
      cgu3[i][j] = (D == 2) ? -grad_uj[i][j] : ((i == D - 1) ? grad_uj[i][j] * (-1. * i_g33) : -grad_uj[i][j]);
    }
    double gu_v[D];  // contract(grad_uj, v, FirstIndex)
    CCB_UNROLL for (int j = 0; j < D; ++j) { double c = 0.; CCB_UNROLL for (int k = 0; k < D; ++k) c += grad_uj[k][j] * v[k]; gu_v[j] = c; }
    double gu0_v = 0.;
    CCB_UNROLL for (int i = 0; i < D; ++i) gu0_v += grad_u0[i] * v[i];

    double shv_grad_u = 0.0;
    CCB_UNROLL for (int i = 0; i < D; ++i) {
      CCB_UNROLL for (int j = 0; j < D; ++j) shv_grad_u += (shv_h[i + 1][j + 1] - v[i] * shv_h[0][j + 1]) * grad_uj[i][j];
      shv_grad_u += (shv_h[i + 1][0] - v[i] * shv_h[0][0]) * grad_u0[i];
    }
    double shv_shvcov = 0.0;
    CCB_UNROLL for (int a = 0; a < 4; ++a) CCB_UNROLL for (int b = 0; b < 4; ++b) shv_shvcov += shv[a][b] * shv_c[a][b];

    const double eta_eff = 2. * tc.eta_pi + tc.lambda_piPi * bulk;
    const double ul = u[D - 1];
    // -- pi^{ij} nabla_i u_j = M du/dt + F (reference :1171-1184)
    CCB_UNROLL for (int i = 0; i < D; ++i) {
      M_shv_nabla_u[i] = shv_h[0][i + 1] - shv_h[0][0] * ucov[i] * ig;
      CCB_UNROLL for (int j = 0; j < D; ++j) F_shv_nabla_u += (shv_h[i + 1][j + 1] - v[i] * shv_h[0][j + 1]) * grad_uj[i][j];
      F_shv_nabla_u += (shv_h[i + 1][0] - v[i] * shv_h[0][0]) * grad_u0[i];
    }
    if (MILNE)
      F_shv_nabla_u += +shv_h[3][3] * gamma * it_ + shv_h[0][0] * gf + (t * shv_h[3][0] + shv_h[0][3] * it_) * ul * die[D - 1];
    if (PASS == 1) {
      double v_gradu_shv[D], grad_u_g_shv_i[D], shv_grad_u_i[D], v_shv_grad_u[D], shvil_shv0l[D];
      CCB_UNROLL for (int i = 0; i < D; ++i) {
        double a = 0., b = 0., c = 0., d = 0., e = 0.;
        CCB_UNROLL for (int j = 0; j < D; ++j) {
          a += gu_v[j] * shv_h[i + 1][j + 1];
          b += cgu3[i][j] * shv_h[0][j + 1];
          c += shv[j + 1][0] * grad_uj[j][i] + shv[j + 1][i + 1] * grad_u0[j];
          d += v[j] * shv[0][0] * grad_uj[j][i] + v[j] * shv[0][i + 1] * grad_u0[j];
        }
        a += gu0_v * shv_h[i + 1][0];
        b += cmd[i + 1] * grad_u0[i] * shv_h[0][0];
        CCB_UNROLL for (int j = 0; j < 4; ++j) e += (shv[0][j] * shv_h[i + 1][j] + shv[i + 1][j] * shv_h[0][j]) * 0.5;
        v_gradu_shv[i] = a; grad_u_g_shv_i[i] = b; shv_grad_u_i[i] = c; v_shv_grad_u[i] = d; shvil_shv0l[i] = e;
      }
      // -- four-acceleration pieces (reference :1101-1169)
      double F_i0_sigma[D], F_i0_D[D], F_i0_domega[D], F_i0_dsigma[D], F_i0_dd[D];
      double M_i0_sigma[D][D], M_i0_D[D][D], M_i0_domega[D][D], M_i0_dsigma[D][D], M_i0_dd[D][D];
      CCB_UNROLL for (int i = 0; i < D; ++i) {
        F_i0_sigma[i] = -(gu_v[i] - grad_u0_contra[i]) * 0.5 + u[i] * gamma * theta_geo * third - gamma * gf * u[i] * 0.5;
        F_i0_D[i] = gamma * (u[i] * shv_h[0][0] + gamma * shv_h[i + 1][0]) * gf + CCB_TT(ul * shv[3][i + 1] * die[i]);
        F_i0_dsigma[i] =
            -(v_gradu_shv[i] - grad_u_g_shv_i[i] - 4. * gamma * u[i] * shv_grad_u * third - shv_grad_u_i[i] + v_shv_grad_u[i]) * 0.25 +
            (shv_h[i + 1][0] - gamma * gamma * shv_h[i + 1][0] - gamma * u[i] * shv_h[0][0] + shv[0][i + 1] +
             4. * u[i] * gamma * shv_h[0][0] * third) * gf * 0.25 +
            (CCB_OT(gamma * cmd[i + 1] * shv_h[0][3] * die[i]) + CCB_OT(gamma * shv[3][0] * die[i]) +
             4. * gamma * u[i] * CCB_OT(gamma * shv_h[3][3]) * third) * 0.25 -
            shv[0][i + 1] * theta_geo * third;
        F_i0_domega[i] =
            -(v_gradu_shv[i] - grad_u_g_shv_i[i] + shv_grad_u_i[i] - v_shv_grad_u[i]) * 0.25 +
            (shv_h[i + 1][0] - gamma * gamma * shv_h[i + 1][0] - gamma * u[i] * shv_h[0][0] - shv[0][i + 1]) * gf * 0.25 +
            (CCB_OT(gamma * cmd[i + 1] * shv_h[0][3] * die[i]) - CCB_OT(gamma * shv[3][0] * die[i])) * 0.25;
        F_i0_dd[i] = shv[i + 1][0] * theta_geo;
        CCB_UNROLL for (int j = 0; j < D; ++j) {
          M_i0_sigma[i][j] = u[i] * ucov[j] * (0.5 * third);
          M_i0_D[i][j] = gamma * u[i] * (shv_h[0][j + 1] - shv_h[0][0] * ucov[j] * ig) +
                         gamma * gamma * (shv_h[i + 1][j + 1] - shv_h[i + 1][0] * ucov[j] * ig);
          M_i0_dsigma[i][j] = (shv_h[i + 1][j + 1] - shv_h[i + 1][0] * ucov[j] * ig - gamma * gamma * shv_h[i + 1][j + 1] -
                               gamma * u[i] * shv_h[0][j + 1] + ucov[j] * gamma * shv_h[i + 1][0] + ucov[j] * u[i] * shv_h[0][0] +
                               4. * gamma * u[i] * (shv_h[0][j + 1] - shv_h[0][0] * ucov[j] * ig) * third +
                               (-shv[0][i + 1] + 4. * shv[0][i + 1] * third) * ucov[j] * ig) * 0.25;
          M_i0_domega[i][j] = (shv_h[i + 1][j + 1] - shv_h[i + 1][0] * ucov[j] * ig - gamma * gamma * shv_h[i + 1][j + 1] -
                               gamma * u[i] * shv_h[0][j + 1] + ucov[j] * gamma * shv_h[i + 1][0] + ucov[j] * u[i] * shv_h[0][0] +
                               shv[0][i + 1] * ucov[j] * ig) * 0.25;
          M_i0_dd[i][j] = -shv[i + 1][0] * ucov[j] * ig;
        }
        M_i0_sigma[i][i] += (1. - gamma * gamma) * 0.5;
        M_i0_domega[i][i] += -shv[0][0] * 0.25;
        M_i0_dsigma[i][i] += +shv[0][0] * 0.25;
        if (MILNE) {
          F_i0_sigma[i] += die[i] * (-(-ul * t - ul * it_ + gamma * ul * ul * t + 2. * gamma * gamma * ul * it_) * 0.5);
          F_i0_D[i] += die[i] * (ul * shv[0][0] * it_ + gamma * shv[3][0] * it_);
          if (D != 2) {
            F_i0_dsigma[i] += (t * ul * die[i] * cmd[i + 1] * shv_h[0][0] + ul * shv_h[i + 1][3] * gamma * it_ * 0.25 +
                               shv_h[i + 1][3] * t * ul + shv[0][0] * die[i] * ul * it_ +
                               t * ul * ul * (gamma * shv_h[i + 1][0] + u[i] * shv_h[0][0]) +
                               2. * gamma * ul * (gamma * shv_h[i + 1][3] + u[i] * shv[0][3]) * it_) * 0.25 +
                              gamma * u[i] * (shv_h[3][0] * ul * t + ul * shv_h[0][3] * it_) * third;
            F_i0_domega[i] += (t * ul * die[i] * cmd[i + 1] * shv_h[0][0] + ul * shv_h[i + 1][3] * gamma * it_ -
                               shv_h[3][i + 1] * t * ul - shv[0][0] * die[i] * ul * it_ +
                               t * ul * ul * (gamma * shv_h[i + 1][0] + u[i] * shv_h[0][0]) +
                               2. * gamma * ul * (gamma * shv_h[i + 1][3] + u[i] * shv[0][3]) * it_) * 0.25;
          }
        }
      }
      // -- shear pieces of the acceleration equation (reference :1186-1213)
      CCB_UNROLL for (int i = 0; i < D; ++i) {
        F_0i_shear[i] = (-shv[i + 1][0] - tc.tau_pi * F_i0_D[i] - tc.delta_pipi * F_i0_dd[i] +
                         2. * tc.tau_pi * F_i0_domega[i] * use_vort - tc.tau_pipi * F_i0_dsigma[i] + eta_eff * F_i0_sigma[i] +
                         tc.phi6 * bulk * shv[i + 1][0] + tc.phi7 * (shvil_shv0l[i] + gamma * u[i] * shv_shvcov * third)) * i_tpg;
        CCB_UNROLL for (int j = 0; j < D; ++j)
          M_0i_shear[i][j] = (-tc.tau_pi * M_i0_D[i][j] - tc.delta_pipi * M_i0_dd[i][j] + 2. * tc.tau_pi * M_i0_domega[i][j] * use_vort -
                              tc.tau_pipi * M_i0_dsigma[i][j] + eta_eff * M_i0_sigma[i][j]) * i_tpg;
      }

    }
    // -- evolution of the independent pi^{ij} components (reference :1216-1409): PASS 2, one (i, j) at a time,
    //    contracted with du/dt at once
    if (PASS == 2) {
      const double* du = out.du_dt;
      const double ueta = fu[2];
      int n = 0;
      CCB_UNROLL for (int i = 0; i < 2; ++i) {
        CCB_UNROLL for (int j = i; j < 3; ++j) {
          double gug = 0., sgu = 0., vsgu = 0., sls = 0.;
          CCB_UNROLL for (int k = 0; k < D; ++k) {
            gug += cgu3[i][k] * shv_h[j + 1][k + 1] + cgu3[j][k] * shv_h[i + 1][k + 1];
            sgu += shv[k + 1][i + 1] * gu3[k][j] + shv[k + 1][j + 1] * gu3[k][i];
            vsgu += v[k] * shv[0][i + 1] * gu3[k][j] + v[k] * shv[0][j + 1] * gu3[k][i];
          }
          CCB_UNROLL for (int k = 0; k < 4; ++k) sls += (shv[i + 1][k] * shv_h[j + 1][k] + shv[j + 1][k] * shv_h[i + 1][k]) * 0.5;
          const double die_i = (i < D) ? die[i < D ? i : 0] : 0.0;  // i < 2 <= D always
          const double die_j = (j < D) ? die[j < D ? j : 0] : 0.0;  // delta_i_eta(jdir) is out of range for D == 2, j == 2 in the reference
          double F_D = gamma * gf * (fu[i] * shv_h[j + 1][0] + fu[j] * shv_h[i + 1][0]);
          double F_sg = cgu3[i][j] * 0.5 + cgu3[j][i] * 0.5 + fu[i] * fu[j] * theta_geo * third;
          const double F_delta = shv[i + 1][j + 1] * theta_geo;
          // common metric block of F_domega / F_dsigma (all terms carry t or 1/t)
          double geo_a = 0., geo_w = 0., geo_s = 0.;
          if (MILNE) {
            geo_a = t * fu[2] * (cmd[i + 1] * shv_h[j + 1][0] * die_i + cmd[j + 1] * shv_h[i + 1][0] * die_j) +
                    gamma * (cmd[i + 1] * shv_h[j + 1][3] * die_i + cmd[j + 1] * shv_h[i + 1][3] * die_j) * it_ +
                    fu[2] * fu[2] * t * (fu[i] * shv_h[j + 1][0] + fu[j] * shv_h[i + 1][0]) +
                    2. * gamma * fu[2] * (fu[i] * shv_h[j + 1][3] + fu[j] * shv_h[i + 1][3]) * it_;
            geo_w = -gamma * (shv[3][i + 1] * die_j + shv[3][j + 1] * die_i) * it_ - fu[2] * (shv[0][i + 1] * die_j + shv[0][j + 1] * die_i) * it_;
            geo_s = +gamma * (shv[3][i + 1] * die_j + shv[3][j + 1] * die_i) * it_ + fu[2] * (shv[0][i + 1] * die_j + shv[0][j + 1] * die_i) * it_ +
                    4. * fu[i] * fu[j] * (shv_h[3][0] * fu[2] * t + gamma * shv_h[3][3] * it_ + fu[2] * shv_h[0][3] * it_) * third;
          }
          const double F_domega = -(-gug + sgu - vsgu) * 0.25 +
                                  (-fu[i] * gamma * shv_h[j + 1][0] - fu[j] * gamma * shv_h[i + 1][0]) * gf * 0.25 + (geo_a + geo_w) * 0.25;
          double F_dsigma = -(-gug - sgu + vsgu - 4. * (fu[i] * fu[j]) * shv_grad_u * third) * 0.25 +
                            (-fu[i] * gamma * shv_h[j + 1][0] - fu[j] * gamma * shv_h[i + 1][0] + 4. * fu[i] * fu[j] * shv_h[0][0] * third) * gf * 0.25 +
                            (geo_a + geo_s) * 0.25 - shv[i + 1][j + 1] * theta_geo * third;
          // diagonal and metric terms for F
          if (MILNE && j == 2) {
            F_D += gamma * shv[3][i + 1] * it_ + ueta * shv[0][i + 1] * it_;
            F_sg += -fu[i] * ueta * ueta * t * 0.5 - fu[i] * ueta * gamma * it_;
          }
          if (i == j) {
            F_sg += -cmd[i + 1] * theta_geo * third;
            F_dsigma += -cmd[i + 1] * shv_grad_u * third;
            if (MILNE)
              F_dsigma += -cmd[i + 1] * shv_h[0][0] * gf * third -
                          cmd[i + 1] * (shv_h[3][0] * fu[2] * t + gamma * shv_h[3][3] * it_ + fu[2] * shv_h[0][3] * it_) * third;
          }
          double F_ext = (-shv[i + 1][j + 1] - tilde_delta * F_delta - tc.tau_pi * F_D + 2. * tc.tau_pi * F_domega * use_vort -
                          tc.tau_pipi * F_dsigma + eta_eff * F_sg + tc.phi6 * bulk * shv[i + 1][j + 1] +
                          tc.phi7 * (sls + fu[i] * fu[j] * shv_shvcov * third)) * i_den;
          if (i == j) F_ext += tc.phi7 * (-cmd[i + 1] * shv_shvcov) * (third * i_den);
          double acc = 0.;
          CCB_UNROLL for (int k = 0; k < D; ++k) {
            const double M_D = gamma * (fu[i] * shv_h[j + 1][k + 1] - fu[i] * shv_h[j + 1][0] * fucov[k] * ig +
                                        fu[j] * shv_h[i + 1][k + 1] - fu[j] * shv_h[i + 1][0] * fucov[k] * ig);
            double M_sg = -fu[i] * fu[j] * fucov[k] * (third * ig);
            const double M_delta = -shv[i + 1][j + 1] * fucov[k] * ig;
            double M_domega = (-fu[i] * shv_h[j + 1][k + 1] * gamma - fu[j] * shv_h[i + 1][k + 1] * gamma +
                               fucov[k] * fu[i] * shv_h[j + 1][0] + fucov[k] * fu[j] * shv_h[i + 1][0]) * 0.25;
            double M_dsigma = (-fu[i] * shv_h[j + 1][k + 1] * gamma - fu[j] * shv_h[i + 1][k + 1] * gamma +
                               fucov[k] * fu[i] * shv_h[j + 1][0] + fucov[k] * fu[j] * shv_h[i + 1][0] +
                               4. * fu[i] * fu[j] * (shv_h[0][k + 1] - shv_h[0][0] * fucov[k] * ig) * third +
                               4. * shv[i + 1][j + 1] * fucov[k] * third * ig) * 0.25;
            if (k == i) {  // diagonal i = k terms
              M_dsigma += shv[0][j + 1] * 0.25;
              M_domega += -shv[0][j + 1] * 0.25;
              M_sg += -gamma * fu[j] * 0.5;
            }
            if (k == j) {  // diagonal j = k terms (j < D)
              M_sg += -gamma * fu[i] * 0.5;
              M_dsigma += shv[0][i + 1] * 0.25;
              M_domega += -shv[0][i + 1] * 0.25;
            }
            if (i == j) {  // diagonal i = j terms
              M_sg += cmd[i + 1] * fucov[k] * (third * ig);
              M_dsigma += -cmd[i + 1] * (shv_h[0][k + 1] - shv_h[0][0] * fucov[k] * ig) * third;
            }
            const double M_ext = (-tc.tau_pi * M_D - tilde_delta * M_delta + 2. * tc.tau_pi * M_domega * use_vort -
                                  tc.tau_pipi * M_dsigma + eta_eff * M_sg) * i_den;
            acc += M_ext * du[k];
          }
          acc += F_ext;
          out.dshv_dt[n] = acc;
          ++n;
        }
      }
    }
  }

  // ------------------------------------------------------------------ bulk / entropy (reference :230-363)
  const double zeta_tilde = tc.zeta_Pi + (tc.delta_PiPi - tc.tau_Pi) * bulk;
  const double mu_vec[3] = {th.muB, th.muS, th.muQ};
  double M_ext_bulk[D], M_ext_entropy[D], R_ext_entropy[3], R_ext_bulk[3];
  CCB_UNROLL for (int i = 0; i < D; ++i) {
    M_ext_bulk[i] = (zeta_tilde * ucov[i] * ig) * i_sgtP;
    M_ext_entropy[i] = (bulk * ucov[i] * ig) * i_sgT;
  }
  CCB_UNROLL for (int c = 0; c < 3; ++c) { R_ext_entropy[c] = -mu_vec[c] * iT; R_ext_bulk[c] = 0.0; }
  double F_ext_bulk = -(zeta_tilde * theta_geo + bulk - tc.phi1 * bulk * bulk) * i_sgtP;
  double F_ext_entropy = (-bulk * theta_geo) * i_sgT;  // external sources j0_ext, j_ext are zero (out of scope)
  if (fl.shear) {
    double shvc_shv = 0.;
    CCB_UNROLL for (int b = 0; b < 4; ++b) CCB_UNROLL for (int a = 0; a < 4; ++a) shvc_shv += shv_c[a][b] * shv[a][b];
    F_ext_entropy += F_shv_nabla_u * i_sgT;
    F_ext_bulk += -(tc.lambda_Pipi * F_shv_nabla_u - tc.phi3 * shvc_shv) * i_sgtP;
    CCB_UNROLL for (int i = 0; i < D; ++i) {
      M_ext_bulk[i] += -tc.lambda_Pipi * M_shv_nabla_u[i] * i_sgtP;
      M_ext_entropy[i] += M_shv_nabla_u[i] * i_sgT;
    }
  }

  // ------------------------------------------------------------------ diffusion M/R/F (reference :1416-1823)
  double F_q0[3] = {0., 0., 0.}, R_q0[3][3], M_q0[D][3];
  CCB_UNROLL for (int a = 0; a < 3; ++a) CCB_UNROLL for (int b = 0; b < 3; ++b) R_q0[a][b] = 0.;
  CCB_UNROLL for (int i = 0; i < D; ++i) CCB_UNROLL for (int c = 0; c < 3; ++c) M_q0[i][c] = 0.;
  if (PASS == 2) { CCB_UNROLL for (int c = 0; c < 3; ++c) CCB_UNROLL for (int i = 0; i < 3; ++i) out.dq_dt[c][i] = 0.0; }
  if (fl.diffusion) {
    const double F_S = F_ext_entropy;
    const double ext_rho[3] = {th.rhoB, th.rhoS, th.rhoQ};
    const double ext_entropy = th.s;
    double M_S[3] = {0., 0., 0.}, ga[3][3];
    CCB_UNROLL for (int i = 0; i < 3; ++i) CCB_UNROLL for (int c = 0; c < 3; ++c) ga[i][c] = 0.;
    CCB_UNROLL for (int i = 0; i < D; ++i) {
      M_S[i] = M_ext_entropy[i];
      CCB_UNROLL for (int c = 0; c < 3; ++c) ga[i][c] = in.grad_alpha[i][c];
    }
    const double* R_S = R_ext_entropy;
    double dads[3], dadr[3][3];
    CCB_UNROLL for (int a = 0; a < 3; ++a) { dads[a] = th.dalpha[a][0]; CCB_UNROLL for (int b = 0; b < 3; ++b) dadr[a][b] = th.dalpha[a][b + 1]; }
    double q[3][4], qc[3][4];
    CCB_UNROLL for (int c = 0; c < 3; ++c) {
      CCB_UNROLL for (int mu = 0; mu < 4; ++mu) q[c][mu] = in.q[c][mu];
      qc[c][0] = q[c][0]; qc[c][1] = -q[c][1]; qc[c][2] = -q[c][2]; qc[c][3] = q[c][3] * (-m.g33);
    }
    double itq[3][3], itt[3][3], itk[3][3];
    inverse3(tau_q, itq);
    CCB_UNROLL for (int a = 0; a < 3; ++a) CCB_UNROLL for (int b = 0; b < 3; ++b) {
      double x = 0., y = 0.;
      CCB_UNROLL for (int c = 0; c < 3; ++c) { x += itq[a][c] * tau_q[c][b]; y += itq[a][c] * kappa_q[c][b]; }
      itt[a][b] = x; itk[a][b] = y;
    }
    if (PASS == 1) {
      const double g2m1 = (1. - gamma * gamma);
      double F_0_D[3], F_0_al[3], A_0_al[3], R_0_al[3][3], M_0_al[3][D], M_0_D[3][D];
      CCB_UNROLL for (int c = 0; c < 3; ++c) {
        F_0_D[c] = MILNE ? (gamma * gamma * gf * q[c][0] + 2. * gamma * gamma * qc[c][3] * fu[2] * it_ +
                            t * gamma * qc[c][0] * fu[2] * fu[2] + t * fu[2] * q[c][3])
                         : 0.0;
        F_0_al[c] = g2m1 * sigma * dads[c] * F_S;
        A_0_al[c] = -g2m1 * dads[c] * ext_entropy * ig;
        CCB_UNROLL for (int i = 0; i < D; ++i) {
          M_0_D[c][i] = gamma * gamma * (qc[c][i + 1] - qc[c][0] * ucov[i] * ig);
          M_0_al[c][i] = g2m1 * sigma * dads[c] * M_S[i];
          F_0_al[c] += -ga[i][c] * u[i] * ig;
        }
        CCB_UNROLL for (int b = 0; b < 3; ++b) {
          R_0_al[c][b] = g2m1 * sigma * (dads[c] * R_S[b] + dadr[c][b]);
          A_0_al[c] += -g2m1 * dadr[c][b] * ext_rho[b] * ig;
        }
      }
      CCB_UNROLL for (int c = 0; c < 3; ++c) {
        CCB_UNROLL for (int i = 0; i < D; ++i) {
          double x = 0.;
          CCB_UNROLL for (int b = 0; b < 3; ++b)
            x += (-itt[c][b] * M_0_D[b][i] + itk[c][b] * M_0_al[b][i] - itk[c][b] * A_0_al[b] * ucov[i] * ig) * ig +
                 itt[c][b] * q[b][0] * ucov[i] * ig * ig;
          M_q0[i][c] = x;
        }
        CCB_UNROLL for (int b = 0; b < 3; ++b) {
          F_q0[c] += (-itt[c][b] * F_0_D[b] - itq[c][b] * q[b][0] + itk[c][b] * F_0_al[b] + itk[c][b] * A_0_al[b] * theta_geo) * ig -
                     itt[c][b] * q[b][0] * theta_geo * ig;
          double x = 0.;
          CCB_UNROLL for (int k = 0; k < 3; ++k) x += itk[c][k] * R_0_al[k][b] * ig;
          R_q0[c][b] = x;
        }
      }
    }
    if (PASS == 2) {
      double F_D_ib[3][3], F_al_ib[3][3], A_al_ib[3][3];
      CCB_UNROLL for (int i = 0; i < 3; ++i) CCB_UNROLL for (int c = 0; c < 3; ++c) {
        // NB the reference terminates this statement after its second term (stray ';',
        // src/eom_default.cpp:1743-1745): the t*u*q_0*u_eta^2 and delta_i_eta terms are dropped.
        F_D_ib[i][c] = MILNE ? (fu[i] * gamma * q[c][0] * gf + 2. * fu[i] * qc[c][3] * fu[2] * gamma * it_) : 0.0;
        F_al_ib[i][c] = -fu[i] * gamma * sigma * dads[c] * F_S + cmd[i + 1] * ga[i][c];
        double A = fu[i] * dads[c] * ext_entropy;
        CCB_UNROLL for (int b = 0; b < 3; ++b) A += fu[i] * dadr[c][b] * ext_rho[b];
        A_al_ib[i][c] = A;
      }
      (void)full_die;
      const double* du = out.du_dt;
      CCB_UNROLL for (int i = 0; i < 3; ++i) CCB_UNROLL for (int c = 0; c < 3; ++c) {
        double f = 0.;
        double rdn[3];
        CCB_UNROLL for (int b = 0; b < 3; ++b) {
          f += (-itt[c][b] * F_D_ib[i][b] - itq[c][b] * q[b][i + 1] + itk[c][b] * F_al_ib[i][b] + itk[c][b] * A_al_ib[i][b] * theta_geo) * i_gs;
          double r = 0.;
          CCB_UNROLL for (int k = 0; k < 3; ++k)
            r += itk[c][k] * (-fu[i] * gamma * sigma * (dads[k] * R_S[b] + dadr[k][b])) * i_gs;
          rdn[b] = r;  // R_extensive_diffusion[i][c][b]
        }
        double x = f;  // F_extensive_diffusion[i][c]
        CCB_UNROLL for (int j = 0; j < 3; ++j) {
          double mq = 0.;
          CCB_UNROLL for (int b = 0; b < 3; ++b) {
            const double M_D_ibj = fu[i] * gamma * (qc[b][j + 1] - qc[b][0] * fucov[j] * ig);
            const double M_al_ibj = -fu[i] * gamma * sigma * dads[b] * M_S[j];
            mq += (-itt[c][b] * M_D_ibj + itk[c][b] * M_al_ibj - itk[c][b] * A_al_ib[i][b] * fucov[j] * ig) * i_gs;
          }
          if (j < D) x += mq * du[j < D ? j : 0];
        }
        CCB_UNROLL for (int b = 0; b < 3; ++b) x += rdn[b] * out.dN_dt[b];
        out.dq_dt[c][i] = x;
      }
    }
  }

  if (PASS == 1) {
    // ------------------------------------------------------------------ charges (reference :369-454)
    double F_ext_N[3], R_ext_N[3][3], M_ext_N[D][3];
    CCB_UNROLL for (int a = 0; a < 3; ++a) {
      CCB_UNROLL for (int b = 0; b < 3; ++b) R_ext_N[a][b] = 0.0;
      R_ext_N[a][a] += gamma * sigma;
      F_ext_N[a] = 0.0;  // rho_ext (external sources) is zero
    }
    CCB_UNROLL for (int i = 0; i < D; ++i) CCB_UNROLL for (int c = 0; c < 3; ++c) M_ext_N[i][c] = 0.0;
    if (fl.diffusion) {
      CCB_UNROLL for (int i = 0; i < D; ++i) CCB_UNROLL for (int c = 0; c < 3; ++c) M_ext_N[i][c] += -M_q0[i][c];
      CCB_UNROLL for (int a = 0; a < 3; ++a) {
        CCB_UNROLL for (int b = 0; b < 3; ++b) R_ext_N[a][b] += -R_q0[a][b];
        F_ext_N[a] += -F_q0[a] + in.grad_qa[a] - in.div_qa[a] - CCB_OT(in.q[a][0]);
      }
    }

    // ------------------------------------------------------------------ du/dt (reference :459-619)
    {
      const double w = th.w, s = th.s;
      const double rhoVec[3] = {th.rhoB, th.rhoS, th.rhoQ}, dwdrho[3] = {th.dwdB, th.dwdS, th.dwdQ};
      double gradP_c[D], gradBulk_c[D];
      raise_vec<D, MILNE>(m, in.gradP, gradP_c);
      raise_vec<D, MILNE>(m, in.gradBulk, gradBulk_c);
      double rho_dw = 0.;
      CCB_UNROLL for (int c = 0; c < 3; ++c) rho_dw += rhoVec[c] * dwdrho[c];
      const double aw = (s * th.dwds * ig + rho_dw * ig);
      double M_w[D], R_w[3];
      CCB_UNROLL for (int i = 0; i < D; ++i) M_w[i] = aw * ucov[i] * ig + sigma * th.dwds * M_ext_entropy[i];
      CCB_UNROLL for (int c = 0; c < 3; ++c) R_w[c] = sigma * dwdrho[c] + sigma * th.dwds * R_ext_entropy[c];
      const double F_w = -aw * theta_geo + sigma * th.dwds * F_ext_entropy;
      const double F_bulk = sigma * F_ext_bulk - bulk * theta_geo * ig;
      double M_u[D][D], F_u[D], R_u[D][3];
      CCB_UNROLL for (int i = 0; i < D; ++i) {
        CCB_UNROLL for (int j = 0; j < D; ++j)
          M_u[i][j] = u[i] * gamma * (M_ext_bulk[j] + M_w[j]) - u[i] * ucov[j] * (w + bulk) * ig;
        M_u[i][i] += gamma * (w + bulk);
        F_u[i] = (gradP_c[i] + gradBulk_c[i]) - gamma * (F_w + F_bulk) * u[i] - (w + bulk) * theta_geo * u[i] -
                 CCB_OT(2. * (w + bulk) * gamma * u[D - 1] * die[i]);
        CCB_UNROLL for (int c = 0; c < 3; ++c) R_u[i][c] = -u[i] * gamma * (R_w[c] + R_ext_bulk[c] * sigma);
      }
      if (fl.shear) {
        CCB_UNROLL for (int i = 0; i < D; ++i) {
          F_u[i] += -CCB_OT(shv[i + 1][0]) - CCB_OT(2. * die[i] * shv[0][3]) - in.divshear[i] + in.gradshear[i] - F_0i_shear[i];
          CCB_UNROLL for (int j = 0; j < D; ++j) M_u[i][j] += M_0i_shear[i][j];
        }
      }
      double RI[3][3];
      inverse3(R_ext_N, RI);
      double A[D][D], rhs[D];
      CCB_UNROLL for (int i = 0; i < D; ++i) {
        double fr = 0.0;
        CCB_UNROLL for (int j = 0; j < D; ++j) {
          double mr = 0.0;
          CCB_UNROLL for (int a = 0; a < 3; ++a) CCB_UNROLL for (int b = 0; b < 3; ++b) mr += R_u[i][a] * RI[a][b] * M_ext_N[j][b];
          A[i][j] = M_u[i][j] - mr;
        }
        CCB_UNROLL for (int a = 0; a < 3; ++a) CCB_UNROLL for (int b = 0; b < 3; ++b) fr += R_u[i][a] * RI[a][b] * F_ext_N[b];
        rhs[i] = F_u[i] + fr;
      }
      if (D == 2) {
        const double A2[2][2] = {{A[0][0], A[0][1]}, {A[1][0], A[1][1]}};
        double Ai[2][2];
        inverse2(A2, Ai);
        CCB_UNROLL for (int i = 0; i < 2; ++i) out.du_dt[i] = Ai[i][0] * rhs[0] + Ai[i][1] * rhs[1];
      } else {
        double A3[3][3], Ai[3][3];
        CCB_UNROLL for (int i = 0; i < 3; ++i) CCB_UNROLL for (int j = 0; j < 3; ++j) A3[i][j] = A[i < D ? i : 0][j < D ? j : 0];
        inverse3(A3, Ai);
        CCB_UNROLL for (int i = 0; i < D; ++i) {
          double x = 0.;
          CCB_UNROLL for (int j = 0; j < D; ++j) x += Ai[i][j] * rhs[j];
          out.du_dt[i] = x;
        }
      }
    }

    // ------------------------------------------------------------------ everything that depends on du/dt (reference :623-826)
    {
      const double* du = out.du_dt;
      double MU_aux[3];
      CCB_UNROLL for (int c = 0; c < 3; ++c) { double x = 0.; CCB_UNROLL for (int i = 0; i < D; ++i) x += M_ext_N[i][c] * du[i]; MU_aux[c] = x; }
      double RNi[3][3];
      inverse3(R_ext_N, RNi);
      CCB_UNROLL for (int a = 0; a < 3; ++a) {
        double x = 0.;
        CCB_UNROLL for (int b = 0; b < 3; ++b) x += RNi[a][b] * (F_ext_N[b] + MU_aux[b]);
        out.dN_dt[a] = x;
      }
    }
  }
  if (PASS == 2) {
    {
      const double* du = out.du_dt;
      double a = 0., b = 0., c = 0., d = 0., e = 0., f = 0.;
      CCB_UNROLL for (int i = 0; i < D; ++i) { a += M_ext_entropy[i] * du[i]; b += M_ext_bulk[i] * du[i]; c += M_shv_nabla_u[i] * du[i]; d += ucov[i] * du[i]; }
      CCB_UNROLL for (int k = 0; k < 3; ++k) { e += R_ext_entropy[k] * out.dN_dt[k]; f += R_ext_bulk[k] * out.dN_dt[k]; }
      out.ds_dt = a + F_ext_entropy + e;
      out.dbulk_dt = b + F_ext_bulk + f;
      out.shv_nabla_u = c + F_shv_nabla_u;
      out.theta = MILNE ? (-1. * d * ig + gamma * divV + gamma * it_ + gf) : (-1. * d * ig + gamma * divV);
    }
    {  // longitudinal work bookkeeping (reference :828-863)
      const double u3 = (D == 3) ? u[D - 1] : 0.0;
      out.dEz = ((th.e + th.p + bulk) * u3 * u3 + (th.p + bulk) * i_g33 + shv[3][3]) * in.mass_s * m.g33 / in.sigma_lab;
    }
  }
#undef CCB_OT
#undef CCB_TT
}

template <int D, bool MILNE>
CCB_HD void eom_evaluate(const Metric<MILNE>& m, const EomIn<D>& in, const Thermo& th, const Transport& tc,
                         const double kappa_q[3][3], const double tau_q[3][3], const EomFlags fl, EomOut<D>& out) {
  eom_pass<D, MILNE, 1>(m, in, th, tc, kappa_q, tau_q, fl, out);
  eom_pass<D, MILNE, 2>(m, in, th, tc, kappa_q, tau_q, fl, out);
}

}  // namespace ccb
