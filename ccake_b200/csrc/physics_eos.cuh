// Equation of state on the device: forward thermodynamics at a phase-diagram
// point, the (s,B,S,Q) -> (T,muB,muQ,muS) inversion, and the thermodynamic
// derivatives the equations of motion consume.
//
// Reference: EquationOfState::{tbqs, evaluate_thermodynamics, set_thermo}
// (src/eos.cpp:128-221, include/eos.h:121-175), the derivative terms
// (src/eos.cpp:276-492, src/eos_derivatives.cpp:34-560) and the analytic EoS
// classes (include/eos_conformal.h). In the reference all of this is serial host
// code between two full-state deep copies (src/sph_workstation.cpp:1737-1762).
#pragma once

#include "physics_common.cuh"

namespace ccb {

// Order of the 17 forward quantities, reference src/eos_table.cpp:33-36 / src/eos.cpp:180-196.
enum { TH_P = 0, TH_S, TH_B, TH_SS, TH_Q, TH_E, TH_CS2, TH_B2, TH_Q2, TH_S2, TH_BQ, TH_BS, TH_QS, TH_TB, TH_TQ, TH_TS, TH_T2, TH_N };

// What one particle carries out of phase 04 (reference include/thermodynamic_info.h:36-68).
struct Thermo {
  double T, muB, muS, muQ, e, s, rhoB, rhoS, rhoQ, p, cs2, w, dwds, dwdB, dwdS, dwdQ;
  double dalpha[3][4];  // [B,S,Q][d/ds, d/dB, d/dS, d/dQ]
  double chiBB, chiBB0;
};

struct ConformalParams { double c, T0, muB0, muS0, muQ0; };

// p = cp x^2, x = (T/T0)^2 + sum (mu_i/mu_i0)^2 ; reference include/eos_conformal.h:33-194
CCB_HD void conformal_full_thermo(const ConformalParams& P, double T, double muB, double muQ, double muS, double* r) {
  const double x = (T / P.T0) * (T / P.T0) + (muB / P.muB0) * (muB / P.muB0) + (muS / P.muS0) * (muS / P.muS0) +
                   (muQ / P.muQ0) * (muQ / P.muQ0);
  const double cp = P.c * P.T0 * P.T0 * P.T0 * P.T0;
  const double T02 = P.T0 * P.T0, B02 = P.muB0 * P.muB0, S02 = P.muS0 * P.muS0, Q02 = P.muQ0 * P.muQ0;
  r[TH_P] = cp * x * x;
  r[TH_S] = 4.0 * cp * T * x / T02;
  r[TH_B] = 4.0 * cp * muB * x / B02;
  r[TH_SS] = 4.0 * cp * muS * x / S02;
  r[TH_Q] = 4.0 * cp * muQ * x / Q02;
  r[TH_E] = (r[TH_S] * T + muB * r[TH_B] + muQ * r[TH_Q] + muS * r[TH_SS]) - r[TH_P];
  r[TH_CS2] = 1.0 / 3.0;
  r[TH_B2] = cp * (8.0 * muB * muB / (B02 * B02) + 4.0 * x / B02);
  r[TH_Q2] = cp * (8.0 * muQ * muQ / (Q02 * Q02) + 4.0 * x / Q02);
  r[TH_S2] = cp * (8.0 * muS * muS / (S02 * S02) + 4.0 * x / S02);
  r[TH_BQ] = 8.0 * cp * muB * muQ / (B02 * Q02);
  r[TH_BS] = 8.0 * cp * muB * muS / (B02 * S02);
  r[TH_QS] = 8.0 * cp * muS * muQ / (S02 * Q02);
  r[TH_TB] = 8.0 * cp * T * muB / (T02 * B02);
  r[TH_TQ] = 8.0 * cp * T * muQ / (T02 * Q02);
  r[TH_TS] = 8.0 * cp * T * muS / (T02 * S02);
  r[TH_T2] = cp * (8.0 * T * T / (T02 * T02) + 4.0 * x / T02);
}

// (s,B,S,Q) -> (T,muB,muS,muQ) for the conformal EoS. With y = (T/T0, mu_i/mu_i0) the
// densities are d = y |y|^2 (d_0 = s T0/(4cp), d_i = rho_i mu_i0/(4cp)), hence
// y = d / |d|^(2/3): no iteration. The reference reaches the same point with GSL
// hybrids from the particle's previous (T,mu) to sum|f| < 1e-10 (src/rootfinder.cpp:206-359).
CCB_HD bool conformal_invert_s(const ConformalParams& P, double s, double B, double S, double Q, double& T, double& muB,
                               double& muS, double& muQ) {
  const double cp4 = 4.0 * P.c * P.T0 * P.T0 * P.T0 * P.T0;
  const double d0 = s * P.T0 / cp4, dB = B * P.muB0 / cp4, dS = S * P.muS0 / cp4, dQ = Q * P.muQ0 / cp4;
  const double n2 = d0 * d0 + dB * dB + dS * dS + dQ * dQ;
  if (!(s > 0.0) || !(n2 > 0.0)) return false;
  const double f = 1.0 / cbrt(n2);  // |d|^(-2/3)
  T = P.T0 * d0 * f; muB = P.muB0 * dB * f; muS = P.muS0 * dS * f; muQ = P.muQ0 * dQ * f;
  return true;
}
// (e,B,S,Q) -> (T,mu): e = 3 cp x^2 fixes x, rho_i = 4 cp mu_i x / mu_i0^2 fixes mu_i, T closes x.
CCB_HD bool conformal_invert_e(const ConformalParams& P, double e, double B, double S, double Q, double& T, double& muB,
                               double& muS, double& muQ) {
  const double cp = P.c * P.T0 * P.T0 * P.T0 * P.T0;
  if (!(e > 0.0)) return false;
  const double x = sqrt(e / (3.0 * cp));
  muB = B * P.muB0 * P.muB0 / (4.0 * cp * x);
  muS = S * P.muS0 * P.muS0 / (4.0 * cp * x);
  muQ = Q * P.muQ0 * P.muQ0 / (4.0 * cp * x);
  const double y2 = x - (muB / P.muB0) * (muB / P.muB0) - (muS / P.muS0) * (muS / P.muS0) - (muQ / P.muQ0) * (muQ / P.muQ0);
  if (!(y2 > 0.0)) return false;
  T = P.T0 * sqrt(y2);
  return true;
}

// a^T m^{-1} b with the reference's diagonal guard (src/eos_derivatives.cpp:457-480):
// |m11|, |m22| below 1e-3 are replaced by +-1e-3 before the factorisation. Solved by
// elimination with partial pivoting (the reference inverts through GSL LU, same pivot rule).
CCB_HD double aT_minv_b(const double a[3], double m[3][3], const double b[3]) {
  const double EPS = 1e-3;
  if (fabs(m[1][1]) < EPS) m[1][1] = (m[1][1] >= 0.0 ? EPS : -EPS);
  if (fabs(m[2][2]) < EPS) m[2][2] = (m[2][2] >= 0.0 ? EPS : -EPS);
  double y[3] = {b[0], b[1], b[2]}, ip[3];
  CCB_UNROLL for (int j = 0; j < 3; ++j) {
    int piv = j;
    double amax = fabs(m[j][j]);
    CCB_UNROLL for (int i = j + 1; i < 3; ++i) if (fabs(m[i][j]) > amax) { amax = fabs(m[i][j]); piv = i; }
    if (piv != j) {
      CCB_UNROLL for (int k = 0; k < 3; ++k) { const double tmp = m[j][k]; m[j][k] = m[piv][k]; m[piv][k] = tmp; }
      const double tmp = y[j]; y[j] = y[piv]; y[piv] = tmp;
    }
    ip[j] = 1.0 / m[j][j];  // one reciprocal per pivot, reused by the back substitution
    CCB_UNROLL for (int i = j + 1; i < 3; ++i) {
      const double f = m[i][j] * ip[j];
      CCB_UNROLL for (int k = j; k < 3; ++k) m[i][k] -= f * m[j][k];
      y[i] -= f * y[j];
    }
  }
  y[2] = y[2] * ip[2];
  y[1] = (y[1] - m[1][2] * y[2]) * ip[1];
  y[0] = (y[0] - m[0][1] * y[1] - m[0][2] * y[2]) * ip[0];
  return a[0] * y[0] + a[1] * y[1] + a[2] * y[2];
}

// The sixteen constrained derivatives (ds/dT at fixed rho, ...). Each is
// h - a^T m^{-1} b with (h, a, m, b) picked from the pressure Hessian exactly as the
// reference lays them out (src/eos_derivatives.cpp:34-455); the layouts are kept
// literal because the 1e-3 guard above makes them differ from the plain cofactor identity.
struct Hess { double t2, tb, ts, tq, b2, bs, bq, s2, sq, q2; };

#define CCB_TERM(h, a0, a1, a2, m00, m01, m02, m10, m11, m12, m20, m21, m22, b0, b1, b2_) \
  { const double a[3] = {a0, a1, a2}; const double b[3] = {b0, b1, b2_};                  \
    double m[3][3] = {{m00, m01, m02}, {m10, m11, m12}, {m20, m21, m22}};                 \
    return (h) - aT_minv_b(a, m, b); }

// mode: 0 full, 1 T-only, 2 baryon-only (EquationOfState::tbqs, src/eos.cpp:139-144)
CCB_HD double dentr_dt(const Hess& H, int mode) {
  if (mode == 1) return H.t2;
  if (mode == 2) return H.t2 - (H.tb * H.tb) / (H.b2 + 1e-30);
  CCB_TERM(H.t2, H.tb, H.ts, H.tq, H.b2, H.bs, H.bq, H.bs, H.s2, H.sq, H.bq, H.sq, H.q2, H.tb, H.ts, H.tq)
}
CCB_HD double dentr_dmub(const Hess& H, int mode) {
  if (mode == 1) return 0.0;
  if (mode == 2) return H.tb - (H.t2 * H.b2) / (H.tb + 1e-30);
  CCB_TERM(H.tb, H.t2, H.ts, H.tq, H.tb, H.bs, H.bq, H.ts, H.s2, H.sq, H.tq, H.sq, H.q2, H.b2, H.bs, H.bq)
}
CCB_HD double dentr_dmus(const Hess& H, int mode) {
  if (mode != 0) return 0.0;
  CCB_TERM(H.ts, H.t2, H.tb, H.tq, H.tb, H.b2, H.bq, H.ts, H.bs, H.sq, H.tq, H.bq, H.q2, H.bs, H.s2, H.sq)
}
CCB_HD double dentr_dmuq(const Hess& H, int mode) {
  if (mode != 0) return 0.0;
  CCB_TERM(H.tq, H.t2, H.tb, H.ts, H.tb, H.b2, H.bs, H.ts, H.bs, H.s2, H.tq, H.bq, H.sq, H.bq, H.sq, H.q2)
}
CCB_HD double db_dt(const Hess& H, int mode) {
  if (mode == 1) return 0.0;
  if (mode == 2) return H.tb - (H.b2 * H.t2) / (H.tb + 1e-30);
  CCB_TERM(H.tb, H.b2, H.bs, H.bq, H.tb, H.ts, H.tq, H.bs, H.s2, H.sq, H.bq, H.sq, H.q2, H.t2, H.ts, H.tq)
}
CCB_HD double ds_dt(const Hess& H, int mode) {
  if (mode != 0) return 0.0;
  CCB_TERM(H.ts, H.bs, H.s2, H.sq, H.tb, H.ts, H.tq, H.b2, H.bs, H.bq, H.bq, H.sq, H.q2, H.t2, H.tb, H.tq)
}
CCB_HD double dq_dt(const Hess& H, int mode) {
  if (mode != 0) return 0.0;
  CCB_TERM(H.tq, H.bq, H.sq, H.q2, H.tb, H.ts, H.tq, H.b2, H.bs, H.bq, H.bs, H.s2, H.sq, H.t2, H.tb, H.ts)
}
// calc_term_4(j, i): "d rho_j / d mu_i"
CCB_HD double db_dmub(const Hess& H, int mode) {  // j=B, i=B
  if (mode == 1) return 0.0;
  if (mode == 2) return H.b2 - (H.tb * H.tb) / (H.t2 + 1e-30);
  CCB_TERM(H.b2, H.tb, H.bs, H.bq, H.t2, H.ts, H.tq, H.ts, H.s2, H.sq, H.tq, H.sq, H.q2, H.tb, H.bs, H.bq)
}
CCB_HD double ds_dmub(const Hess& H, int mode) {  // j=S, i=B
  if (mode != 0) return 0.0;
  CCB_TERM(H.bs, H.ts, H.s2, H.sq, H.t2, H.ts, H.tq, H.tb, H.bs, H.bq, H.tq, H.sq, H.q2, H.tb, H.b2, H.bq)
}
CCB_HD double dq_dmub(const Hess& H, int mode) {  // j=Q, i=B
  if (mode != 0) return 0.0;
  CCB_TERM(H.bq, H.tq, H.sq, H.q2, H.t2, H.ts, H.tq, H.tb, H.bs, H.bq, H.ts, H.s2, H.sq, H.tb, H.b2, H.bs)
}
CCB_HD double db_dmus(const Hess& H, int mode) {  // j=B, i=S
  if (mode != 0) return 0.0;
  CCB_TERM(H.bs, H.tb, H.b2, H.bq, H.t2, H.tb, H.tq, H.ts, H.bs, H.sq, H.tq, H.bq, H.q2, H.ts, H.s2, H.sq)
}
CCB_HD double ds_dmus(const Hess& H, int mode) {  // j=S, i=S
  if (mode != 0) return 0.0;
  CCB_TERM(H.s2, H.ts, H.bs, H.sq, H.t2, H.tb, H.tq, H.tb, H.b2, H.bq, H.tq, H.bq, H.q2, H.ts, H.bs, H.sq)
}
CCB_HD double dq_dmus(const Hess& H, int mode) {  // j=Q, i=S
  if (mode != 0) return 0.0;
  CCB_TERM(H.sq, H.tq, H.bq, H.q2, H.t2, H.tb, H.tq, H.tb, H.b2, H.bq, H.ts, H.bs, H.sq, H.ts, H.bs, H.s2)
}
CCB_HD double db_dmuq(const Hess& H, int mode) {  // j=B, i=Q
  if (mode != 0) return 0.0;
  CCB_TERM(H.bq, H.tb, H.b2, H.bs, H.t2, H.tb, H.ts, H.ts, H.bs, H.s2, H.tq, H.bq, H.sq, H.tq, H.sq, H.q2)
}
CCB_HD double ds_dmuq(const Hess& H, int mode) {  // j=S, i=Q
  if (mode != 0) return 0.0;
  CCB_TERM(H.sq, H.ts, H.bs, H.s2, H.t2, H.tb, H.ts, H.tb, H.b2, H.bs, H.tq, H.bq, H.sq, H.tq, H.bq, H.q2)
}
CCB_HD double dq_dmuq(const Hess& H, int mode) {  // j=Q, i=Q
  if (mode != 0) return 0.0;
  CCB_TERM(H.q2, H.tq, H.bq, H.sq, H.t2, H.tb, H.ts, H.tb, H.b2, H.bs, H.ts, H.bs, H.s2, H.tq, H.bq, H.sq)
}
#undef CCB_TERM

// EquationOfState::set_thermo (include/eos.h:121-148) from the 17 forward values at (T,muB,muQ,muS).
// chiBB0 is the caller's business (table EoS re-evaluates at mu = 0, src/eos.cpp:199-210).
CCB_HD void thermo_from_forward(double T, double muB, double muQ, double muS, const double* r, int mode, Thermo& th) {
  th.T = T; th.muB = muB; th.muS = muS; th.muQ = muQ;
  th.p = r[TH_P]; th.s = r[TH_S]; th.rhoB = r[TH_B]; th.rhoS = r[TH_SS]; th.rhoQ = r[TH_Q]; th.e = r[TH_E];
  th.w = r[TH_E] + r[TH_P]; th.cs2 = r[TH_CS2];
  th.chiBB = r[TH_B2]; th.chiBB0 = r[TH_B2];
  Hess H;
  H.t2 = r[TH_T2]; H.tb = r[TH_TB]; H.ts = r[TH_TS]; H.tq = r[TH_TQ]; H.b2 = r[TH_B2]; H.bs = r[TH_BS];
  H.bq = r[TH_BQ]; H.s2 = r[TH_S2]; H.sq = r[TH_QS]; H.q2 = r[TH_Q2];
  const double B = r[TH_B], S = r[TH_SS], Q = r[TH_Q], s = r[TH_S];
  const bool hasB = fabs(B) > kTiny, hasS = fabs(S) > kTiny, hasQ = fabs(Q) > kTiny;
  const double sT = dentr_dt(H, mode);
  const double sB = hasB ? dentr_dmub(H, mode) : 0.0, sS = hasS ? dentr_dmus(H, mode) : 0.0,
               sQ = hasQ ? dentr_dmuq(H, mode) : 0.0;
  // Reciprocals of T and of each constrained derivative are formed once (the reference divides ~40 times).
  const double iT = 1.0 / T, iT2 = iT * iT;
  const double isT = 1.0 / sT, isB = hasB ? 1.0 / sB : 0.0, isS = hasS ? 1.0 / sS : 0.0, isQ = hasQ ? 1.0 / sQ : 0.0;
  {  // dwds, src/eos.cpp:276-292
    double ct = 0.0;
    if (hasB) ct += B * isB;
    if (hasQ) ct += Q * isQ;
    if (hasS) ct += S * isS;
    th.dwds = T + s * isT + ct;
  }
  CCB_UNROLL for (int a = 0; a < 3; ++a) CCB_UNROLL for (int x = 0; x < 4; ++x) th.dalpha[a][x] = 0.0;
  // dalpha_X/ds, src/eos.cpp:336-366
  th.dalpha[0][0] = -muB * iT2 * isT + (hasB ? iT * isB : 0.0);
  th.dalpha[1][0] = -muS * iT2 * isT + (hasS ? iT * isS : 0.0);
  th.dalpha[2][0] = -muQ * iT2 * isT + (hasQ ? iT * isQ : 0.0);
  th.dwdB = muB; th.dwdS = muS; th.dwdQ = muQ;
  if (hasB) {  // src/eos.cpp:294-305, :369-404
    const double ibT = 1.0 / db_dt(H, mode), ibB = 1.0 / db_dmub(H, mode);
    double ct = s * ibT + B * ibB;
    const double ibQ = hasQ ? 1.0 / db_dmuq(H, mode) : 0.0, ibS = hasS ? 1.0 / db_dmus(H, mode) : 0.0;
    if (hasQ) ct += Q * ibQ;
    if (hasS) ct += S * ibS;
    th.dwdB = muB + ct;
    th.dalpha[0][1] = iT * ibB - muB * iT2 * ibT;
    th.dalpha[1][1] = -muS * iT2 * ibT + (hasS ? iT * ibS : 0.0);
    th.dalpha[2][1] = -muQ * iT2 * ibT + (hasQ ? iT * ibQ : 0.0);
  }
  if (hasS) {  // src/eos.cpp:307-318, :406-446
    const double ixT = 1.0 / ds_dt(H, mode), ixS = 1.0 / ds_dmus(H, mode);
    double ct = s * ixT + S * ixS;
    const double ixB = hasB ? 1.0 / ds_dmub(H, mode) : 0.0, ixQ = hasQ ? 1.0 / ds_dmuq(H, mode) : 0.0;
    if (hasB) ct += B * ixB;
    if (hasQ) ct += Q * ixQ;
    th.dwdS = muS + ct;
    th.dalpha[0][2] = -muB * iT2 * ixT + (hasB ? iT * ixB : 0.0);
    th.dalpha[1][2] = -muS * iT2 * ixT + iT * ixS;
    th.dalpha[2][2] = -muQ * iT2 * ixT + (hasQ ? iT * ixQ : 0.0);
  }
  if (hasQ) {  // src/eos.cpp:320-331, :448-488
    const double ixT = 1.0 / dq_dt(H, mode), ixQ = 1.0 / dq_dmuq(H, mode);
    double ct = s * ixT + Q * ixQ;
    const double ixB = hasB ? 1.0 / dq_dmub(H, mode) : 0.0, ixS = hasS ? 1.0 / dq_dmus(H, mode) : 0.0;
    if (hasB) ct += B * ixB;
    if (hasS) ct += S * ixS;
    th.dwdQ = muQ + ct;
    th.dalpha[0][3] = -muB * iT2 * ixT + (hasB ? iT * ixB : 0.0);
    th.dalpha[1][3] = -muS * iT2 * ixT + (hasS ? iT * ixS : 0.0);
    th.dalpha[2][3] = -muQ * iT2 * ixT + iT * ixQ;
  }
}

}  // namespace ccb
