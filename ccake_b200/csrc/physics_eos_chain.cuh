// The tabulated equation of state and its fall-back chain on the device:
//   4-D multilinear table interpolation        InterpolatorND<4>::evaluate       (src/interpolatorND.cpp:250-390)
//   table / analytic EoS forward evaluation    EoS_table::get_*                  (src/eos_table.cpp:131-412)
//                                              EoS_tanh_conformal                (include/eos_tanh_conformal.h)
//                                              EoS_conformal_diagonal            (include/eos_conformal_diagonal.h)
//   densities -> (T, muB, muQ, muS)            Rootfinder::find_root             (src/rootfinder.cpp:423-716)
//   seed cascade and EoS chain                 EquationOfState::rootfinder_update_s / rootfinder_s_out
//                                                                                 (src/eos.cpp:638-804, 834-977)
//
// The reference solves with GSL's hybrids (Powell dogleg, finite-difference Jacobian + Broyden updates,
// sum|f| < 1e-10). Here the root is found by Newton's method with the EXACT Jacobian: for the table that is the
// derivative of the multilinear interpolant itself (it comes out of the same 16 vertex loads as the values), for
// the analytic EoSs the pressure Hessian. Both stop on the reference's criterion sum|f| < 1e-10 (fm^-3), inside
// the reference's bounds, from the reference's sequence of seeds, so where the reference finds a root this finds
// the same one (the interpolant is piecewise multilinear: the root in a cell is unique); the iterates differ.
// The conformal and diagonal-conformal members of the chain are inverted in closed form.
//
// Order conventions (the reference's): phase-diagram point x = (T, muB, muQ, muS); densities returned by an
// EoS = (s or e, rhoB, rhoS, rhoQ); residuals f = (s, B, Q, S) - targets (src/rootfinder.cpp:72-160).
#pragma once

#include "physics_eos.cuh"

namespace ccb {

enum { EOS_KIND_TABLE = 0, EOS_KIND_TANH = 1, EOS_KIND_CONFORMAL = 2, EOS_KIND_DIAGONAL = 3 };
constexpr double kTbqsInfinity = 1e10;  // reference include/eos_header.h:15
constexpr double kRootTolerance = 1e-10;  // reference include/rootfinder.h:37
constexpr int kNewtonMaxIter = 80;

struct TableView {
  int n[4];                          // nodes along T, muB, muQ, muS
  double min[4], max[4], delta[4];   // InterpolatorND::construct_interpolant (src/interpolatorND.cpp:195-226)
  const double* hot_s;               // [node][4]  s, rhoB, rhoS, rhoQ   (34 MB for 155 x 19^3: L2-resident)
  const double* hot_e;               // [node][4]  e, rhoB, rhoS, rhoQ   (initialisation only)
  const double* full;                // [node][17] every field, touched once per particle after convergence
};

struct AnalyticParams { double c, T0, muB0, muS0, muQ0, Tc, Ts; };

struct EosChain {
  int n;            // members, in the reference's order: table, tanh_conformal, conformal, conformal_diagonal
  int kind[4];
  TableView table;
  AnalyticParams tanh_p, conf_p, diag_p;
  int prohibit_unstable_cs2, prohibit_acausal_cs2, restrict_mu_T_ratios;
};

CCB_HD void eos_bounds(const EosChain& ch, int k, double lo[4], double hi[4]) {
  if (ch.kind[k] == EOS_KIND_TABLE) {
    CCB_UNROLL for (int d = 0; d < 4; ++d) { lo[d] = ch.table.min[d]; hi[d] = ch.table.max[d]; }
  } else {  // src/eos_initialization.cpp:96-97
    lo[0] = 0.0; hi[0] = kTbqsInfinity;
    CCB_UNROLL for (int d = 1; d < 4; ++d) { lo[d] = -kTbqsInfinity; hi[d] = kTbqsInfinity; }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Table. Cell location exactly as the reference computes it (modf of the scaled coordinate, the last node folded
// into the last cell, degenerate axes pinned).
struct TableCell {
  long long base;
  double fr[4];       // the reference's "fracs": weight of the LOWER node along each axis
  long long stride[4];
};

CCB_HD bool table_locate(const TableView& tb, const double x[4], TableCell& c) {
  int ind[4];
  bool bad = false;
  CCB_UNROLL for (int d = 0; d < 4; ++d) {
    if (x[d] != x[d]) return false;  // NaN coordinates: fill value (src/interpolatorND.cpp:255-260)
    if (tb.n[d] <= 1 || fabs(tb.delta[d]) < 1e-30) { ind[d] = 0; c.fr[d] = 1.0; continue; }
    double whole = 0.0;
    c.fr[d] = 1.0 - modf((x[d] - tb.min[d]) / tb.delta[d], &whole);
    if (!(fabs(whole) < 2.0e9)) return false;
    ind[d] = (int)whole;
    if (ind[d] + 1 == tb.n[d]) { c.fr[d] = 1.0 - c.fr[d]; ind[d]--; }
    if (ind[d] < 0 || ind[d] >= tb.n[d]) bad = true;
  }
  if (bad) return false;
  c.stride[3] = 1; c.stride[2] = tb.n[3]; c.stride[1] = (long long)tb.n[3] * tb.n[2]; c.stride[0] = c.stride[1] * tb.n[1];
  c.base = 0;
  CCB_UNROLL for (int d = 0; d < 4; ++d) {
    c.base += ind[d] * c.stride[d];
    if (tb.n[d] <= 1) c.stride[d] = 0;  // the upper vertex of a pinned axis carries weight 0: never dereference it
  }
  return true;
}

#if defined(__CUDA_ARCH__)
#define CCB_LDG(p) __ldg(p)
#else
#define CCB_LDG(p) (*(p))
#endif

// Weight of hypercube vertex h (bit 3 = T ... bit 0 = muS, the reference's vertex order) as the reference forms it:
// the running product of (frac + b - 2 b frac) over the axes (src/interpolatorND.cpp:295-299).
CCB_HD double vertex_weight(const double fr[4], int h) {
  double w = 1.0;
  CCB_UNROLL for (int d = 0; d < 4; ++d) {
    const double b = (double)((h >> (3 - d)) & 1);
    w *= (fr[d] + b - 2.0 * b * fr[d]);
  }
  return w;
}

// Four densities (s|e, B, S, Q) at x and, optionally, their exact derivatives with respect to x: jac[field][axis].
CCB_HD bool table_densities(const TableView& tb, const double* __restrict__ hot, const double x[4], double val[4], double (*jac)[4]) {
  TableCell c;
  if (!table_locate(tb, x, c)) return false;
  CCB_UNROLL for (int f = 0; f < 4; ++f) val[f] = 0.0;
  if (jac) { CCB_UNROLL for (int f = 0; f < 4; ++f) CCB_UNROLL for (int d = 0; d < 4; ++d) jac[f][d] = 0.0; }
  double a[4][2];
  CCB_UNROLL for (int d = 0; d < 4; ++d) { a[d][0] = c.fr[d]; a[d][1] = c.fr[d] + 1.0 - 2.0 * c.fr[d]; }
  CCB_UNROLL for (int h = 0; h < 16; ++h) {
    const int b0 = (h >> 3) & 1, b1 = (h >> 2) & 1, b2 = (h >> 1) & 1, b3 = h & 1;
    const long long node = c.base + b0 * c.stride[0] + b1 * c.stride[1] + b2 * c.stride[2] + b3 * c.stride[3];
#if defined(__CUDA_ARCH__)
    const double2* p = reinterpret_cast<const double2*>(hot + 4 * node);  // one 32-byte sector per vertex
    const double2 v01 = __ldg(p), v23 = __ldg(p + 1);
    const double v[4] = {v01.x, v01.y, v23.x, v23.y};
#else
    const double* p = hot + 4 * node;
    const double v[4] = {p[0], p[1], p[2], p[3]};
#endif
    const double w = ((1.0 * a[0][b0]) * a[1][b1]) * a[2][b2] * a[3][b3];
    CCB_UNROLL for (int f = 0; f < 4; ++f) val[f] += w * v[f];
    if (jac) {
      const double w0 = (b0 ? 1.0 : -1.0) * a[1][b1] * a[2][b2] * a[3][b3];
      const double w1 = (b1 ? 1.0 : -1.0) * a[0][b0] * a[2][b2] * a[3][b3];
      const double w2 = (b2 ? 1.0 : -1.0) * a[0][b0] * a[1][b1] * a[3][b3];
      const double w3 = (b3 ? 1.0 : -1.0) * a[0][b0] * a[1][b1] * a[2][b2];
      CCB_UNROLL for (int f = 0; f < 4; ++f) { jac[f][0] += w0 * v[f]; jac[f][1] += w1 * v[f]; jac[f][2] += w2 * v[f]; jac[f][3] += w3 * v[f]; }
    }
  }
  if (jac) {
    CCB_UNROLL for (int d = 0; d < 4; ++d) {
      const double s = (tb.n[d] <= 1 || fabs(tb.delta[d]) < 1e-30) ? 0.0 : 1.0 / tb.delta[d];
      CCB_UNROLL for (int f = 0; f < 4; ++f) jac[f][d] *= s;
    }
  }
  return true;
}

// All 17 fields at x (EoS_table::get_full_thermo, src/eos_table.cpp:345-412). Out of range: NaN fill.
CCB_HD bool table_full_thermo(const TableView& tb, const double x[4], double r[TH_N]) {
  TableCell c;
  if (!table_locate(tb, x, c)) {
    CCB_UNROLL for (int f = 0; f < TH_N; ++f) r[f] = nan("");
    return false;
  }
  CCB_UNROLL for (int f = 0; f < TH_N; ++f) r[f] = 0.0;
  for (int h = 0; h < 16; ++h) {
    const int b0 = (h >> 3) & 1, b1 = (h >> 2) & 1, b2 = (h >> 1) & 1, b3 = h & 1;
    const long long node = c.base + b0 * c.stride[0] + b1 * c.stride[1] + b2 * c.stride[2] + b3 * c.stride[3];
    const double w = vertex_weight(c.fr, h);
    const double* p = tb.full + TH_N * node;
    CCB_UNROLL for (int f = 0; f < TH_N; ++f) r[f] += w * CCB_LDG(p + f);
  }
  return true;
}
CCB_HD double table_field(const TableView& tb, const double x[4], int field) {
  TableCell c;
  if (!table_locate(tb, x, c)) return nan("");
  double r = 0.0;
  for (int h = 0; h < 16; ++h) {
    const int b0 = (h >> 3) & 1, b1 = (h >> 2) & 1, b2 = (h >> 1) & 1, b3 = h & 1;
    const long long node = c.base + b0 * c.stride[0] + b1 * c.stride[1] + b2 * c.stride[2] + b3 * c.stride[3];
    r += vertex_weight(c.fr, h) * CCB_LDG(tb.full + TH_N * node + field);
  }
  return r;
}

// Gaussian elimination with partial pivoting of the 4 x 5 augmented system A, every index a compile-time constant after unrolling
// (a row addressed by a run-time pivot index would put the whole matrix into local memory: the Newton loop then issues ~20
// local stores per iterate). Rows are exchanged with the arg-max row only, as a single swap, so the arithmetic is that of the
// textbook algorithm. Returns false for a singular matrix. On success x solves the system.
CCB_HD bool solve4_pivoted(double A[4][5], double x[4]) {
  CCB_UNROLL for (int j = 0; j < 4; ++j) {
    int piv = j;
    double amax = fabs(A[j][j]);
    CCB_UNROLL for (int i = j + 1; i < 4; ++i) { const double v = fabs(A[i][j]); if (v > amax) { amax = v; piv = i; } }
    if (!(amax > 0.0)) return false;
    CCB_UNROLL for (int i = j + 1; i < 4; ++i) {
      const bool sw = piv == i;
      CCB_UNROLL for (int c = 0; c < 5; ++c) { const double a = A[j][c], b = A[i][c]; A[j][c] = sw ? b : a; A[i][c] = sw ? a : b; }
    }
    const double ip = 1.0 / A[j][j];
    CCB_UNROLL for (int i = j + 1; i < 4; ++i) {
      const double m = A[i][j] * ip;
      CCB_UNROLL for (int c = j; c < 5; ++c) A[i][c] -= m * A[j][c];
    }
  }
  x[3] = A[3][4] / A[3][3];
  x[2] = (A[2][4] - A[2][3] * x[3]) / A[2][2];
  x[1] = (A[1][4] - A[1][2] * x[2] - A[1][3] * x[3]) / A[1][1];
  x[0] = (A[0][4] - A[0][1] * x[1] - A[0][2] * x[2] - A[0][3] * x[3]) / A[0][0];
  return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// tanh-modulated conformal EoS: p = t(T) * c T0^4 x^2, t = (1 + tanh((T - Tc)/Ts))/2 (include/eos_tanh_conformal.h:22-29, 215-262)
CCB_HD void tanh_full_thermo(const AnalyticParams& P, double T, double muB, double muQ, double muS, double* r) {
  const ConformalParams C{P.c, P.T0, P.muB0, P.muS0, P.muQ0};
  double c[TH_N];
  conformal_full_thermo(C, T, muB, muQ, muS, c);
  const double arg = (T - P.Tc) / P.Ts, th = tanh(arg), sech = 1.0 / cosh(arg);
  const double tf = 0.5 * (1.0 + th), tp = 0.5 * (sech * sech) / P.Ts, tpp = -2.0 * th * tp / P.Ts;
  r[TH_P] = tf * c[TH_P];
  r[TH_S] = tf * c[TH_S] + tp * c[TH_P];
  r[TH_B] = tf * c[TH_B]; r[TH_SS] = tf * c[TH_SS]; r[TH_Q] = tf * c[TH_Q];
  r[TH_E] = (r[TH_S] * T + muB * r[TH_B] + muQ * r[TH_Q] + muS * r[TH_SS]) - r[TH_P];
  r[TH_B2] = tf * c[TH_B2]; r[TH_Q2] = tf * c[TH_Q2]; r[TH_S2] = tf * c[TH_S2];
  r[TH_BQ] = tf * c[TH_BQ]; r[TH_BS] = tf * c[TH_BS]; r[TH_QS] = tf * c[TH_QS];
  r[TH_TB] = tf * c[TH_TB] + tp * c[TH_B];
  r[TH_TQ] = tf * c[TH_TQ] + tp * c[TH_Q];
  r[TH_TS] = tf * c[TH_TS] + tp * c[TH_SS];
  r[TH_T2] = tf * c[TH_T2] + 2.0 * tp * c[TH_S] + tpp * c[TH_P];
  // Speed of sound at fixed s/rho_X. The reference spells out the cofactor expansion of n^T H^-1 n / w
  // (include/eos_tanh_conformal.h:283-301: n = (s, B, Q, S), H = pressure Hessian, w = T s + mu.rho); the same
  // number is obtained here from one 4x4 elimination.
  double H[4][5] = {{r[TH_T2], r[TH_TB], r[TH_TQ], r[TH_TS], r[TH_S]},
                    {r[TH_TB], r[TH_B2], r[TH_BQ], r[TH_BS], r[TH_B]},
                    {r[TH_TQ], r[TH_BQ], r[TH_Q2], r[TH_QS], r[TH_Q]},
                    {r[TH_TS], r[TH_BS], r[TH_QS], r[TH_S2], r[TH_SS]}};
  const double n[4] = {r[TH_S], r[TH_B], r[TH_Q], r[TH_SS]};
  double y[4];
  if (!solve4_pivoted(H, y)) { y[0] = y[1] = y[2] = y[3] = nan(""); }
  const double w = T * r[TH_S] + muB * r[TH_B] + muQ * r[TH_Q] + muS * r[TH_SS];
  r[TH_CS2] = (n[0] * y[0] + n[1] * y[1] + n[2] * y[2] + n[3] * y[3]) / w;
}

// Diagonal conformal EoS: p = c T0^4 [(T/T0)^4 + sum (mu_i/mu_i0)^4] (include/eos_conformal_diagonal.h:78-205)
CCB_HD void diagonal_full_thermo(const AnalyticParams& P, double T, double muB, double muQ, double muS, double* r) {
  const double cp = P.c * P.T0 * P.T0 * P.T0 * P.T0;
  const double T04 = P.T0 * P.T0 * P.T0 * P.T0, B04 = P.muB0 * P.muB0 * P.muB0 * P.muB0, S04 = P.muS0 * P.muS0 * P.muS0 * P.muS0,
               Q04 = P.muQ0 * P.muQ0 * P.muQ0 * P.muQ0;
  const double x = (T / P.T0) * (T / P.T0) * (T / P.T0) * (T / P.T0) + (muB / P.muB0) * (muB / P.muB0) * (muB / P.muB0) * (muB / P.muB0) +
                   (muS / P.muS0) * (muS / P.muS0) * (muS / P.muS0) * (muS / P.muS0) + (muQ / P.muQ0) * (muQ / P.muQ0) * (muQ / P.muQ0) * (muQ / P.muQ0);
  r[TH_P] = cp * x;
  r[TH_S] = 4.0 * cp * T * T * T / T04;
  r[TH_B] = 4.0 * cp * muB * muB * muB / B04;
  r[TH_SS] = 4.0 * cp * muS * muS * muS / S04;
  r[TH_Q] = 4.0 * cp * muQ * muQ * muQ / Q04;
  r[TH_E] = (r[TH_S] * T + muB * r[TH_B] + muQ * r[TH_Q] + muS * r[TH_SS]) - r[TH_P];
  r[TH_CS2] = 1.0 / 3.0;
  r[TH_B2] = 12.0 * cp * muB * muB / B04; r[TH_Q2] = 12.0 * cp * muQ * muQ / Q04; r[TH_S2] = 12.0 * cp * muS * muS / S04;
  r[TH_BQ] = kTiny; r[TH_BS] = kTiny; r[TH_QS] = kTiny; r[TH_TB] = kTiny; r[TH_TQ] = kTiny; r[TH_TS] = kTiny;
  r[TH_T2] = 12.0 * cp * T * T / T04;
}
CCB_HD double sgn(double v) { return (double)((0.0 < v) - (v < 0.0)); }
// EoS_conformal_diagonal::get_tbqs_seed_from_{s,e}BSQ (include/eos_conformal_diagonal.h:48-76): also the exact inverse of this EoS
CCB_HD void diagonal_seed_s(const AnalyticParams& P, double s0, double B0, double S0, double Q0, double x[4]) {
  const double den = cbrt(4.0) * cbrt(P.c);
  x[0] = cbrt(P.T0 * P.T0 * P.T0 * P.T0 * s0) / den;
  x[1] = sgn(B0) * cbrt(P.muB0 * P.muB0 * P.muB0 * P.muB0 * fabs(B0)) / den;
  x[2] = sgn(Q0) * cbrt(P.muQ0 * P.muQ0 * P.muQ0 * P.muQ0 * fabs(Q0)) / den;
  x[3] = sgn(S0) * cbrt(P.muS0 * P.muS0 * P.muS0 * P.muS0 * fabs(S0)) / den;
}
CCB_HD void diagonal_seed_e(const AnalyticParams& P, double e0, double B0, double S0, double Q0, double x[4]) {
  const double den = cbrt(4.0) * cbrt(P.c);
  const double f43 = 4.0 / 3.0;
  const double disc = e0 - 3.0 * (pow(P.muB0 * fabs(B0), f43) + pow(P.muS0 * fabs(S0), f43) + pow(P.muQ0 * fabs(Q0), f43)) / (4.0 * den);
  x[0] = P.T0 * pow(disc / (3.0 * P.c), 0.25);  // NaN when the EoS has no solution (disc < 0), as in the reference
  x[1] = sgn(B0) * cbrt(P.muB0 * P.muB0 * P.muB0 * P.muB0 * fabs(B0)) / den;
  x[2] = sgn(Q0) * cbrt(P.muQ0 * P.muQ0 * P.muQ0 * P.muQ0 * fabs(Q0)) / den;
  x[3] = sgn(S0) * cbrt(P.muS0 * P.muS0 * P.muS0 * P.muS0 * fabs(S0)) / den;
}

// ---------------------------------------------------------------------------------------------------------------------
// Forward evaluation of member k of the chain.
CCB_HD bool chain_full_thermo(const EosChain& ch, int k, const double x[4], double r[TH_N]) {
  switch (ch.kind[k]) {
    case EOS_KIND_TABLE: return table_full_thermo(ch.table, x, r);
    case EOS_KIND_TANH: tanh_full_thermo(ch.tanh_p, x[0], x[1], x[2], x[3], r); return true;
    case EOS_KIND_CONFORMAL: {
      const ConformalParams C{ch.conf_p.c, ch.conf_p.T0, ch.conf_p.muB0, ch.conf_p.muS0, ch.conf_p.muQ0};
      conformal_full_thermo(C, x[0], x[1], x[2], x[3], r);
      return true;
    }
    default: diagonal_full_thermo(ch.diag_p, x[0], x[1], x[2], x[3], r); return true;
  }
}

// Residual densities (s|e, B, S, Q) and Jacobian d(dens)/dx of member k; analytic members: from the Hessian
// (d e / d x_j = sum_i x_i H_ij because dp/dx_j = n_j).
// TABLE (here and below): member k is known to be the table at compile time (the fast path of phase 04), so the analytic members'
// code -- most of the instructions of these functions -- is not generated.
template <bool TABLE = false>
CCB_HD bool chain_densities(const EosChain& ch, int k, bool energy, const double x[4], double val[4], double jac[4][4]) {
  if (TABLE || ch.kind[k] == EOS_KIND_TABLE) return table_densities(ch.table, energy ? ch.table.hot_e : ch.table.hot_s, x, val, jac);
  if (TABLE) return false;
  double r[TH_N];
  chain_full_thermo(ch, k, x, r);
  // Hessian rows in x order (T, B, Q, S)
  const double H[4][4] = {{r[TH_T2], r[TH_TB], r[TH_TQ], r[TH_TS]}, {r[TH_TB], r[TH_B2], r[TH_BQ], r[TH_BS]},
                          {r[TH_TQ], r[TH_BQ], r[TH_Q2], r[TH_QS]}, {r[TH_TS], r[TH_BS], r[TH_QS], r[TH_S2]}};
  val[0] = energy ? r[TH_E] : r[TH_S]; val[1] = r[TH_B]; val[2] = r[TH_SS]; val[3] = r[TH_Q];
  CCB_UNROLL for (int d = 0; d < 4; ++d) {
    jac[0][d] = energy ? (x[0] * H[0][d] + x[1] * H[1][d] + x[2] * H[2][d] + x[3] * H[3][d]) : H[0][d];
    jac[1][d] = H[1][d]; jac[2][d] = H[3][d]; jac[3][d] = H[2][d];  // val order is (.., B, S, Q)
  }
  return true;
}

// One run of the root search from x (Rootfinder::rootfinder4D, src/rootfinder.cpp:206-417).
// mode: 0 full, 1 T-only, 2 baryon-only (the pinned potentials are driven to zero, :131-160).
template <bool TABLE = false>
CCB_HD bool newton_solve(const EosChain& ch, int k, bool energy, int mode, const double target[4] /* s|e, B, S, Q */,
                         const double lo[4], const double hi[4], double x[4]) {
  double f[4], J[4][4], val[4], jac[4][4];
  double fnorm = 0.0;
  int clamped = 0;
  bool have = false;  // val / jac already hold the evaluation at x (the accepted trial point of the previous iterate)
  for (int it = 0; it <= kNewtonMaxIter; ++it) {
    if (!have && !chain_densities<TABLE>(ch, k, energy, x, val, jac)) return false;
    have = false;
    // residuals in the reference's order: (s, B, Q, S)
    f[0] = val[0] - target[0]; f[1] = val[1] - target[1]; f[2] = val[3] - target[3]; f[3] = val[2] - target[2];
    CCB_UNROLL for (int d = 0; d < 4; ++d) { J[0][d] = jac[0][d]; J[1][d] = jac[1][d]; J[2][d] = jac[3][d]; J[3][d] = jac[2][d]; }
    if (mode == 1) { CCB_UNROLL for (int r = 1; r < 4; ++r) { f[r] = x[r]; CCB_UNROLL for (int d = 0; d < 4; ++d) J[r][d] = (r == d) ? 1.0 : 0.0; } }
    else if (mode == 2) { CCB_UNROLL for (int r = 2; r < 4; ++r) { f[r] = x[r]; CCB_UNROLL for (int d = 0; d < 4; ++d) J[r][d] = (r == d) ? 1.0 : 0.0; } }
    fnorm = fabs(f[0]) + fabs(f[1]) + fabs(f[2]) + fabs(f[3]);
    if (!(fnorm == fnorm) || fnorm > 1e300) return false;  // GSL_EBADFUNC
    if (it > 0 && fnorm < kRootTolerance) return true;     // gsl_multiroot_test_residual after at least one iterate
    if (it == kNewtonMaxIter) return false;
    // Newton step: J dx = -f, elimination with partial pivoting
    double A[4][5];
    CCB_UNROLL for (int r = 0; r < 4; ++r) { CCB_UNROLL for (int d = 0; d < 4; ++d) A[r][d] = J[r][d]; A[r][4] = -f[r]; }
    double dx[4];
    if (!solve4_pivoted(A, dx)) return false;  // GSL_ENOPROGJ
    // A step that would leave the EoS's range is shortened to stop just inside it; an iterate that keeps pressing
    // against the boundary is the reference's "out-of-bounds" failure (src/rootfinder.cpp:283-337).
    double scale = 1.0;
    CCB_UNROLL for (int d = 0; d < 4; ++d) {
      const double xn = x[d] + dx[d];
      if (xn > hi[d]) scale = fmin(scale, 0.9 * (hi[d] - x[d]) / dx[d]);
      else if (xn < lo[d]) scale = fmin(scale, 0.9 * (lo[d] - x[d]) / dx[d]);
    }
    if (scale < 1.0) { if (++clamped > 12 || !(scale > 1e-6)) return false; } else clamped = 0;
    // backtracking on sum|f| (the interpolant is only piecewise smooth)
    // The trial point is evaluated WITH its Jacobian: when it is accepted it is the next iterate, and the same 16 vertex loads
    // would otherwise be repeated at the top of the loop (a third of all table evaluations of a converging search).
    double xt[4], ft = 0.0, v2[4], j2[4][4];
    bool accepted = false;
    for (int bt = 0; bt < 8; ++bt) {
      CCB_UNROLL for (int d = 0; d < 4; ++d) xt[d] = x[d] + scale * dx[d];
      const bool ok = chain_densities<TABLE>(ch, k, energy, xt, v2, j2);
      if (ok) {
        double g[4] = {v2[0] - target[0], v2[1] - target[1], v2[3] - target[3], v2[2] - target[2]};
        if (mode == 1) { g[1] = xt[1]; g[2] = xt[2]; g[3] = xt[3]; }
        else if (mode == 2) { g[2] = xt[2]; g[3] = xt[3]; }
        ft = fabs(g[0]) + fabs(g[1]) + fabs(g[2]) + fabs(g[3]);
        if (ft == ft && (ft < fnorm || ft < kRootTolerance)) { accepted = true; break; }
      }
      scale *= 0.5;
    }
    if (!accepted) return false;  // GSL_ENOPROG
    CCB_UNROLL for (int d = 0; d < 4; ++d) {
      x[d] = xt[d]; val[d] = v2[d];
      CCB_UNROLL for (int e = 0; e < 4; ++e) jac[d][e] = j2[d][e];
    }
    have = true;
  }
  return false;
}

CCB_HD bool in_range(const double x[4], const double lo[4], const double hi[4]) {
  CCB_UNROLL for (int d = 0; d < 4; ++d) if (x[d] < lo[d] || x[d] > hi[d] || x[d] != x[d]) return false;
  return true;
}

// Rootfinder::find_root (src/rootfinder.cpp:423-716): the seed, then the seed with T, muB, muS, muQ moved by
// +-20 %, then the "mu = 0" seed. `x` holds the seed on entry and the root on success.
// FIRST_ONLY: stop after the search from the seed itself (the fast path of phase 04: a particle that is on the table and stays on
// it converges there; everything else is left to the full chain, which repeats this first search with the same outcome).
template <bool FIRST_ONLY = false, bool TABLE = false>
CCB_HD bool find_root(const EosChain& ch, int k, bool energy, double t0_, double Bin, double Sin, double Qin, double x[4]) {
  double lo[4], hi[4];
  eos_bounds(ch, k, lo, hi);
  const double target[4] = {t0_, Bin, Sin, Qin};
  // closed-form members
  if (!TABLE && ch.kind[k] == EOS_KIND_CONFORMAL) {
    const ConformalParams C{ch.conf_p.c, ch.conf_p.T0, ch.conf_p.muB0, ch.conf_p.muS0, ch.conf_p.muQ0};
    double T, muB, muS, muQ;
    const bool ok = energy ? conformal_invert_e(C, t0_, Bin, Sin, Qin, T, muB, muS, muQ) : conformal_invert_s(C, t0_, Bin, Sin, Qin, T, muB, muS, muQ);
    if (!ok) return false;
    x[0] = T; x[1] = muB; x[2] = muQ; x[3] = muS;
    return in_range(x, lo, hi);
  }
  if (!TABLE && ch.kind[k] == EOS_KIND_DIAGONAL) {
    if (energy) diagonal_seed_e(ch.diag_p, t0_, Bin, Sin, Qin, x); else diagonal_seed_s(ch.diag_p, t0_, Bin, Sin, Qin, x);
    return in_range(x, lo, hi) && x[0] > 0.0;
  }
  const bool frozenB = fabs(hi[1] - lo[1]) < 1e-14, frozenQ = fabs(hi[2] - lo[2]) < 1e-14, frozenS = fabs(hi[3] - lo[3]) < 1e-14;
  const bool zB = fabs(Bin) < 1e-30, zQ = fabs(Qin) < 1e-30, zS = fabs(Sin) < 1e-30;
  const bool t_only = frozenB && frozenQ && frozenS && zB && zQ && zS;
  const bool b_only = !t_only && frozenQ && frozenS && zQ && zS;
  const int mode = t_only ? 1 : (b_only ? 2 : 0);
  double pos[4] = {x[0], x[1], x[2], x[3]};
  if (t_only) { pos[1] = pos[2] = pos[3] = 0.0; } else if (b_only) { pos[2] = pos[3] = 0.0; }
  const double sg[4] = {1.0, sgn(Bin), sgn(Qin), sgn(Sin)};
  CCB_UNROLL for (int d = 0; d < 4; ++d) if (pos[d] * pos[d] < kRootTolerance) pos[d] = fmin(0.5 * hi[d], 1.0) * sg[d];
  double y[4];
#define CCB_TRY()                                                                              \
  { CCB_UNROLL for (int d = 0; d < 4; ++d) y[d] = pos[d];                                      \
    if (newton_solve<TABLE>(ch, k, energy, mode, target, lo, hi, y)) { CCB_UNROLL for (int d = 0; d < 4; ++d) x[d] = y[d]; return true; } }
  CCB_TRY()
  if (FIRST_ONLY) return false;
  const double base[4] = {pos[0], pos[1], pos[2], pos[3]};
  // Rootfinder::tbqs keeps the previous position when the proposed one is out of range (src/rootfinder.cpp:172-203)
  // (arrays are addressed with compile-time indices only -- a run-time index would move them to local memory for good)
  auto at = [](const double* v, int d) { return d == 0 ? v[0] : (d == 1 ? v[1] : (d == 2 ? v[2] : v[3])); };
  auto propose = [&](int d, double v) {
    double q[4];
    CCB_UNROLL for (int e = 0; e < 4; ++e) q[e] = (e == d) ? v : base[e];
    if (in_range(q, lo, hi)) { CCB_UNROLL for (int e = 0; e < 4; ++e) pos[e] = q[e]; }
  };
  // order of the reference: T, muB, muS, muQ (indices 0, 1, 3, 2 of x)
  for (int a = 0; a < 4; ++a) {
    const int d = (a == 2) ? 3 : (a == 3 ? 2 : a);
    if (d == 1 && t_only) continue;
    if (d >= 2 && (t_only || b_only)) continue;
    const double bd = at(base, d), hd = at(hi, d), ld = at(lo, d);
    const double step = 0.2 * bd;
    propose(d, (bd + step > hd) ? hd - 1.0 : bd + step);
    CCB_TRY()
    propose(d, (bd - step < ld) ? ld + 1.0 : bd - step);
    CCB_TRY()
  }
  {
    double q[4] = {base[0], 0.0, 0.0, 0.0};
    if (!t_only) q[1] = fmin(0.5 * hi[1], 1.0) * sgn(Bin);
    if (!t_only && !b_only) { q[2] = fmin(0.5 * hi[2], 1.0) * sgn(Qin); q[3] = fmin(0.5 * hi[3], 1.0) * sgn(Sin); }
    if (in_range(q, lo, hi)) { CCB_UNROLL for (int e = 0; e < 4; ++e) pos[e] = q[e]; }
    CCB_TRY()
  }
#undef CCB_TRY
  return false;
}

// EquationOfState::tbqs(): derivative mode from the frozen axes of the EoS alone (src/eos.cpp:132-144)
CCB_HD int derivative_mode(const EosChain& ch, int k) {
  double lo[4], hi[4];
  eos_bounds(ch, k, lo, hi);
  const bool fB = fabs(lo[1] - hi[1]) < 1e-30, fQ = fabs(lo[2] - hi[2]) < 1e-30, fS = fabs(lo[3] - hi[3]) < 1e-30;
  if (fB && fQ && fS) return 1;
  if (fQ && fS) return 2;
  return 0;
}

struct ChainResult {
  double x[4];     // T, muB, muQ, muS
  double r[TH_N];  // forward thermodynamics at x
  double chiBB0;
  int eos;         // member of the chain that produced it
};

// EquationOfState::evaluate_thermodynamics (src/eos.cpp:174-221)
template <bool TABLE = false>
CCB_HD void chain_evaluate(const EosChain& ch, int k, const double x[4], ChainResult& out) {
  CCB_UNROLL for (int d = 0; d < 4; ++d) out.x[d] = x[d];
  if (TABLE) table_full_thermo(ch.table, x, out.r); else chain_full_thermo(ch, k, x, out.r);
  out.chiBB0 = out.r[TH_B2];
  if ((TABLE || ch.kind[k] == EOS_KIND_TABLE) && (x[1] != 0.0 || x[2] != 0.0 || x[3] != 0.0)) {
    const double x0[4] = {x[0], 0.0, 0.0, 0.0};
    out.chiBB0 = table_field(ch.table, x0, TH_B2);
  }
  out.eos = k;
}

CCB_HD bool chain_accepts(const EosChain& ch, int k, const ChainResult& res) {
  const double cs2 = res.r[TH_CS2];
  if (ch.prohibit_unstable_cs2 && cs2 < 0.0) return false;
  if (ch.prohibit_acausal_cs2 && cs2 > 1.0) return false;
  if (ch.restrict_mu_T_ratios && ch.kind[k] == EOS_KIND_TABLE &&
      fmax(fmax(fabs(res.x[1]), fabs(res.x[3])), fabs(res.x[2])) > 3.5 * res.x[0]) return false;
  return true;
}

// EquationOfState::rootfinder_update_s (src/eos.cpp:638-804). `pos` = the particle's previous point, `start` = the
// member it last used (members before it are skipped: skip_failed_EoS, include/eos_header.h:8).
CCB_HD bool chain_update_s(const EosChain& ch, int start, const double pos_in[4], double sin_, double Bin, double Sin, double Qin,
                           ChainResult& out) {
  const double hc = kHbarcMeVfm;
  double pos[4] = {pos_in[0], pos_in[1], pos_in[2], pos_in[3]};
  bool found = false;
  for (int k = start < 0 ? 0 : start; k < ch.n; ++k) {
    double x[4];
    found = false;
    for (int seed = 0; seed < 4 && !found; ++seed) {
      if (seed == 0) { CCB_UNROLL for (int d = 0; d < 4; ++d) x[d] = pos[d]; }
      else if (seed == 1) { x[0] = pos[0]; x[1] = x[2] = x[3] = 0.0; }
      else if (seed == 2) { x[0] = 580.0 / hc; x[1] = sgn(Bin); x[2] = sgn(Qin); x[3] = sgn(Sin); }
      else diagonal_seed_s(ch.diag_p, sin_, Bin, Sin, Qin, x);
      found = find_root(ch, k, false, sin_, Bin, Sin, Qin, x);
    }
    if (found) {
      chain_evaluate(ch, k, x, out);
      CCB_UNROLL for (int d = 0; d < 4; ++d) pos[d] = x[d];
      if (chain_accepts(ch, k, out)) break;
    }
  }
  return found;
}

// EquationOfState::rootfinder_s_out (src/eos.cpp:834-977): every member is tried, forced seed first.
CCB_HD bool chain_s_out(const EosChain& ch, double ein, double Bin, double Sin, double Qin, ChainResult& out) {
  const double hc = kHbarcMeVfm;
  bool found = false;
  for (int k = 0; k < ch.n; ++k) {
    double x[4] = {580.0 / hc, sgn(Bin), sgn(Qin), sgn(Sin)};
    found = find_root(ch, k, true, ein, Bin, Sin, Qin, x);
    if (!found) {
      diagonal_seed_e(ch.diag_p, ein, Bin, Sin, Qin, x);
      found = find_root(ch, k, true, ein, Bin, Sin, Qin, x);
    }
    if (found) {
      chain_evaluate(ch, k, x, out);
      if (chain_accepts(ch, k, out)) break;
    }
  }
  return found;
}

// EquationOfState::set_thermo for a chain result
CCB_HD void chain_thermo(const EosChain& ch, const ChainResult& res, Thermo& th) {
  thermo_from_forward(res.x[0], res.x[1], res.x[2], res.x[3], res.r, derivative_mode(ch, res.eos), th);
  th.chiBB0 = (ch.kind[res.eos] == EOS_KIND_TABLE) ? res.chiBB0 : th.chiBB;
}

}  // namespace ccb
