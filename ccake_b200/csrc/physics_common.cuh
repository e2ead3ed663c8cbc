// Per-pair and per-particle physics of the CCAKE SPH hot path, written as plain
// inline functions over registers so that the CUDA kernels can fuse what the
// reference runs as ~22 separate Kokkos launches.
//
// Everything here is fp64. Functions are __host__ __device__ only so that
// tests/host_harness.cpp can exercise the same arithmetic on the CPU against the
// oracle; the shipped library calls them from device code exclusively.
#pragma once

#include <cmath>

#if defined(__CUDACC__)
#define CCB_HD __host__ __device__ __forceinline__
#define CCB_UNROLL _Pragma("unroll")
#else
#define CCB_HD inline
#define CCB_UNROLL
#endif

namespace ccb {

constexpr double kHbarcGeVfm = 0.1973269804;  // reference include/constants.h:15
constexpr double kHbarcMeVfm = 197.3269804;   // reference include/constants.h:16
constexpr double kTiny = 1e-25;               // reference include/eos_header.h:16
constexpr double kPi = 3.14159265358979323846;

// ---------------------------------------------------------------------------
// Cubic-spline SPH kernel (reference src/kernel.cpp:11-118).
// The four normalisations are divided by h^D / h^(D+1) once on the host (the
// reference re-evaluates pow(h, D) for every pair).
struct KernelNorm {
  double norm;        // knorm  / h^D      (q < 1 branch of W)
  double norm2;       // knorm2 / h^D      (1 <= q < 2 branch of W)
  double norm_grad;   // kgrad  / h^(D+1)  (1 <= q < 2 branch of grad W)
  double norm_grad2;  // kgrad2 / h^(D+1)  (q < 1 branch of grad W)
  double h;
  double inv_h;
};

inline KernelNorm make_kernel_norm(int D, double h) {
  double knorm, knorm2, kgrad, kgrad2;
  if (D == 1) { knorm = 2. / 3.; knorm2 = 1. / 6.; kgrad = -1. / 2.; kgrad2 = 1.; }
  else if (D == 2) { knorm = 10. / (7 * kPi); knorm2 = 5. / (14. * kPi); kgrad = -15. / (14. * kPi); kgrad2 = 15. / (7. * kPi); }
  else { knorm = 1. / kPi; knorm2 = 1. / (4 * kPi); kgrad = -3. / (4. * kPi); kgrad2 = 3. / (2. * kPi); }
  KernelNorm k;
  k.norm = knorm / std::pow(h, D);
  k.norm2 = knorm2 / std::pow(h, D);
  k.norm_grad = kgrad / std::pow(h, D + 1);
  k.norm_grad2 = kgrad2 / std::pow(h, D + 1);
  k.h = h;
  k.inv_h = 1.0 / h;
  return k;
}

// W(r, h), reference src/kernel.cpp:36-48
CCB_HD double sph_w(const KernelNorm& k, double r) {
  const double q = r / k.h;
  if (q >= 2.0) return 0.0;
  if (q >= 1.0) return k.norm2 * (2.0 - q) * (2.0 - q) * (2.0 - q);
  const double qq = q * q;
  return k.norm * (1 - 1.5 * qq + 0.75 * q * qq);
}

// Scalar factor f such that grad_a W = f * (r_a - r_b), reference src/kernel.cpp:97-118.
// (q < 1: norm_grad2 * (1.5 q - 2) / h ; 1 <= q < 2: (norm_grad / r) * (2 - q)^2 ; else 0.)
CCB_HD double sph_gradw_factor(const KernelNorm& k, double r) {
  const double q = r / k.h;
  if (q >= 2.0) return 0.0;
  if (q >= 1.0) return (k.norm_grad / r) * (2.0 - q) * (2.0 - q);
  return k.norm_grad2 * (1.5 * q - 2) / k.h;
}

// ---------------------------------------------------------------------------
// Metric policy. MILNE: g = diag(1,-1,-1,-t^2); otherwise Minkowski.
// The reference has one 1227-line header per metric (include/milne.hpp,
// include/cartesian.hpp) differing only in the factor carried by the last slot.
template <bool MILNE> struct Metric {
  double t, t2, g33;  // g33 = t^2 (Milne) or 1 (Cartesian): |g_{eta eta}|
  CCB_HD explicit Metric(double t_) : t(t_), t2(t_ * t_), g33(MILNE ? t_ * t_ : 1.0) {}
};

// u^i -> u_i for a D-vector (reference milne.hpp:92-95; D==2 specialisation :353-354
// flips both signs and never applies t^2).
template <int D, bool MILNE> CCB_HD void lower_vec(const Metric<MILNE>& m, const double* in, double* out) {
  if (D == 2) { out[0] = -in[0]; out[1] = -in[1]; }
  else {
    CCB_UNROLL for (int i = 0; i < D - 1; ++i) out[i] = -in[i];
    out[D - 1] = in[D - 1] * (-m.g33);
  }
}
// fixed-size 3-vector: always (-,-,-g33)
template <bool MILNE> CCB_HD void lower_vec3(const Metric<MILNE>& m, const double* in, double* out) {
  out[0] = -in[0]; out[1] = -in[1]; out[2] = in[2] * (-m.g33);
}
// v_i -> v^i (reference milne.hpp:103-107, :359-360)
template <int D, bool MILNE> CCB_HD void raise_vec(const Metric<MILNE>& m, const double* in, double* out) {
  if (D == 2) { out[0] = -in[0]; out[1] = -in[1]; }
  else {
    CCB_UNROLL for (int i = 0; i < D - 1; ++i) out[i] = -in[i];
    out[D - 1] = in[D - 1] * (-1. / m.g33);
  }
}
// contravariant metric diagonal used by the shear terms: get_contra_metric_diagonal<3>(t)
// (reference milne.hpp:1130-1138): (1, -1, -1, -1/t/t)
template <bool MILNE> CCB_HD void contra_metric_diag(const Metric<MILNE>& m, double* g) {
  g[0] = 1.; g[1] = -1.; g[2] = -1.; g[3] = MILNE ? -1. / m.t / m.t : -1.;
}

// gamma = sqrt(1 - u_i u^i), reference src/eom_default.cpp:24-37
template <int D, bool MILNE> CCB_HD double gamma_of_u(const Metric<MILNE>& m, const double* u) {
  double uc[D];
  lower_vec<D, MILNE>(m, u, uc);
  double c = 0.0;
  CCB_UNROLL for (int i = 0; i < D; ++i) c += uc[i] * u[i];
  return sqrt(1 - c);
}

// lab frame -> local rest frame, reference src/eom_default.cpp:199-203 (Milne: lab/gamma/t),
// src/eom_cartesian.cpp:197-202 (lab/gamma)
template <bool MILNE> CCB_HD double to_lrf(const Metric<MILNE>& m, double lab, double gamma) {
  return MILNE ? lab / gamma / m.t : lab / gamma;
}

// 3x3 inverse by cofactors, same expression order as reference milne.hpp:810-827
CCB_HD void inverse3(const double A[3][3], double Ai[3][3]) {
  const double det = A[0][0] * A[1][1] * A[2][2] + A[0][1] * A[1][2] * A[2][0] + A[0][2] * A[1][0] * A[2][1] -
                     A[0][2] * A[1][1] * A[2][0] - A[0][1] * A[1][0] * A[2][2] - A[0][0] * A[1][2] * A[2][1];
  const double id = 1.0 / det;  // one division instead of the reference's nine (<= 1 ulp per entry)
  Ai[0][0] = (A[1][1] * A[2][2] - A[1][2] * A[2][1]) * id;
  Ai[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * id;
  Ai[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * id;
  Ai[1][0] = (A[1][2] * A[2][0] - A[1][0] * A[2][2]) * id;
  Ai[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * id;
  Ai[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * id;
  Ai[2][0] = (A[1][0] * A[2][1] - A[1][1] * A[2][0]) * id;
  Ai[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * id;
  Ai[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * id;
}
CCB_HD void inverse2(const double A[2][2], double Ai[2][2]) {  // reference milne.hpp:797-808
  const double det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
  const double id = 1.0 / det;
  Ai[0][0] = A[1][1] * id; Ai[0][1] = -A[0][1] * id; Ai[1][0] = -A[1][0] * id; Ai[1][1] = A[0][0] * id;
}

// ---------------------------------------------------------------------------
// pi^{mu nu}, Pi and q^mu from the evolved extensive components
// (reference src/eom_default.cpp:57-172, src/eom_cartesian.cpp:57-170).
// ext_shv = (xx, xy, x-eta, yy, y-eta); shv is the full symmetric 4x4, index 0=tau,1=x,2=y,3=eta.
template <int D, bool MILNE>
CCB_HD void rebuild_shear(const Metric<MILNE>& m, const double* u, double gamma, double sigma, const double* ext_shv,
                          double shv[4][4]) {
  double fu[3] = {0., 0., 0.}, fuc[3];
  CCB_UNROLL for (int i = 0; i < D; ++i) fu[i] = u[i];
  lower_vec3<MILNE>(m, fu, fuc);
  shv[1][1] = ext_shv[0] * sigma; shv[1][2] = ext_shv[1] * sigma; shv[1][3] = ext_shv[2] * sigma;
  shv[2][2] = ext_shv[3] * sigma; shv[2][3] = ext_shv[4] * sigma;
  shv[2][1] = shv[1][2]; shv[3][1] = shv[1][3]; shv[3][2] = shv[2][3];
  // pi^00 = u_i u_j pi^ij / gamma^2 with pi^33 eliminated through tracelessness
  if (MILNE) {
    shv[0][0] = (fuc[0] * fuc[0] * shv[1][1] + fuc[1] * fuc[1] * shv[2][2] + 2. * fuc[1] * fuc[0] * shv[2][1] +
                 2. * fuc[2] * fuc[0] * shv[3][1] + 2. * fuc[2] * fuc[1] * shv[3][2] -
                 fuc[2] * fuc[2] * (shv[1][1] + shv[2][2]) / m.t2) /
                (gamma * gamma - fuc[2] * fuc[2] / (m.t2));
    shv[3][3] = (shv[0][0] - shv[1][1] - shv[2][2]) / m.t2;
  } else {
    shv[0][0] = (fuc[0] * fuc[0] * shv[1][1] + fuc[1] * fuc[1] * shv[2][2] + 2. * fuc[1] * fuc[0] * shv[2][1] +
                 2. * fuc[2] * fuc[0] * shv[3][1] + 2. * fuc[2] * fuc[1] * shv[3][2] -
                 fuc[2] * fuc[2] * (shv[1][1] + shv[2][2])) /
                (gamma * gamma - fuc[2] * fuc[2]);
    shv[3][3] = (shv[0][0] - shv[1][1] - shv[2][2]);
  }
  const double ig = 1.0 / gamma;
  CCB_UNROLL for (int i = 0; i < 3; ++i) {
    double p = 0.;
    CCB_UNROLL for (int j = 0; j < 3; ++j) p += -shv[i + 1][j + 1] * fuc[j] * ig;
    shv[0][i + 1] = p;
    shv[i + 1][0] = p;
  }
}

// q^mu_a (a = B,S,Q) from extensive diffusion; q^0 = -q^i u_i / gamma
template <int D, bool MILNE>
CCB_HD void rebuild_diffusion(const Metric<MILNE>& m, const double* u, double gamma, double sigma, const double* ext_q,
                              double q[3][4]) {
  double fu[3] = {0., 0., 0.}, fuc[3];
  CCB_UNROLL for (int i = 0; i < D; ++i) fu[i] = u[i];
  lower_vec3<MILNE>(m, fu, fuc);
  const double ig = 1.0 / gamma;
  CCB_UNROLL for (int c = 0; c < 3; ++c) {
    double q0 = 0.;
    CCB_UNROLL for (int i = 0; i < 3; ++i) {
      q[c][i + 1] = ext_q[c * 3 + i] * sigma;
      q0 += -q[c][i + 1] * fuc[i] * ig;
    }
    q[c][0] = q0;
  }
}

// SPHWorkstation::regulate_diffusion (src/sph_workstation.cpp:133-218; PhysRevC.98.034916 Eqs. C6-C8): where the spacelike
// norm |q_a| exceeds rq_max |rho_a| chi0 (F(e) - F(0)), F the Fermi function of (e - e0) / xi0, the current q^mu_a is
// scaled back onto that bound and extensive_diffusion is made consistent with it. Returns true if anything changed.
CCB_HD bool regulate_diffusion_current(bool milne, double t, double chi0, double e0, double xi0, double rq_max, double e, double sigma,
                                       const double rho[3], double q[3][4], double ext_q[9]) {
  const double fermi_offset = 1.0 / (exp(e0 / xi0) + 1.0);
  const double t2 = t * t;
  const double ferm = 1.0 / (exp(-(e - e0) / xi0) + 1.0);
  const double f_str = chi0 * (ferm - fermi_offset);
  bool changed = false;
  CCB_UNROLL for (int c = 0; c < 3; ++c) {
    const double n_abs = fabs(rho[c]);
    if (n_abs < 1.0e-15) continue;
    const double q0 = q[c][0], q1 = q[c][1], q2 = q[c][2], q3 = q[c][3];
    const double q3_sq = milne ? t2 * q3 * q3 : q3 * q3;
    const double q_sq = q1 * q1 + q2 * q2 + q3_sq - q0 * q0;
    if (q_sq <= 0.0) continue;
    const double q_mag = sqrt(q_sq);
    const double threshold = rq_max * n_abs * f_str;
    if (q_mag <= threshold) continue;
    const double scale = threshold / q_mag;
    q[c][0] = q0 * scale; q[c][1] = q1 * scale; q[c][2] = q2 * scale; q[c][3] = q3 * scale;
    CCB_UNROLL for (int d = 0; d < 3; ++d) ext_q[c * 3 + d] = q[c][d + 1] / sigma;
    changed = true;
  }
  return changed;
}

}  // namespace ccb
