// Per-particle kernels: the work the reference spreads over ~18 launches per RK
// stage collapsed into four streaming kernels:
//   prologue      : 02 gamma/v (src/eom_default.cpp:18-38) + packing of the field-gather records
//   eos_local     : 04 EoS inversion (src/sph_workstation.cpp:1714-1772; serial host code in the
//                   reference), 05 pi^{mu nu}/q^mu rebuild (src/eom_default.cpp:57-172),
//                   08 transport coefficients (src/sph_workstation.cpp:1623-1699) + packing of
//                   the gradient-gather records
//   eom_rk        : 10 equations of motion (src/eom_default.cpp:212-869) fused with the
//                   Runge-Kutta stage update of Evolver::step_rk (src/evolver.cpp:345-485)
//   reduce_totals : conservation sums (src/system_state.cpp:739-885) in one pass
// One thread per particle, every column is read or written once: HBM-bound.
#pragma once

#include "gather_params.cuh"
#include "stage_params.h"

namespace ccb {

template <int D, bool MILNE>
__global__ void prologue_kernel(DeviceView dv, double t) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dv.n) return;
  Metric<MILNE> m(t);
  double u[D], x[3];
#pragma unroll
  for (int d = 0; d < D; ++d) u[d] = col(dv.f[CCB_F_U], dv.cap, d, i);
#pragma unroll
  for (int d = 0; d < 3; ++d) x[d] = col(dv.f[CCB_F_POSITION], dv.cap, d, i);
  const double gamma = gamma_of_u<D, MILNE>(m, u);
  dv.f[CCB_F_GAMMA][i] = gamma;
#pragma unroll
  for (int d = 0; d < 3; ++d) col(dv.f[CCB_F_V], dv.cap, d, i) = d < D ? u[d < D ? d : 0] / gamma : 0.0;
  double* r = dv.rec_fields + i * REC_F;
  const double ms = col(dv.f[CCB_F_SPH_MASS], dv.cap, 0, i);
  double2* q = reinterpret_cast<double2*>(r);
  q[0] = make_double2(x[0], x[1]);
  q[1] = make_double2(x[2], ms);
  q[2] = make_double2(ms * col(dv.f[CCB_F_EXTENSIVE], dv.cap, 0, i),
                      col(dv.f[CCB_F_SPH_MASS], dv.cap, 1, i) * col(dv.f[CCB_F_EXTENSIVE], dv.cap, 1, i));
  q[3] = make_double2(col(dv.f[CCB_F_SPH_MASS], dv.cap, 2, i) * col(dv.f[CCB_F_EXTENSIVE], dv.cap, 2, i),
                      col(dv.f[CCB_F_SPH_MASS], dv.cap, 3, i) * col(dv.f[CCB_F_EXTENSIVE], dv.cap, 3, i));
  // a run without conserved charges (all three *_charge_enabled switches off) smooths exact zeros: tell the field
  // gather when that is NOT the case
  if (q[2].y != 0.0 || q[3].x != 0.0 || q[3].y != 0.0) dv.qflag[2] = 1;
}

__device__ __forceinline__ void raise_error(DeviceView& dv, int code, int id) {
  if (atomicCAS(&dv.err[0], 0, code) == 0) dv.err[1] = id;
}

// Read of an input column that is streamed once (or re-read from L2): ld.global.cg, cached in L2 only. The per-particle kernels
// with a register-heavy body (eom_rk_kernel: 254 registers + 672 B of stack) keep their stack in L1 this way instead of having it
// evicted by ~130 input columns passing through.
__device__ __forceinline__ double colr(const double* p, int64_t cap, int c, int64_t i) { return __ldcg(p + (int64_t)c * cap + i); }

__device__ __forceinline__ void load_thermo(const DeviceView& dv, int64_t i, Thermo& th) {
  double* T = dv.f[CCB_F_THERMO];
  th.T = colr(T, dv.cap, 0, i); th.muB = colr(T, dv.cap, 1, i); th.muS = colr(T, dv.cap, 2, i); th.muQ = colr(T, dv.cap, 3, i);
  th.e = colr(T, dv.cap, 4, i); th.s = colr(T, dv.cap, 5, i); th.rhoB = colr(T, dv.cap, 6, i); th.rhoS = colr(T, dv.cap, 7, i);
  th.rhoQ = colr(T, dv.cap, 8, i); th.p = colr(T, dv.cap, 9, i); th.cs2 = colr(T, dv.cap, 10, i); th.w = colr(T, dv.cap, 11, i);
  th.dwds = colr(T, dv.cap, 12, i); th.dwdB = colr(T, dv.cap, 13, i); th.dwdS = colr(T, dv.cap, 14, i); th.dwdQ = colr(T, dv.cap, 15, i);
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int x = 0; x < 4; ++x) th.dalpha[a][x] = colr(dv.f[CCB_F_DALPHA], dv.cap, a * 4 + x, i);
  th.chiBB = colr(dv.f[CCB_F_CHIBB], dv.cap, 0, i); th.chiBB0 = colr(dv.f[CCB_F_CHIBB], dv.cap, 1, i);
}
__device__ __forceinline__ void store_thermo(const DeviceView& dv, int64_t i, const Thermo& th) {
  double* T = dv.f[CCB_F_THERMO];
  col(T, dv.cap, 0, i) = th.T; col(T, dv.cap, 1, i) = th.muB; col(T, dv.cap, 2, i) = th.muS; col(T, dv.cap, 3, i) = th.muQ;
  col(T, dv.cap, 4, i) = th.e; col(T, dv.cap, 5, i) = th.s; col(T, dv.cap, 6, i) = th.rhoB; col(T, dv.cap, 7, i) = th.rhoS;
  col(T, dv.cap, 8, i) = th.rhoQ; col(T, dv.cap, 9, i) = th.p; col(T, dv.cap, 10, i) = th.cs2; col(T, dv.cap, 11, i) = th.w;
  col(T, dv.cap, 12, i) = th.dwds; col(T, dv.cap, 13, i) = th.dwdB; col(T, dv.cap, 14, i) = th.dwdS; col(T, dv.cap, 15, i) = th.dwdQ;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int x = 0; x < 4; ++x) col(dv.f[CCB_F_DALPHA], dv.cap, a * 4 + x, i) = th.dalpha[a][x];
  col(dv.f[CCB_F_CHIBB], dv.cap, 0, i) = th.chiBB; col(dv.f[CCB_F_CHIBB], dv.cap, 1, i) = th.chiBB0;
  col(dv.f[CCB_F_TSEED], dv.cap, 0, i) = th.T; col(dv.f[CCB_F_TSEED], dv.cap, 1, i) = th.muB;
  col(dv.f[CCB_F_TSEED], dv.cap, 2, i) = th.muS; col(dv.f[CCB_F_TSEED], dv.cap, 3, i) = th.muQ;
}

// (s, rhoB, rhoS, rhoQ) in the local rest frame -> full thermodynamics.
// Reference: SPHWorkstation::locate_phase_diagram_point_sBSQ (src/sph_workstation.cpp:659-741)
// -> EquationOfState::update_s (src/eos.cpp:574-804) -> set_thermo (include/eos.h:121-148).
__device__ __forceinline__ bool eos_from_sBSQ(const StageParams& sp, double s, double B, double S, double Q, Thermo& th) {
  double T, muB, muS, muQ;
  if (!conformal_invert_s(sp.conf, s, B, S, Q, T, muB, muS, muQ)) return false;
  double r[TH_N];
  conformal_full_thermo(sp.conf, T, muB, muQ, muS, r);
  thermo_from_forward(T, muB, muQ, muS, r, 0, th);
  return true;
}

// One record of the gradient gather (layout: GradRec<D, DIFF>, gather_params.cuh), written with 128-bit stores.
template <int D, bool DIFF>
__device__ __forceinline__ void pack_grad_record(double* __restrict__ rec, int64_t i, const double* x, const double* v, double m_s,
                                                 double sig_lab, double bulk, const Thermo& th, const double shv[4][4],
                                                 const double q[3][4], int* __restrict__ qflag) {
  using R = GradRec<D, DIFF>;
  double r[R::REC];
#pragma unroll
  for (int k = 0; k < R::REC; ++k) r[k] = 0.0;
  const double sigsqr = 1.0 / (sig_lab * sig_lab);
#pragma unroll
  for (int d = 0; d < D; ++d) { r[R::X + d] = x[d]; r[R::V + d] = v[d]; r[R::S0 + d] = sigsqr * shv[0][d + 1]; }
  r[R::M] = m_s;
  r[R::PT] = sigsqr * th.p;
  r[R::BT] = bulk / sig_lab / sig_lab;
  r[R::E] = th.e;
#pragma unroll
  for (int d = 0; d < D; ++d)
#pragma unroll
    for (int e = d; e < D; ++e) r[R::SM + R::sym(d, e)] = sigsqr * shv[d + 1][e + 1];
  r[R::SIGLAB] = sig_lab;
  if (DIFF) {
    r[R::ALPHA + 0] = th.muB / th.T; r[R::ALPHA + 1] = th.muS / th.T; r[R::ALPHA + 2] = th.muQ / th.T;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      r[R::Q0 + c] = sigsqr * q[c][0];
#pragma unroll
      for (int d = 0; d < D; ++d) r[R::QI + c * D + d] = sigsqr * q[c][d + 1];
    }
    // zero diffusion currents everywhere (kappa = 0 and no initial diffusion, as in the reference's own inputs) let the
    // gradient gather skip the q^mu_a part of every record: tell it when that is not the case
    bool nz = false;
#pragma unroll
    for (int k = R::Q0; k < R::FULL; ++k) nz |= (r[k] != 0.0);
    if (nz) *qflag = 1;
  }
  double2* o = reinterpret_cast<double2*>(rec + i * R::REC);
#pragma unroll
  for (int k = 0; k < R::REC / 2; ++k) o[k] = make_double2(r[2 * k], r[2 * k + 1]);
}

template <int D, bool MILNE>
__global__ void eos_local_kernel(DeviceView dv, StageParams sp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dv.n) return;
  Metric<MILNE> m(sp.t);
  Thermo th;
  if (sp.do_eos) {
    const double s = col(dv.f[CCB_F_THERMO], dv.cap, 5, i), B = col(dv.f[CCB_F_THERMO], dv.cap, 6, i),
                 S = col(dv.f[CCB_F_THERMO], dv.cap, 7, i), Q = col(dv.f[CCB_F_THERMO], dv.cap, 8, i);
    if (!eos_from_sBSQ(sp, s, B, S, Q, th)) { raise_error(dv, CCB_ERR_EOS_NO_ROOT, dv.id[i]); return; }
    store_thermo(dv, i, th);
  } else {
    load_thermo(dv, i, th);
  }
  if (!sp.do_local) return;
  double u[D];
#pragma unroll
  for (int d = 0; d < D; ++d) u[d] = col(dv.f[CCB_F_U], dv.cap, d, i);
  const double gamma = dv.f[CCB_F_GAMMA][i], sigma = dv.f[CCB_F_SIGMA][i], sig_lab = dv.f[CCB_F_SIGMA_LAB][i];
  const double bulk = dv.f[CCB_F_EXTENSIVE_BULK][i] * sigma;
  dv.f[CCB_F_BULK][i] = bulk;
  double ext_shv[5], shv[4][4], ext_q[9], q[3][4];
#pragma unroll
  for (int c = 0; c < 5; ++c) ext_shv[c] = col(dv.f[CCB_F_EXTENSIVE_SHV], dv.cap, c, i);
  rebuild_shear<D, MILNE>(m, u, gamma, sigma, ext_shv, shv);
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) col(dv.f[CCB_F_SHV], dv.cap, a * 4 + b, i) = shv[a][b];
#pragma unroll
  for (int c = 0; c < 9; ++c) ext_q[c] = sp.fl.diffusion ? col(dv.f[CCB_F_EXTENSIVE_DIFFUSION], dv.cap, c, i) : 0.0;
  rebuild_diffusion<D, MILNE>(m, u, gamma, sigma, ext_q, q);
  if (sp.reg_enabled) {  // phase 06
    const double rho[3] = {th.rhoB, th.rhoS, th.rhoQ};
    if (regulate_diffusion_current(MILNE, sp.t, sp.reg_chi0, sp.reg_e0, sp.reg_xi0, sp.reg_rq_max, th.e, sigma, rho, q, ext_q)) {
#pragma unroll
      for (int c = 0; c < 9; ++c) col(dv.f[CCB_F_EXTENSIVE_DIFFUSION], dv.cap, c, i) = ext_q[c];
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int mu = 0; mu < 4; ++mu) col(dv.f[CCB_F_DIFFUSION], dv.cap, c * 4 + mu, i) = q[c][mu];
  Transport tc;
  transport_coefficients(th, sp.tp, tc);
  {
    double* TC = dv.f[CCB_F_TRANSPORT];
    const double v[13] = {tc.eta_pi, tc.tau_pi, tc.zeta_Pi, tc.tau_Pi, tc.delta_PiPi, tc.delta_pipi, tc.lambda_piPi,
                          tc.lambda_Pipi, tc.phi1, tc.phi3, tc.phi6, tc.phi7, tc.tau_pipi};
#pragma unroll
    for (int c = 0; c < 13; ++c) col(TC, dv.cap, c, i) = v[c];
  }
  // gradient-gather record, pre-scaled by 1/sigma_lab^2 where the pair formula needs it
  double x[3], v[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) { x[d] = col(dv.f[CCB_F_POSITION], dv.cap, d, i); v[d] = col(dv.f[CCB_F_V], dv.cap, d, i); }
  const double m_s = col(dv.f[CCB_F_SPH_MASS], dv.cap, 0, i);
  if (sp.fl.diffusion) pack_grad_record<D, true>(dv.rec_grad, i, x, v, m_s, sig_lab, bulk, th, shv, q, dv.qflag);
  else pack_grad_record<D, false>(dv.rec_grad, i, x, v, m_s, sig_lab, bulk, th, shv, q, dv.qflag);
}

// the 13 transport scalars phase 08 stored (column order of CCB_F_TRANSPORT, include/ccake_b200.h)
__device__ __forceinline__ void load_transport(const DeviceView& dv, int64_t i, Transport& tc) {
  const double* TC = dv.f[CCB_F_TRANSPORT];
  tc.eta_pi = colr(TC, dv.cap, 0, i); tc.tau_pi = colr(TC, dv.cap, 1, i); tc.zeta_Pi = colr(TC, dv.cap, 2, i); tc.tau_Pi = colr(TC, dv.cap, 3, i);
  tc.delta_PiPi = colr(TC, dv.cap, 4, i); tc.delta_pipi = colr(TC, dv.cap, 5, i); tc.lambda_piPi = colr(TC, dv.cap, 6, i);
  tc.lambda_Pipi = colr(TC, dv.cap, 7, i); tc.phi1 = colr(TC, dv.cap, 8, i); tc.phi3 = colr(TC, dv.cap, 9, i); tc.phi6 = colr(TC, dv.cap, 10, i);
  tc.phi7 = colr(TC, dv.cap, 11, i); tc.tau_pipi = colr(TC, dv.cap, 12, i);
}

// inputs of the equations of motion for particle i, from the SoA columns
template <int D>
__device__ __forceinline__ void load_eom_inputs(const DeviceView& dv, const StageParams& sp, int64_t i, EomIn<D>& in) {
#pragma unroll
  for (int d = 0; d < D; ++d) {
    in.u[d] = colr(dv.f[CCB_F_U], dv.cap, d, i);
    in.v[d] = colr(dv.f[CCB_F_V], dv.cap, d, i);
    in.gradP[d] = colr(dv.f[CCB_F_GRADP], dv.cap, d, i);
    in.gradBulk[d] = colr(dv.f[CCB_F_GRADBULK], dv.cap, d, i);
    in.gradshear[d] = colr(dv.f[CCB_F_GRADSHEAR], dv.cap, d, i);
    in.divshear[d] = colr(dv.f[CCB_F_DIVSHEAR], dv.cap, d, i);
#pragma unroll
    for (int e = 0; e < D; ++e) in.gradV[d][e] = colr(dv.f[CCB_F_GRADV], dv.cap, d * 3 + e, i);
#pragma unroll
    for (int c = 0; c < 3; ++c) in.grad_alpha[d][c] = sp.fl.diffusion ? colr(dv.f[CCB_F_GRAD_ALPHA], dv.cap, d * 3 + c, i) : 0.0;
  }
  in.gamma = colr(dv.f[CCB_F_GAMMA], dv.cap, 0, i);
  in.sigma = colr(dv.f[CCB_F_SIGMA], dv.cap, 0, i);
  in.sigma_lab = colr(dv.f[CCB_F_SIGMA_LAB], dv.cap, 0, i);
  in.mass_s = colr(dv.f[CCB_F_SPH_MASS], dv.cap, 0, i);
  in.bulk = colr(dv.f[CCB_F_BULK], dv.cap, 0, i);
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) in.shv[a][b] = colr(dv.f[CCB_F_SHV], dv.cap, a * 4 + b, i);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int mu = 0; mu < 4; ++mu) in.q[c][mu] = sp.fl.diffusion ? colr(dv.f[CCB_F_DIFFUSION], dv.cap, c * 4 + mu, i) : 0.0;
    in.div_qa[c] = sp.fl.diffusion ? colr(dv.f[CCB_F_DIV_QA], dv.cap, c, i) : 0.0;
    in.grad_qa[c] = sp.fl.diffusion ? colr(dv.f[CCB_F_GRAD_QA], dv.cap, c, i) : 0.0;
  }
}

// The equations of motion read ~130 columns per particle at 8 resident warps per SM (254 registers): too few loads in
// flight to cover DRAM latency. Every 16th thread therefore asks L2 for the 128-byte line of each input column of its 16
// neighbours before the arithmetic starts, so that the whole input of the CTA is in flight at once.
__device__ __forceinline__ void prefetch_l2(const double* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template <int D>
__device__ __forceinline__ void prefetch_eom_inputs(const DeviceView& dv, const StageParams& sp, int64_t i) {
  if ((threadIdx.x & 15) != 0) return;
#define CCB_PF(F, N) _Pragma("unroll") for (int c = 0; c < (N); ++c) prefetch_l2(dv.f[F] + (int64_t)c * dv.cap + i)
#define CCB_PF0(F, N) _Pragma("unroll") for (int c = 0; c < (N); ++c) prefetch_l2(dv.y0[F] + (int64_t)c * dv.cap + i)
  CCB_PF(CCB_F_THERMO, 16); CCB_PF(CCB_F_DALPHA, 12); CCB_PF(CCB_F_CHIBB, 2); CCB_PF(CCB_F_TRANSPORT, 13);
  CCB_PF(CCB_F_U, D); CCB_PF(CCB_F_V, D); CCB_PF(CCB_F_GRADP, D); CCB_PF(CCB_F_GRADBULK, D); CCB_PF(CCB_F_GRADSHEAR, D);
  CCB_PF(CCB_F_DIVSHEAR, D); CCB_PF(CCB_F_GRADV, 3 * D);
  CCB_PF(CCB_F_GAMMA, 1); CCB_PF(CCB_F_SIGMA, 1); CCB_PF(CCB_F_SIGMA_LAB, 1); CCB_PF(CCB_F_SPH_MASS, 1); CCB_PF(CCB_F_BULK, 1);
  CCB_PF(CCB_F_SHV, 16);
  if (sp.fl.diffusion) { CCB_PF(CCB_F_GRAD_ALPHA, 3 * D); CCB_PF(CCB_F_DIFFUSION, 12); CCB_PF(CCB_F_DIV_QA, 3); CCB_PF(CCB_F_GRAD_QA, 3); }
  if (sp.do_update) {
    CCB_PF0(CCB_F_EXTENSIVE, 4); CCB_PF0(CCB_F_EXTENSIVE_BULK, 1); CCB_PF0(CCB_F_EZ, 1); CCB_PF0(CCB_F_POSITION, D); CCB_PF0(CCB_F_U, D);
    CCB_PF0(CCB_F_EXTENSIVE_SHV, 5);
    if (sp.fl.diffusion) { CCB_PF0(CCB_F_EXTENSIVE_DIFFUSION, 9); }
  }
#undef CCB_PF
#undef CCB_PF0
}

template <int D, bool MILNE>
__global__ void __launch_bounds__(128) eom_rk_kernel(DeviceView dv, StageParams sp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dv.n) return;
  prefetch_eom_inputs<D>(dv, sp, i);
  Metric<MILNE> m(sp.t);
  EomOut<D> out;
  // Pass 1: the linear solve for du/dt and dN/dt. Pass 2 re-reads its inputs (L1/L2 hits) behind a compiler
  // barrier, so that nothing of pass 1 but the two results stays live: see physics_eom.cuh.
  {
    Thermo th;
    load_thermo(dv, i, th);
    Transport tc;
    load_transport(dv, i, tc);
    double kappa[3][3], tauq[3][3];
    diffusion_coefficients(th, sp.tp, sp.eos_type == CCB_EOS_TABLE && dv.eos_id[i] == 0, kappa, tauq);
    EomIn<D> in;
    load_eom_inputs<D>(dv, sp, i, in);
    eom_pass<D, MILNE, 1>(m, in, th, tc, kappa, tauq, sp.fl, out);
  }
  asm volatile("" ::: "memory");
  Thermo th;
  load_thermo(dv, i, th);
  Transport tc;
  load_transport(dv, i, tc);
  double kappa[3][3], tauq[3][3];
  diffusion_coefficients(th, sp.tp, sp.eos_type == CCB_EOS_TABLE && dv.eos_id[i] == 0, kappa, tauq);
  EomIn<D> in;
  load_eom_inputs<D>(dv, sp, i, in);
  eom_pass<D, MILNE, 2>(m, in, th, tc, kappa, tauq, sp.fl, out);

  // derivatives are kept: Output, RK4 accumulation and the parity tests read them
#pragma unroll
  for (int d = 0; d < 3; ++d) col(dv.f[CCB_F_DU_DT], dv.cap, d, i) = d < D ? out.du_dt[d < D ? d : 0] : 0.0;
  col(dv.f[CCB_F_D_EXTENSIVE_DT], dv.cap, 0, i) = out.ds_dt;
#pragma unroll
  for (int c = 0; c < 3; ++c) col(dv.f[CCB_F_D_EXTENSIVE_DT], dv.cap, 1 + c, i) = out.dN_dt[c];
  dv.f[CCB_F_D_BULK_DT][i] = out.dbulk_dt;
#pragma unroll
  for (int c = 0; c < 5; ++c) col(dv.f[CCB_F_D_SHV_DT], dv.cap, c, i) = out.dshv_dt[c];
  if (sp.fl.diffusion) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int d = 0; d < 3; ++d) col(dv.f[CCB_F_D_DIFFUSION_DT], dv.cap, c * 3 + d, i) = out.dq_dt[c][d];
  }
  dv.f[CCB_F_THETA][i] = out.theta;
  dv.f[CCB_F_SHV_NABLA_U][i] = out.shv_nabla_u;
  dv.f[CCB_F_DEZ][i] = out.dEz;

  if (!sp.do_update) return;
  // Evolver::step_rk, src/evolver.cpp:363-475: y = y0 + h * f
  const double h = sp.h_stage;
  double ext_s = dv.y0[CCB_F_EXTENSIVE][i] + h * out.ds_dt;
  if (ext_s < 0.0) {
    if (dv.freeze[i] > 3) ext_s = 1.e-3;
    else raise_error(dv, CCB_ERR_NEG_ENTROPY, dv.id[i]);
  }
  col(dv.f[CCB_F_EXTENSIVE], dv.cap, 0, i) = ext_s;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    col(dv.f[CCB_F_EXTENSIVE], dv.cap, 1 + c, i) = col(dv.y0[CCB_F_EXTENSIVE], dv.cap, 1 + c, i) + h * out.dN_dt[c];
  dv.f[CCB_F_EXTENSIVE_BULK][i] = dv.y0[CCB_F_EXTENSIVE_BULK][i] + h * out.dbulk_dt;
  dv.f[CCB_F_EZ][i] = dv.y0[CCB_F_EZ][i] + h * out.dEz;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    if (d < D) {
      col(dv.f[CCB_F_POSITION], dv.cap, d, i) = col(dv.y0[CCB_F_POSITION], dv.cap, d, i) + h * in.v[d < D ? d : 0];
      col(dv.f[CCB_F_U], dv.cap, d, i) = col(dv.y0[CCB_F_U], dv.cap, d, i) + h * out.du_dt[d < D ? d : 0];
    } else {
      col(dv.f[CCB_F_POSITION], dv.cap, d, i) = 0.0;
      col(dv.f[CCB_F_U], dv.cap, d, i) = 0.0;
    }
  }
#pragma unroll
  for (int c = 0; c < 5; ++c)
    col(dv.f[CCB_F_EXTENSIVE_SHV], dv.cap, c, i) = col(dv.y0[CCB_F_EXTENSIVE_SHV], dv.cap, c, i) + h * out.dshv_dt[c];
  if (sp.fl.diffusion) {
#pragma unroll
    for (int c = 0; c < 9; ++c)
      col(dv.f[CCB_F_EXTENSIVE_DIFFUSION], dv.cap, c, i) =
          col(dv.y0[CCB_F_EXTENSIVE_DIFFUSION], dv.cap, c, i) + h * out.dq_dt[c / 3][c % 3];
  }
}

// ---- initialisation kernels (BSQHydro::initialize_hydrodynamics on the device) ---
// tmp columns (current order): 0 e_or_s, 1 rhoB, 2 rhoS, 3 rhoQ, 4 bulk, 5..9 shv (xx,xy,xeta,yy,yeta), 10..18 q^i_a [charge][i]
// input_initial_shear: bit 0 = Settings::input_initial_shear, bit 1 = Settings::input_initial_diffusion
template <int D, bool MILNE>
__global__ void init_particles_kernel(DeviceView dv, StageParams sp, const double* __restrict__ tmp, int input_as_entropy,
                                      int input_initial_shear, int bulk_from_trace, double efcheck, int eos_done) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dv.n) return;
  const int64_t cap = dv.cap;
  const double hc = kHbarcGeVfm;
  const double rhoB = tmp[1 * cap + i], rhoS = tmp[2 * cap + i], rhoQ = tmp[3 * cap + i];
  double e_in = 0., s_in = 0.;
  Thermo th;
  // SPHWorkstation::initialize_entropy_and_charge_densities, src/sph_workstation.cpp:253-375
  if (input_as_entropy) s_in = tmp[i]; else e_in = tmp[i] / hc;
  if (eos_done) {  // tabulated EoS: eos_chain_init_kernel has already located the particle in the phase diagram
    load_thermo(dv, i, th);
  } else {
    double T, muB, muS, muQ;
    const bool ok = input_as_entropy ? conformal_invert_s(sp.conf, s_in, rhoB, rhoS, rhoQ, T, muB, muS, muQ)
                                     : conformal_invert_e(sp.conf, e_in, rhoB, rhoS, rhoQ, T, muB, muS, muQ);
    if (!ok) { raise_error(dv, CCB_ERR_EOS_NO_ROOT, dv.id[i]); return; }
    double r[TH_N];
    conformal_full_thermo(sp.conf, T, muB, muQ, muS, r);
    thermo_from_forward(T, muB, muQ, muS, r, 0, th);
    store_thermo(dv, i, th);
  }
  if (input_as_entropy) e_in = th.e; else s_in = th.s;
  dv.freeze[i] = (e_in > efcheck && dv.freeze[i] != 4) ? 0 : 4;
  const double sigma = dv.f[CCB_F_SIGMA][i];
  col(dv.f[CCB_F_EXTENSIVE], cap, 0, i) = s_in / sigma;
  col(dv.f[CCB_F_EXTENSIVE], cap, 1, i) = rhoB / sigma;
  col(dv.f[CCB_F_EXTENSIVE], cap, 2, i) = rhoS / sigma;
  col(dv.f[CCB_F_EXTENSIVE], cap, 3, i) = rhoQ / sigma;
  if (s_in < 0.0) dv.freeze[i] = 4;
  // calculate_extensive_shv / set_bulk_Pi / set_diffusion, src/sph_workstation.cpp:1350-1460
#pragma unroll
  for (int c = 0; c < 5; ++c) col(dv.f[CCB_F_EXTENSIVE_SHV], cap, c, i) = (input_initial_shear & 1) ? (tmp[(5 + c) * cap + i] / hc) / sigma : 0.0;
  if (sp.fl.diffusion) {  // SPHWorkstation::set_diffusion, src/sph_workstation.cpp:1441-1461
#pragma unroll
    for (int c = 0; c < 9; ++c) col(dv.f[CCB_F_EXTENSIVE_DIFFUSION], cap, c, i) = (input_initial_shear & 2) ? tmp[(10 + c) * cap + i] / sigma : 0.0;
  }
  dv.f[CCB_F_EXTENSIVE_BULK][i] = bulk_from_trace ? (tmp[4 * cap + i] / hc - th.p) / sigma : 0.0;
}

// S, B, S, Q, E, Ez, n_frozen in one pass (block reduce, one partial per block).
// Mirrors the reference sums term by term, including its use of sph_mass.s in the
// S and Q diffusion terms (src/system_state.cpp:770, 803, 812).
template <bool MILNE>
__global__ void reduce_totals_kernel(DeviceView dv, double t, double* __restrict__ out, unsigned long long* __restrict__ nfrozen) {
  double acc[6] = {0., 0., 0., 0., 0., 0.};
  unsigned int frozen = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < dv.n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ms = col(dv.f[CCB_F_SPH_MASS], dv.cap, 0, i), mB = col(dv.f[CCB_F_SPH_MASS], dv.cap, 1, i),
                 mS = col(dv.f[CCB_F_SPH_MASS], dv.cap, 2, i), mQ = col(dv.f[CCB_F_SPH_MASS], dv.cap, 3, i);
    const double gamma = dv.f[CCB_F_GAMMA][i], sigma = dv.f[CCB_F_SIGMA][i], sig_lab = dv.f[CCB_F_SIGMA_LAB][i];
    acc[0] += col(dv.f[CCB_F_EXTENSIVE], dv.cap, 0, i) * ms;
    const double qb0 = col(dv.f[CCB_F_DIFFUSION], dv.cap, 0, i), qs0 = col(dv.f[CCB_F_DIFFUSION], dv.cap, 4, i),
                 qq0 = col(dv.f[CCB_F_DIFFUSION], dv.cap, 8, i);
    acc[1] += col(dv.f[CCB_F_EXTENSIVE], dv.cap, 1, i) * mB + (qb0 * mB / sigma / gamma) * mB / gamma;
    acc[2] += col(dv.f[CCB_F_EXTENSIVE], dv.cap, 2, i) * mS + (qs0 * ms / sigma / gamma);
    acc[3] += col(dv.f[CCB_F_EXTENSIVE], dv.cap, 3, i) * mQ + (qq0 * ms / sigma / gamma);
    const double w = col(dv.f[CCB_F_THERMO], dv.cap, 11, i), p = col(dv.f[CCB_F_THERMO], dv.cap, 9, i);
    const double bulk = dv.f[CCB_F_BULK][i], shv00 = dv.f[CCB_F_SHV][i];
    const double C = w + bulk, g2 = gamma * gamma;
    acc[4] += (C * g2 - p - bulk + shv00) * ms * (MILNE ? t : 1.0) / sig_lab;
    acc[5] += dv.f[CCB_F_EZ][i];
    frozen += (dv.freeze[i] >= 4) ? 1u : 0u;
  }
  __shared__ double sh[6][8];
  __shared__ unsigned int shf[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sh[k][warp] = v;
  }
  for (int o = 16; o > 0; o >>= 1) frozen += __shfl_down_sync(0xffffffffu, frozen, o);
  if (lane == 0) shf[warp] = frozen;
  __syncthreads();
  // one partial per block, summed by the host in block order: deterministic totals
  if (threadIdx.x < 6) {
    double v = 0.;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += sh[threadIdx.x][w];
    out[blockIdx.x * 6 + threadIdx.x] = v;
  }
  if (threadIdx.x == 6) {
    unsigned int v = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += shf[w];
    nfrozen[blockIdx.x] = (unsigned long long)v;
  }
}

}  // namespace ccb
