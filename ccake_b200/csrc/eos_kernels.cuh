// Phase 04 for a tabulated equation of state: one thread locates one particle in the phase diagram
// (physics_eos_chain.cuh) and writes its thermodynamics. The reference does this as a serial host loop between
// two deep copies of the whole particle array (src/sph_workstation.cpp:1714-1772).
// Bound: fp64 pipe + L2 latency (16 scattered 32-byte vertex reads per Newton iterate from the 4-field "hot"
// table, which is sized to stay L2-resident); the 17-field table is touched once per particle, after convergence.
#pragma once

#include "eos_preinverted.cuh"
#include "local.cuh"
#include "physics_eos_chain.cuh"

// The kernels below inline the whole fall-back chain and take minutes each to compile; inst_eos.cu is therefore built as four
// objects (-DCCB_EOS_PART=1..4: stage + fast path | initial location | host queries | pre-inverted tables and their generator) that
// compile in parallel. Without the macro every kernel is visible (the host test harness, editors).
#ifndef CCB_EOS_PART
#define CCB_EOS_PART 0
#endif
#define CCB_EOS_HAS(part) (CCB_EOS_PART == 0 || CCB_EOS_PART == (part))

namespace ccb {

#if CCB_EOS_HAS(1)
// (s, rhoB, rhoS, rhoQ) of the local rest frame (thermo columns 5..8, written by the field gather) -> thermo.
// SPHWorkstation::locate_phase_diagram_point_sBSQ, src/sph_workstation.cpp:659-741
// `todo` (nullable): per-particle flags of the fast path below; a particle it has already located is skipped.
__global__ void __launch_bounds__(128) eos_chain_stage_kernel(DeviceView dv, EosChain ch, const uint32_t* __restrict__ todo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dv.n) return;
  if (todo && !todo[i]) return;
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

// Fast path of phase 04 on a table run, in two lean kernels. A particle that was on the table after the previous stage (eos_id = 0)
// almost always stays there, and the full chain would locate it with its FIRST search -- Newton from the previous point on the table.
//   eos_table_root_kernel     does exactly that search (find_root<FIRST_ONLY, TABLE>) and parks the root in a scratch column block;
//   eos_table_epilogue_kernel evaluates the 17 table fields there, applies the member's acceptance test and stores the thermodynamics.
// Whatever they cannot finish (no convergence, acceptance test failed, particle not on the table) is flagged in `todo` and done by
// eos_chain_stage_kernel, which repeats the same first search with the same outcome: results are identical bit for bit. Why two
// kernels: ncu put the one-kernel form (14 k instructions = 228 KB of code, profiles/r02af) at 5.7 warps stalled on instruction
// fetch per issue; the search alone is 3.4 k instructions, the epilogue 4.9 k, each inside the instruction cache, and neither
// carries the analytic members' code (TABLE). todo: 1 = full chain, 2 = root parked, 0 = done.
template <int MINB>
__global__ void __launch_bounds__(128, MINB) eos_table_root_kernel(DeviceView dv, EosChain ch, uint32_t* __restrict__ todo, double* __restrict__ xs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dv.n) return;
  if (dv.eos_id[i] != 0 || ch.kind[0] != EOS_KIND_TABLE) { todo[i] = 1u; return; }
  const double s = col(dv.f[CCB_F_THERMO], dv.cap, 5, i), B = col(dv.f[CCB_F_THERMO], dv.cap, 6, i),
               S = col(dv.f[CCB_F_THERMO], dv.cap, 7, i), Q = col(dv.f[CCB_F_THERMO], dv.cap, 8, i);
  double x[4] = {col(dv.f[CCB_F_TSEED], dv.cap, 0, i), col(dv.f[CCB_F_TSEED], dv.cap, 1, i),
                 col(dv.f[CCB_F_TSEED], dv.cap, 3, i), col(dv.f[CCB_F_TSEED], dv.cap, 2, i)};
  if (!find_root<true, true>(ch, 0, false, s, B, S, Q, x)) { todo[i] = 1u; return; }
  todo[i] = 2u;
#pragma unroll
  for (int d = 0; d < 4; ++d) xs[(int64_t)d * dv.cap + i] = x[d];   // the seed columns stay untouched: the full chain may still need them
}
template <int MINB>
__global__ void __launch_bounds__(128, MINB) eos_table_epilogue_kernel(DeviceView dv, EosChain ch, uint32_t* __restrict__ todo, const double* __restrict__ xs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dv.n || todo[i] != 2u) return;
  const double x[4] = {xs[i], xs[dv.cap + i], xs[2 * dv.cap + i], xs[3 * dv.cap + i]};
  ChainResult res;
  chain_evaluate<true>(ch, 0, x, res);
  if (!chain_accepts(ch, 0, res)) { todo[i] = 1u; return; }
  todo[i] = 0u;
  Thermo th;
  chain_thermo(ch, res, th);
  store_thermo(dv, i, th);   // eos_id stays 0
}
#endif  // part 1

#if CCB_EOS_HAS(2)
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

#endif  // part 2

#if CCB_EOS_HAS(3)
// ---- host-callable queries (C ABI: ccb_eos_full_thermo, ccb_eos_locate) -----------------------------------------------
// EoS_base::get_full_thermo (include/eos_base.h:12-50) of chain member `member` at n points (T, muB, muQ, muS); a point
// outside the member's domain gives a row of NaN. conf != nullptr: the closed-form conformal EoS of a conformal run.
__global__ void eos_query_kernel(EosChain ch, const ConformalParams* conf_or_null, ConformalParams conf, int member, int64_t n,
                                 const double* __restrict__ pts, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x[4] = {pts[i * 4 + 0], pts[i * 4 + 1], pts[i * 4 + 2], pts[i * 4 + 3]};
  double r[TH_N];
  bool ok = true;
  if (conf_or_null) conformal_full_thermo(conf, x[0], x[1], x[2], x[3], r);
  else ok = chain_full_thermo(ch, member, x, r);
  for (int k = 0; k < TH_N; ++k) out[i * TH_N + k] = ok ? r[k] : nan("");
}
// EquationOfState::update_s / s_out (src/eos.cpp:574-804, 834-977) + set_thermo for n density vectors (s | e, rhoB, rhoS, rhoQ):
// located from `seed` (T, muB, muS, muQ; ignored when energy != 0) through the chain, starting at member start[i].
// thermo: 16 columns per row in CCB_F_THERMO order; eos_out: the member that answered, -1 = no solution.
__global__ void eos_locate_kernel(EosChain ch, ConformalParams conf, int use_chain, int energy, int64_t n, const double* __restrict__ dens,
                                  const double* __restrict__ seed, const int32_t* __restrict__ start, double* __restrict__ thermo,
                                  int32_t* __restrict__ eos_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a = dens[i * 4 + 0], B = dens[i * 4 + 1], S = dens[i * 4 + 2], Q = dens[i * 4 + 3];
  Thermo th;
  int eos = -1;
  bool ok;
  if (use_chain) {
    ChainResult res;
    if (energy) ok = chain_s_out(ch, a, B, S, Q, res);
    else {
      const double pos[4] = {seed[i * 4 + 0], seed[i * 4 + 1], seed[i * 4 + 3], seed[i * 4 + 2]};  // solver order (T, muB, muQ, muS)
      ok = chain_update_s(ch, start ? start[i] : 0, pos, a, B, S, Q, res);
    }
    if (ok) { chain_thermo(ch, res, th); eos = res.eos; }
  } else {
    double T, muB, muS, muQ;
    // closed-form conformal EoS: (s | e, rho) -> (T, mu) (EoS_conformal::get_tbqs_seed_from_{s,e}BSQ are exact inverses there)
    ok = energy ? conformal_invert_e(conf, a, B, S, Q, T, muB, muS, muQ) : conformal_invert_s(conf, a, B, S, Q, T, muB, muS, muQ);
    if (ok) {
      double r[TH_N];
      conformal_full_thermo(conf, T, muB, muQ, muS, r);
      thermo_from_forward(T, muB, muQ, muS, r, 0, th);
      eos = 0;
    }
  }
  const double v[16] = {th.T, th.muB, th.muS, th.muQ, th.e, th.s, th.rhoB, th.rhoS, th.rhoQ, th.p, th.cs2, th.w, th.dwds, th.dwdB, th.dwdS, th.dwdQ};
  for (int k = 0; k < 16; ++k) thermo[i * 16 + k] = ok ? v[k] : nan("");
  eos_out[i] = eos;
}

#endif  // part 3

#if CCB_EOS_HAS(4)
// ---- pre-inverted equation of state (eos_preinverted.cuh) ---------------------------------------------------------------
// Phase 04 with Settings::online_inverter_enabled = false: EoS_Interpolator::fill_thermodynamics (src/eos_interpolator.cpp:531-627).
// Writes T, muB, muS, muQ, e, p, cs2, w = e + p and dw/d(s, B, S, Q); s and the charge densities stay what the field gather stored.
__global__ void __launch_bounds__(128) pre_fill_kernel(DeviceView dv, PreTable tb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dv.n) return;
  double* T = dv.f[CCB_F_THERMO];
  double o[PRE_NVAR];
  if (!pre_fill(tb, col(T, dv.cap, 5, i), col(T, dv.cap, 6, i), col(T, dv.cap, 7, i), col(T, dv.cap, 8, i), o)) {
    raise_error(dv, CCB_ERR_EOS_NO_ROOT, dv.id[i]);
    return;
  }
  col(T, dv.cap, 0, i) = o[PRE_T]; col(T, dv.cap, 1, i) = o[PRE_MUB]; col(T, dv.cap, 2, i) = o[PRE_MUS]; col(T, dv.cap, 3, i) = o[PRE_MUQ];
  col(T, dv.cap, 4, i) = o[PRE_E]; col(T, dv.cap, 9, i) = o[PRE_P]; col(T, dv.cap, 10, i) = o[PRE_CS2]; col(T, dv.cap, 11, i) = o[PRE_E] + o[PRE_P];
  col(T, dv.cap, 12, i) = o[PRE_DWDS]; col(T, dv.cap, 13, i) = o[PRE_DWDB]; col(T, dv.cap, 14, i) = o[PRE_DWDSS]; col(T, dv.cap, 15, i) = o[PRE_DWDQ];
  col(dv.f[CCB_F_TSEED], dv.cap, 0, i) = o[PRE_T]; col(dv.f[CCB_F_TSEED], dv.cap, 1, i) = o[PRE_MUB];
  col(dv.f[CCB_F_TSEED], dv.cap, 2, i) = o[PRE_MUS]; col(dv.f[CCB_F_TSEED], dv.cap, 3, i) = o[PRE_MUQ];
}
// Initial location on the flat (s, nB) table (SPHWorkstation::initialize_entropy_and_charge_densities with use_preinverted,
// src/sph_workstation.cpp:260-320): query_thermo_host for entropy input, invert_e_to_s_host for energy input.
__global__ void __launch_bounds__(128) pre_init_kernel(DeviceView dv, PreFlat fl, const double* __restrict__ tmp, int input_as_entropy) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dv.n) return;
  const int64_t cap = dv.cap;
  const double rhoB = tmp[1 * cap + i];
  PreThermo q;
  const bool ok = input_as_entropy ? pre_query_thermo(fl, tmp[i], rhoB, q) : pre_invert_e_to_s(fl, tmp[i] / kHbarcGeVfm, rhoB, q);
  if (!ok) { raise_error(dv, CCB_ERR_EOS_NO_ROOT, dv.id[i]); return; }
  Thermo th;
  th.T = q.T; th.muB = q.muB; th.muS = q.muS; th.muQ = q.muQ; th.e = q.e; th.s = q.s; th.rhoB = q.rhoB; th.rhoS = 0.0; th.rhoQ = 0.0;
  th.p = q.p; th.cs2 = q.cs2; th.w = q.w; th.dwds = q.dwds; th.dwdB = q.dwdB; th.dwdS = 0.0; th.dwdQ = 0.0;
  for (int a = 0; a < 3; ++a) for (int x = 0; x < 4; ++x) th.dalpha[a][x] = 0.0;
  th.chiBB = 0.0; th.chiBB0 = 0.0;
  store_thermo(dv, i, th);
  dv.eos_id[i] = 0;
}
// One tile of the pre-inverted table (utilities/generate_eos_table.cpp:223-330): node (is, iB, iS, iQ) of the decade tile
// `exponent` is x = 10^e + i * 10^(e + 1) / N per axis; every node is located from the generator's seed (T = 750 / 197 fm^-1,
// mu = 0) with the run's own inversion (update_s through the chain, or the closed-form conformal inversion).
// vars: [PRE_NVAR][N^4]; eos: answering member per node, -1 (and NaN) where the reference aborts ("root-finding failed").
__global__ void __launch_bounds__(128) pre_generate_kernel(EosChain ch, ConformalParams conf, int use_chain, int e0, int e1, int e2, int e3, int N,
                                                           double* __restrict__ vars, int32_t* __restrict__ eos_out) {
  const int64_t total = (int64_t)N * N * N * N;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= total) return;
  const int iQ = (int)(k % N), iS = (int)((k / N) % N), iB = (int)((k / ((int64_t)N * N)) % N), is = (int)(k / ((int64_t)N * N * N));
  const int ex[4] = {e0, e1, e2, e3}, ii[4] = {is, iB, iS, iQ};
  double x[4];
  for (int j = 0; j < 4; ++j) { const double xmin = pow(10., (double)ex[j]), xmax = pow(10., (double)(ex[j] + 1)); x[j] = xmin + ii[j] * (xmax / N); }
  Thermo th;
  int eos = -1;
  bool ok;
  if (use_chain) {
    const double pos[4] = {750. / 197., 0.0, 0.0, 0.0};
    ChainResult res;
    ok = chain_update_s(ch, 0, pos, x[0], x[1], x[2], x[3], res);
    if (ok) { chain_thermo(ch, res, th); eos = res.eos; }
  } else {
    double T, muB, muS, muQ;
    ok = conformal_invert_s(conf, x[0], x[1], x[2], x[3], T, muB, muS, muQ);
    if (ok) {
      double r[TH_N];
      conformal_full_thermo(conf, T, muB, muQ, muS, r);
      thermo_from_forward(T, muB, muQ, muS, r, 0, th);
      eos = 0;
    }
  }
  const double v[PRE_NVAR] = {th.T, th.muB, th.muQ, th.muS, th.e, th.p, th.cs2, th.dwds, th.dwdB, th.dwdQ, th.dwdS};
  for (int q = 0; q < PRE_NVAR; ++q) vars[(int64_t)q * total + k] = ok ? v[q] : nan("");
  eos_out[k] = eos;
}

#endif  // part 4

}  // namespace ccb
