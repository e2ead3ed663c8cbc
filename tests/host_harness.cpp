// TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// Compiles the per-particle physics headers of the library
// (ccake_b200/csrc/physics_*.cuh, stage_params.h) with the HOST compiler and exposes
// them over arrays, so that the arithmetic the CUDA kernels run can be compared with
// the oracle here, without a GPU. The shipped library never calls these on the host.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../ccake_b200/csrc/eos_preinverted.cuh"
#include "../ccake_b200/csrc/stage_params.h"

using namespace ccb;

extern "C" {

// thermo[16] + dalpha[12] from LRF densities (conformal EoS), per particle
int hh_eos(const ccb_config* cfg, int64_t n, const double* sBSQ, double* thermo, double* dalpha) {
  const StageParams sp = make_stage_params(*cfg, cfg->t0, 0., 0);
  for (int64_t i = 0; i < n; ++i) {
    double T, muB, muS, muQ;
    if (!conformal_invert_s(sp.conf, sBSQ[4 * i], sBSQ[4 * i + 1], sBSQ[4 * i + 2], sBSQ[4 * i + 3], T, muB, muS, muQ)) return 1;
    double r[TH_N];
    conformal_full_thermo(sp.conf, T, muB, muQ, muS, r);
    Thermo th;
    thermo_from_forward(T, muB, muQ, muS, r, 0, th);
    const double v[16] = {th.T, th.muB, th.muS, th.muQ, th.e, th.s, th.rhoB, th.rhoS, th.rhoQ, th.p, th.cs2, th.w, th.dwds, th.dwdB, th.dwdS, th.dwdQ};
    for (int k = 0; k < 16; ++k) thermo[16 * i + k] = v[k];
    for (int a = 0; a < 3; ++a) for (int x = 0; x < 4; ++x) dalpha[12 * i + a * 4 + x] = th.dalpha[a][x];
  }
  return 0;
}

}  // extern "C"
template <int D, bool MILNE>
static void local_impl(const StageParams& sp, double t, int64_t n, const double* u, const double* gamma, const double* sigma,
                       const double* ext_shv, const double* ext_q, const double* thermo, double* shv, double* q, double* transport) {
  Metric<MILNE> m(t);
  for (int64_t i = 0; i < n; ++i) {
    double uu[D];
    for (int d = 0; d < D; ++d) uu[d] = u[3 * i + d];
    double s4[4][4] = {}, qq[3][4];
    rebuild_shear<D, MILNE>(m, uu, gamma[i], sigma[i], ext_shv + 5 * i, s4);
    rebuild_diffusion<D, MILNE>(m, uu, gamma[i], sigma[i], ext_q + 9 * i, qq);
    for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) shv[16 * i + 4 * a + b] = s4[a][b];
    for (int c = 0; c < 3; ++c) for (int mu = 0; mu < 4; ++mu) q[12 * i + 4 * c + mu] = qq[c][mu];
    Thermo th{};
    const double* T = thermo + 16 * i;
    th.T = T[0]; th.muB = T[1]; th.muS = T[2]; th.muQ = T[3]; th.e = T[4]; th.s = T[5]; th.rhoB = T[6]; th.rhoS = T[7];
    th.rhoQ = T[8]; th.p = T[9]; th.cs2 = T[10]; th.w = T[11];
    Transport tc;
    transport_coefficients(th, sp.tp, tc);
    const double v[13] = {tc.eta_pi, tc.tau_pi, tc.zeta_Pi, tc.tau_Pi, tc.delta_PiPi, tc.delta_pipi, tc.lambda_piPi, tc.lambda_Pipi,
                          tc.phi1, tc.phi3, tc.phi6, tc.phi7, tc.tau_pipi};
    for (int k = 0; k < 13; ++k) transport[13 * i + k] = v[k];
  }
}

extern "C" {
int hh_local(const ccb_config* cfg, double t, int64_t n, const double* u, const double* gamma, const double* sigma,
             const double* ext_shv, const double* ext_q, const double* thermo, double* shv, double* q, double* transport) {
  const StageParams sp = make_stage_params(*cfg, t, 0., 0);
  const bool milne = cfg->coordinate_system == CCB_COORD_HYPERBOLIC;
  if (cfg->dim == 2) { if (milne) local_impl<2, true>(sp, t, n, u, gamma, sigma, ext_shv, ext_q, thermo, shv, q, transport); else local_impl<2, false>(sp, t, n, u, gamma, sigma, ext_shv, ext_q, thermo, shv, q, transport); }
  else { if (milne) local_impl<3, true>(sp, t, n, u, gamma, sigma, ext_shv, ext_q, thermo, shv, q, transport); else local_impl<3, false>(sp, t, n, u, gamma, sigma, ext_shv, ext_q, thermo, shv, q, transport); }
  return 0;
}

struct hh_eom_arrays {
  const double *u, *v, *gamma, *sigma, *sigma_lab, *mass_s, *bulk, *shv, *q, *thermo, *dalpha;
  const double *gradP, *gradBulk, *gradshear, *divshear, *gradV, *div_qa, *grad_qa, *grad_alpha;
  double *du_dt, *d_ext, *d_bulk, *d_shv, *d_q, *theta, *shv_nabla_u, *dEz;
};

}  // extern "C"
template <int D, bool MILNE> static void eom_impl(const StageParams& sp, double t, int64_t n, const hh_eom_arrays& A) {
  Metric<MILNE> m(t);
  for (int64_t i = 0; i < n; ++i) {
    Thermo th{};
    const double* T = A.thermo + 16 * i;
    th.T = T[0]; th.muB = T[1]; th.muS = T[2]; th.muQ = T[3]; th.e = T[4]; th.s = T[5]; th.rhoB = T[6]; th.rhoS = T[7];
    th.rhoQ = T[8]; th.p = T[9]; th.cs2 = T[10]; th.w = T[11]; th.dwds = T[12]; th.dwdB = T[13]; th.dwdS = T[14]; th.dwdQ = T[15];
    for (int a = 0; a < 3; ++a) for (int x = 0; x < 4; ++x) th.dalpha[a][x] = A.dalpha[12 * i + 4 * a + x];
    Transport tc;
    transport_coefficients(th, sp.tp, tc);
    double kappa[3][3], tauq[3][3];
    diffusion_coefficients(th, sp.tp, false, kappa, tauq);
    EomIn<D> in;
    for (int d = 0; d < D; ++d) {
      in.u[d] = A.u[3 * i + d]; in.v[d] = A.v[3 * i + d]; in.gradP[d] = A.gradP[3 * i + d]; in.gradBulk[d] = A.gradBulk[3 * i + d];
      in.gradshear[d] = A.gradshear[3 * i + d]; in.divshear[d] = A.divshear[3 * i + d];
      for (int e = 0; e < D; ++e) in.gradV[d][e] = A.gradV[9 * i + 3 * d + e];
      for (int c = 0; c < 3; ++c) in.grad_alpha[d][c] = A.grad_alpha[9 * i + 3 * d + c];
    }
    in.gamma = A.gamma[i]; in.sigma = A.sigma[i]; in.sigma_lab = A.sigma_lab[i]; in.mass_s = A.mass_s[i]; in.bulk = A.bulk[i];
    for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) in.shv[a][b] = A.shv[16 * i + 4 * a + b];
    for (int c = 0; c < 3; ++c) {
      for (int mu = 0; mu < 4; ++mu) in.q[c][mu] = A.q[12 * i + 4 * c + mu];
      in.div_qa[c] = A.div_qa[3 * i + c]; in.grad_qa[c] = A.grad_qa[3 * i + c];
    }
    EomOut<D> out;
    eom_evaluate<D, MILNE>(m, in, th, tc, kappa, tauq, sp.fl, out);
    for (int d = 0; d < 3; ++d) A.du_dt[3 * i + d] = d < D ? out.du_dt[d < D ? d : 0] : 0.;
    A.d_ext[4 * i] = out.ds_dt;
    for (int c = 0; c < 3; ++c) A.d_ext[4 * i + 1 + c] = out.dN_dt[c];
    A.d_bulk[i] = out.dbulk_dt;
    for (int c = 0; c < 5; ++c) A.d_shv[5 * i + c] = out.dshv_dt[c];
    for (int c = 0; c < 3; ++c) for (int d = 0; d < 3; ++d) A.d_q[9 * i + 3 * c + d] = out.dq_dt[c][d];
    A.theta[i] = out.theta; A.shv_nabla_u[i] = out.shv_nabla_u; A.dEz[i] = out.dEz;
  }
}

extern "C" {
int hh_eom(const ccb_config* cfg, double t, int64_t n, const hh_eom_arrays* A) {
  const StageParams sp = make_stage_params(*cfg, t, 0., 0);
  const bool milne = cfg->coordinate_system == CCB_COORD_HYPERBOLIC;
  if (cfg->dim == 2) { if (milne) eom_impl<2, true>(sp, t, n, *A); else eom_impl<2, false>(sp, t, n, *A); }
  else { if (milne) eom_impl<3, true>(sp, t, n, *A); else eom_impl<3, false>(sp, t, n, *A); }
  return 0;
}

// kernel values for the golden vectors of unit_tests/kernel_test.cpp
double hh_kernel(int D, double r, double h) { const KernelNorm k = make_kernel_norm(D, h); return sph_w(k, r); }
double hh_grad_kernel_factor(int D, double r, double h) { const KernelNorm k = make_kernel_norm(D, h); return sph_gradw_factor(k, r); }

}  // extern "C"

// ---- tabulated EoS + fall-back chain (physics_eos_chain.cuh, eos_table_host.h) on the host -------------------
#include "../ccake_b200/csrc/eos_table_host.h"

namespace {
struct HostChain {
  HostTable table;
  std::vector<double> hot_s, hot_e;
  EosChain chain;
};
}  // namespace

extern "C" {

void* hh_chain_create(const char* thermo_dat, int normalize_by_T, double freeze_out_T_MeV, int prohibit_unstable_cs2,
                      int prohibit_acausal_cs2, int restrict_mu_T_ratios) {
  auto* h = new HostChain();
  std::string err;
  std::memset(&h->chain, 0, sizeof(h->chain));
  if (!read_thermo_dat(thermo_dat, h->table, err) || !prepare_table(h->table, normalize_by_T != 0, err) ||
      !match_fallbacks(h->table, freeze_out_T_MeV, h->chain, err)) {
    std::fprintf(stderr, "hh_chain_create: %s\n", err.c_str());
    delete h;
    return nullptr;
  }
  const long long nodes = h->table.nodes();
  h->hot_s.resize((size_t)nodes * 4); h->hot_e.resize((size_t)nodes * 4);
  for (long long i = 0; i < nodes; ++i) {
    const double* f = &h->table.fields[(size_t)i * TH_N];
    h->hot_s[4 * i] = f[TH_S]; h->hot_e[4 * i] = f[TH_E];
    for (int c = 1; c < 4; ++c) { const double v = c == 1 ? f[TH_B] : c == 2 ? f[TH_SS] : f[TH_Q]; h->hot_s[4 * i + c] = v; h->hot_e[4 * i + c] = v; }
  }
  h->chain.table = host_view(h->table);
  h->chain.table.hot_s = h->hot_s.data(); h->chain.table.hot_e = h->hot_e.data();
  h->chain.prohibit_unstable_cs2 = prohibit_unstable_cs2; h->chain.prohibit_acausal_cs2 = prohibit_acausal_cs2;
  h->chain.restrict_mu_T_ratios = restrict_mu_T_ratios;
  return h;
}
void hh_chain_destroy(void* h) { delete static_cast<HostChain*>(h); }

// interpolated 17 fields at points x = (T, muB, muQ, muS)
int hh_chain_table_thermo(void* hv, int64_t n, const double* x, double* r17) {
  auto* h = static_cast<HostChain*>(hv);
  for (int64_t i = 0; i < n; ++i) table_full_thermo(h->chain.table, x + 4 * i, r17 + TH_N * i);
  return 0;
}
// forward thermodynamics of chain member k at points x
int hh_chain_member_thermo(void* hv, int k, int64_t n, const double* x, double* r17) {
  auto* h = static_cast<HostChain*>(hv);
  for (int64_t i = 0; i < n; ++i) chain_full_thermo(h->chain, k, x + 4 * i, r17 + TH_N * i);
  return 0;
}
void hh_chain_parameters(void* hv, double* out17) {
  auto* h = static_cast<HostChain*>(hv);
  const AnalyticParams* P[3] = {&h->chain.tanh_p, &h->chain.conf_p, &h->chain.diag_p};
  int o = 0;
  for (int k = 0; k < 3; ++k) {
    out17[o++] = P[k]->c; out17[o++] = P[k]->T0; out17[o++] = P[k]->muB0; out17[o++] = P[k]->muS0; out17[o++] = P[k]->muQ0;
    if (k == 0) { out17[o++] = P[k]->Tc; out17[o++] = P[k]->Ts; }
  }
}

// phase 04 per particle: sBSQ[4n] targets, seed[4n] = previous (T, muB, muS, muQ), eos[n] in/out.
// energy != 0: the initial e -> s location (targets[0] = e in fm^-4), seeds unused.
int hh_chain_solve(void* hv, int energy, int64_t n, const double* targets, const double* seed, int32_t* eos, double* thermo, double* dalpha,
                   double* chi) {
  auto* h = static_cast<HostChain*>(hv);
  int failures = 0;
  for (int64_t i = 0; i < n; ++i) {
    ChainResult res;
    bool ok;
    if (energy) ok = chain_s_out(h->chain, targets[4 * i], targets[4 * i + 1], targets[4 * i + 2], targets[4 * i + 3], res);
    else {
      const double pos[4] = {seed[4 * i], seed[4 * i + 1], seed[4 * i + 3], seed[4 * i + 2]};
      ok = chain_update_s(h->chain, eos[i], pos, targets[4 * i], targets[4 * i + 1], targets[4 * i + 2], targets[4 * i + 3], res);
    }
    if (!ok) { ++failures; eos[i] = -1; continue; }
    Thermo th;
    chain_thermo(h->chain, res, th);
    const double v[16] = {th.T, th.muB, th.muS, th.muQ, th.e, th.s, th.rhoB, th.rhoS, th.rhoQ, th.p, th.cs2, th.w, th.dwds, th.dwdB, th.dwdS, th.dwdQ};
    for (int k = 0; k < 16; ++k) thermo[16 * i + k] = v[k];
    for (int a = 0; a < 3; ++a) for (int x = 0; x < 4; ++x) dalpha[12 * i + a * 4 + x] = th.dalpha[a][x];
    chi[2 * i] = th.chiBB; chi[2 * i + 1] = th.chiBB0;
    eos[i] = res.eos;
  }
  return failures;
}

}  // extern "C"

// ---- pre-inverted EoS (eos_preinverted.cuh) on the host: the arithmetic of pre_fill_kernel / pre_init_kernel -----------------
extern "C" {
// tiles: ntiles x {slot[4], N[4], log[4]} ints, {min[4], step[4]} doubles, vars pointers; dens[n][4] -> out[n][11], ok[n]
int hh_pre_fill(int ntiles, const int* tile_ints, const double* tile_doubles, const double* const* tile_vars, int flat, int64_t n, const double* dens,
                double* out, int* ok) {
  std::vector<PreTile> tiles(kPreTileCount);
  for (auto& t : tiles) std::memset(&t, 0, sizeof(t));
  for (int k = 0; k < ntiles; ++k) {
    const int* ti = tile_ints + 12 * k;
    PreTile& t = tiles[((ti[0] * 4 + ti[1]) * 4 + ti[2]) * 4 + ti[3]];
    for (int i = 0; i < 4; ++i) { t.N[i] = ti[4 + i]; t.log_spacing[i] = ti[8 + i]; t.min[i] = tile_doubles[8 * k + i]; t.step[i] = tile_doubles[8 * k + 4 + i]; }
    t.vars = tile_vars[k];
  }
  const PreTable tb{tiles.data(), flat};
  for (int64_t i = 0; i < n; ++i) ok[i] = pre_fill(tb, dens[4 * i], dens[4 * i + 1], dens[4 * i + 2], dens[4 * i + 3], out + PRE_NVAR * i) ? 1 : 0;
  return 0;
}
// flat-table initialisation queries: arrays[11] = s_axis nB_axis T muB muS muQ e p cs2 dwds dwdB (fm units);
// mode 0 query_thermo(a = s), 1 invert_e_to_s(a = e) -> out[n][12] = T muB muS muQ e p cs2 s rhoB w dwds dwdB
int hh_pre_flat_query(int Ns, int NB, const double* const* arrays, int mode, int64_t n, const double* a, const double* rhoB, double* out, int* ok,
                      double T_fo, double* efreeze) {
  PreFlat f;
  f.Ns = Ns; f.NB = NB;
  f.s_axis = arrays[0]; f.nB_axis = arrays[1]; f.T = arrays[2]; f.muB = arrays[3]; f.muS = arrays[4]; f.muQ = arrays[5];
  f.e = arrays[6]; f.p = arrays[7]; f.cs2 = arrays[8]; f.dwds = arrays[9]; f.dwdB = arrays[10];
  f.log_s_min = std::log10(f.s_axis[0]); f.dlog_s = (std::log10(f.s_axis[Ns - 1]) - f.log_s_min) / (Ns - 1);
  f.log_nB_min = std::log10(f.nB_axis[0]); f.dlog_nB = (std::log10(f.nB_axis[NB - 1]) - f.log_nB_min) / (NB - 1);
  for (int64_t i = 0; i < n; ++i) {
    PreThermo q{};
    ok[i] = (mode ? pre_invert_e_to_s(f, a[i], rhoB[i], q) : pre_query_thermo(f, a[i], rhoB[i], q)) ? 1 : 0;
    const double v[12] = {q.T, q.muB, q.muS, q.muQ, q.e, q.p, q.cs2, q.s, q.rhoB, q.w, q.dwds, q.dwdB};
    for (int k = 0; k < 12; ++k) out[12 * i + k] = v[k];
  }
  if (efreeze) *efreeze = pre_efreeze_at_T(f, T_fo);
  return 0;
}
}  // extern "C"
