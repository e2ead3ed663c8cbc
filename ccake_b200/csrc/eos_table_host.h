// Host side of the tabulated equation of state: reading `thermo.dat`, the unit conversions of the EoS_table
// constructor and the matching of the three analytic fall-back EoSs to the table. Host code of the library
// (the reference does the same work on the host at start-up); nothing here runs per timestep.
//
// Reference: InterpolatorND<4>::load_data_from_dat / construct_interpolant (src/interpolatorND.cpp:125-226),
// EoS_table::EoS_table (src/eos_table.cpp:14-122), EquationOfState::set_up_chosen_EOSs (src/eos_initialization.cpp:273-455).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "physics_eos_chain.cuh"

namespace ccb {

struct HostTable {
  int n[4] = {0, 0, 0, 0};
  std::vector<double> grid;    // nodes x 4, fm^-1 after prepare()
  std::vector<double> fields;  // nodes x 17, fm^-n after prepare()
  double min[4], max[4], delta[4];
  long long nodes() const { return (long long)n[0] * n[1] * n[2] * n[3]; }
};

// first line: the grid sizes; every other line: 4 coordinates then the fields (17 expected)
inline bool read_thermo_dat(const std::string& path, HostTable& t, std::string& err) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) { err = "File " + path + " could not be opened"; return false; }
  std::fseek(f, 0, SEEK_END);
  const long sz = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<char> buf((size_t)sz + 1);
  if (std::fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) { std::fclose(f); err = "short read of " + path; return false; }
  std::fclose(f);
  buf[(size_t)sz] = 0;
  char* p = buf.data();
  char* eol = std::strchr(p, '\n');
  if (!eol) { err = "empty table file"; return false; }
  *eol = 0;
  int dims = 0;
  for (char* q = p; dims < 5;) {
    char* e;
    const long v = std::strtol(q, &e, 10);
    if (e == q) break;
    if (dims < 4) t.n[dims] = (int)v;
    ++dims; q = e;
  }
  if (dims != 4) { err = "Your equation of state table has the wrong dimension!"; return false; }
  const long long nodes = t.nodes();
  t.grid.resize((size_t)nodes * 4);
  t.fields.resize((size_t)nodes * TH_N);
  p = eol + 1;
  for (long long row = 0; row < nodes; ++row) {
    for (int c = 0; c < 4 + TH_N; ++c) {
      char* e;
      const double v = std::strtod(p, &e);
      if (e == p) { err = "table file ends early or has fewer than 21 columns (row " + std::to_string(row) + ")"; return false; }
      p = e;
      if (c < 4) t.grid[(size_t)row * 4 + c] = v; else t.fields[(size_t)row * TH_N + (c - 4)] = v;
    }
  }
  return true;
}

// EoS_table constructor: axes to fm^-1, fields to fm^-n, then the interpolant's geometry.
inline bool prepare_table(HostTable& t, bool normalize_by_T, std::string& err) {
  const long long nodes = t.nodes();
  if (nodes <= 0 || (long long)t.grid.size() != nodes * 4 || (long long)t.fields.size() != nodes * TH_N) { err = "table arrays do not match the grid sizes"; return false; }
  const double hc1 = 1.0 / kHbarcMeVfm;
  for (auto& v : t.grid) v *= hc1;
  // field order: p s B S Q e cs2 chiBB chiQQ chiSS chiBQ chiBS chiQS chiTB chiTQ chiTS chiTT
  static const int power[TH_N] = {4, 3, 3, 3, 3, 4, 0, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2};
  if (normalize_by_T) {
    for (long long i = 0; i < nodes; ++i) {
      const double T = t.grid[(size_t)i * 4];
      for (int c = 0; c < TH_N; ++c)
        if (power[c]) t.fields[(size_t)i * TH_N + c] *= std::pow(T, power[c]);
    }
  } else {
    const double hc2 = hc1 * hc1, hc3 = hc2 * hc1, hc4 = hc3 * hc1;
    const double fac[5] = {1.0, hc1, hc2, hc3, hc4};
    for (long long i = 0; i < nodes; ++i)
      for (int c = 0; c < TH_N; ++c)
        if (power[c]) t.fields[(size_t)i * TH_N + c] *= fac[power[c]];
  }
  for (int d = 0; d < 4; ++d) {
    double lo = t.grid[d], hi = t.grid[d];
    for (long long i = 1; i < nodes; ++i) { const double v = t.grid[(size_t)i * 4 + d]; lo = std::min(lo, v); hi = std::max(hi, v); }
    t.min[d] = lo; t.max[d] = hi;
    t.delta[d] = (t.n[d] <= 1) ? 0.0 : (hi - lo) / (t.n[d] - 1.0);
  }
  return true;
}

inline TableView host_view(const HostTable& t) {
  TableView v;
  for (int d = 0; d < 4; ++d) { v.n[d] = t.n[d]; v.min[d] = t.min[d]; v.max[d] = t.max[d]; v.delta[d] = t.delta[d]; }
  v.hot_s = nullptr; v.hot_e = nullptr; v.full = t.fields.data();
  return v;
}

// EquationOfState::set_up_chosen_EOSs for eos_type "table" (src/eos_initialization.cpp:273-455): the chain
// table -> tanh_conformal -> conformal -> conformal_diagonal with every coefficient matched to the table's
// pressure at T = 1.1 T_FO and at the table's largest chemical potentials.
inline bool match_fallbacks(const HostTable& t, double freeze_out_T_MeV, EosChain& ch, std::string& err) {
  const TableView tv = host_view(t);
  const double Tmax = 1.1 * (freeze_out_T_MeV / kHbarcMeVfm);
  auto pressure = [&](double T, double muB, double muQ, double muS, double& p) {
    const double x[4] = {T, muB, muQ, muS};
    for (int d = 0; d < 4; ++d) if (x[d] < t.min[d] || x[d] > t.max[d]) return false;  // EquationOfState::tbqs exits here
    double r[TH_N];
    if (!table_full_thermo(tv, x, r)) return false;
    p = r[TH_P];
    return true;
  };
  double pT = 0.0;
  if (!pressure(Tmax, 0.0, 0.0, 0.0, pT)) { err = "EoS table does not cover (1.1 T_FO, 0, 0, 0)"; return false; }
  const double c = pT / (Tmax * Tmax * Tmax * Tmax);
  const double mumax[3] = {t.max[1], t.max[2], t.max[3]};  // B, Q, S
  double mu0[3], mu0d[3];
  for (int a = 0; a < 3; ++a) {
    if (std::fabs(mumax[a]) < 1e-30) { mu0[a] = 1.0; mu0d[a] = 1.0; continue; }
    double pv = 0.0;
    if (!pressure(Tmax, a == 0 ? mumax[0] : 0.0, a == 1 ? mumax[1] : 0.0, a == 2 ? mumax[2] : 0.0, pv)) { err = "EoS table does not cover (1.1 T_FO, mu_max)"; return false; }
    mu0[a] = std::pow(c, 0.25) * mumax[a] / std::sqrt(std::sqrt(pv) - std::sqrt(pT));
    mu0d[a] = std::pow(c / (pv - pT), 0.25) * mumax[a];
  }
  ch.n = 4;
  ch.kind[0] = EOS_KIND_TABLE; ch.kind[1] = EOS_KIND_TANH; ch.kind[2] = EOS_KIND_CONFORMAL; ch.kind[3] = EOS_KIND_DIAGONAL;
  const double Tc = 220.0 / kHbarcMeVfm, Ts = 120.0 / kHbarcMeVfm;
  ch.tanh_p = AnalyticParams{c, 1.0, mu0[0], mu0[2], mu0[1], Tc, Ts};
  ch.conf_p = AnalyticParams{c, 1.0, mu0[0], mu0[2], mu0[1], 0.0, 0.0};
  ch.diag_p = AnalyticParams{c, 1.0, mu0d[0], mu0d[2], mu0d[1], 0.0, 0.0};
  return true;
}

}  // namespace ccb
