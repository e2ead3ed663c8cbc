// Pre-inverted equation of state: phase 04 when Settings::online_inverter_enabled = false, the reference's DEFAULT mode
// (include/defaults.h:15). Reference: EoS_Interpolator (include/eos_interpolator.h:45-163, src/eos_interpolator.cpp):
//   fill_thermodynamics  :531-627   (s, rhoB, rhoS, rhoQ) of the local rest frame -> T, muB, muS, muQ, e, p, cs2, dw/d(s,B,S,Q), w
//   interpolate4D        :639-737   multilinear over 16 vertices, with the 1-D (s only) and 2-D (s, rhoB) fall-backs
//   load_flat_table      :262-521   the flat (s, nB) format: unit conversion + dw/ds, dw/dB from the chi tensor
//   query_thermo_host / invert_e_to_s_host / efreeze_at_T :745-930   initialisation queries of the flat table
//
// Tables arrive as arrays through the C ABI (ccb_set_preinverted_table / ccb_set_preinverted_flat): the reference reads
// them from HDF5 (absent here, and I/O anyway). Tiled format: up to 5 x 4 x 4 x 4 tiles, one per decade of (s, |rhoB|, |rhoS|,
// |rhoQ|) from 1e-3 upwards, 20 points per axis (utilities/generate_eos_table.cpp:189-203); flat format: ONE tile in slot
// [0][0][0][0] with log-spaced s and nB axes and degenerate S, Q axes.
//
// One thread per particle: 11 fields x 16 vertices of one tile, all of them in the same few cache lines per field; the
// particles of a warp are neighbours in space, hence in the table. Bound: L2 latency / bandwidth of a ~0.1-2 GB table that is
// read at 11 x 16 x 8 B = 1.4 KB per particle.
#pragma once

#include "physics_common.cuh"

namespace ccb {

// order of ccake::eos_variables (include/eos_interpolator.h:36-43)
enum { PRE_T = 0, PRE_MUB, PRE_MUQ, PRE_MUS, PRE_E, PRE_P, PRE_CS2, PRE_DWDS, PRE_DWDB, PRE_DWDQ, PRE_DWDSS, PRE_NVAR };
constexpr int kPreTileCount = 5 * 4 * 4 * 4;  // NTBL_s x NTBL_B x NTBL_S x NTBL_Q (include/eos_interpolator.h:45-48)

// table_attributes of the four axes + the 11 fields of one tile; vars: [ivar][is][iB][iS][iQ] (Q fastest), nullptr = tile absent
struct PreTile {
  double min[4], step[4];
  int N[4];
  int log_spacing[4];
  const double* vars;
};
struct PreTable {
  const PreTile* tiles;  // kPreTileCount entries, slot ((ts * 4 + tB) * 4 + tS) * 4 + tQ
  int flat;
};

// One field of one tile at the cell idx, position pos (interpolate4D, src/eos_interpolator.cpp:639-737). A vertex index beyond
// the last node of an axis (the reference clamps idx to N - 1 and then reads idx + 1: out of bounds there, :593-607) is held
// at the last node.
CCB_HD double pre_interpolate4D(const PreTile& t, const int idx[4], const double pos[4], int ivar) {
  const double* f = t.vars + (long long)ivar * t.N[0] * t.N[1] * t.N[2] * t.N[3];
  auto at = [&](int i0, int i1, int i2, int i3) {
    i0 = i0 < t.N[0] ? i0 : t.N[0] - 1; i1 = i1 < t.N[1] ? i1 : t.N[1] - 1;
    i2 = i2 < t.N[2] ? i2 : t.N[2] - 1; i3 = i3 < t.N[3] ? i3 : t.N[3] - 1;
    return f[(((long long)i0 * t.N[1] + i1) * t.N[2] + i2) * t.N[3] + i3];
  };
  const double ds = t.step[0], dB = t.step[1], dS = t.step[2], dQ = t.step[3];
  if (dB == 0.0 && dS == 0.0 && dQ == 0.0) {  // 1-D fall-back along s
    const double x0 = (pos[0] - t.min[0] - idx[0] * ds) / ds;
    return (1.0 - x0) * at(idx[0], idx[1], idx[2], idx[3]) + x0 * at(idx[0] + 1, idx[1], idx[2], idx[3]);
  }
  if (dS == 0.0 && dQ == 0.0) {  // 2-D bilinear in (s, rhoB)
    const double x0 = (pos[0] - t.min[0] - idx[0] * ds) / ds;
    const double x1 = (pos[1] - t.min[1] - idx[1] * dB) / dB;
    double f2 = 0.0;
    CCB_UNROLL for (int k = 0; k < 4; ++k) {
      double phi = 1.0;
      phi *= (k & 1) ? x0 : (1.0 - x0);
      phi *= ((k >> 1) & 1) ? x1 : (1.0 - x1);
      f2 += phi * at(idx[0] + (k & 1), idx[1] + ((k >> 1) & 1), idx[2], idx[3]);
    }
    return f2;
  }
  double x[4];
  CCB_UNROLL for (int i = 0; i < 4; ++i) x[i] = (fabs(t.step[i]) < 1e-30) ? 0.0 : (pos[i] - t.min[i] - idx[i] * t.step[i]) / t.step[i];
  double fsum = 0.0;
  CCB_UNROLL for (int k = 0; k < 16; ++k) {
    int b[4];
    double phi = 1.0;
    CCB_UNROLL for (int i = 0; i < 4; ++i) {
      b[i] = (t.N[i] <= 1) ? 0 : ((k >> i) & 1);   // never step to the + 1 corner of a single-node axis
      phi *= b[i] ? x[i] : (1.0 - x[i]);
    }
    fsum += phi * at(idx[0] + b[0], idx[1] + b[1], idx[2] + b[2], idx[3] + b[3]);
  }
  return fsum;
}

// EoS_Interpolator::fill_thermodynamics for one particle: out[PRE_NVAR]; returns false when the tile is absent
CCB_HD bool pre_fill(const PreTable& tb, double s_in, double rhoB_in, double rhoS_in, double rhoQ_in, double out[PRE_NVAR]) {
  const double s = fmax(.001, s_in), rhoB = fmax(.001, fabs(rhoB_in)), rhoS = fmax(.001, fabs(rhoS_in)), rhoQ = fmax(.001, fabs(rhoQ_in));
  int tl[4] = {0, 0, 0, 0};
  if (!tb.flat) {
    const double v[4] = {s, rhoB, rhoS, rhoQ};
    CCB_UNROLL for (int i = 0; i < 4; ++i) {
      int ie = (int)floor(log10(fabs(v[i])) + 3);
      ie = ie > 0 ? ie : 0;
      const int last = (i == 0) ? 4 : 3;  // NTBL_s - 1, NTBL_{B,S,Q} - 1
      tl[i] = ie < last ? ie : last;
    }
  }
  const PreTile& t = tb.tiles[((tl[0] * 4 + tl[1]) * 4 + tl[2]) * 4 + tl[3]];
  if (!t.vars) return false;
  double pos[4] = {s, rhoB, rhoS, rhoQ};
  int idx[4];
  CCB_UNROLL for (int i = 0; i < 4; ++i) {
    if (t.log_spacing[i]) pos[i] = log10(fmax(pos[i], 1e-30));
    if (fabs(t.step[i]) < 1e-30) idx[i] = 0;
    else {
      int k = (int)floor((pos[i] - t.min[i]) / t.step[i]);
      k = k > 0 ? k : 0;
      idx[i] = k < t.N[i] - 1 ? k : t.N[i] - 1;
    }
  }
  CCB_UNROLL for (int v = 0; v < PRE_NVAR; ++v) out[v] = pre_interpolate4D(t, idx, pos, v);
  return true;
}

// ---- the flat (s, nB) table: initialisation queries (src/eos_interpolator.cpp:745-930) ---------------------------------
struct PreFlat {
  int Ns, NB;
  double log_s_min, dlog_s, log_nB_min, dlog_nB;
  const double *s_axis, *nB_axis;                                    // physical, fm^-3
  const double *T, *muB, *muS, *muQ, *e, *p, *cs2, *dwds, *dwdB;     // [is * NB + iB]
};
struct PreThermo { double T, muB, muS, muQ, e, p, cs2, s, rhoB, w, dwds, dwdB; };

// EoS_Interpolator::query_thermo_host
CCB_HD bool pre_query_thermo(const PreFlat& f, double s, double rhoB, PreThermo& out) {
  const double abs_rhoB = fmax(fabs(rhoB), f.nB_axis[0]);
  const double s_clamped = fmax(s, f.s_axis[0]);
  if (s_clamped <= 0.0) return false;
  const int Ns = f.Ns, NB = f.NB;
  const double L_s_max = f.log_s_min + f.dlog_s * (Ns - 1), L_nB_max = f.log_nB_min + f.dlog_nB * (NB - 1);
  const double L_s = fmin(fmax(log10(s_clamped), f.log_s_min), L_s_max);
  const double L_nB = fmin(fmax(log10(abs_rhoB), f.log_nB_min), L_nB_max);
  const double xi_s = (L_s - f.log_s_min) / f.dlog_s, xi_B = (L_nB - f.log_nB_min) / f.dlog_nB;
  int is = (int)floor(xi_s), iB = (int)floor(xi_B);
  is = is > 0 ? is : 0; is = is < Ns - 2 ? is : Ns - 2;
  iB = iB > 0 ? iB : 0; iB = iB < NB - 2 ? iB : NB - 2;
  const double a = xi_s - is, b = xi_B - iB;
  auto bil = [&](const double* q) {
    return (1.0 - a) * (1.0 - b) * q[(long long)is * NB + iB] + a * (1.0 - b) * q[(long long)(is + 1) * NB + iB] +
           (1.0 - a) * b * q[(long long)is * NB + iB + 1] + a * b * q[(long long)(is + 1) * NB + iB + 1];
  };
  out.T = bil(f.T); out.muB = bil(f.muB); out.muS = bil(f.muS); out.muQ = bil(f.muQ);
  out.e = bil(f.e); out.p = bil(f.p); out.cs2 = bil(f.cs2);
  out.s = s; out.rhoB = rhoB; out.w = out.e + out.p;
  out.dwds = bil(f.dwds); out.dwdB = bil(f.dwdB);
  return true;
}
// EoS_Interpolator::invert_e_to_s_host
CCB_HD bool pre_invert_e_to_s(const PreFlat& f, double e_input, double rhoB, PreThermo& out) {
  const int Ns = f.Ns, NB = f.NB;
  const double L_nB_max = f.log_nB_min + f.dlog_nB * (NB - 1);
  const double abs_rhoB = fmax(fabs(rhoB), f.nB_axis[0]);
  const double L_nB = fmin(log10(abs_rhoB), L_nB_max);
  const double xi_B = (L_nB - f.log_nB_min) / f.dlog_nB;
  int iB = (int)floor(xi_B);
  iB = iB > 0 ? iB : 0; iB = iB < NB - 2 ? iB : NB - 2;
  const double b = xi_B - iB;
  auto e_at = [&](int k) { return (1.0 - b) * f.e[(long long)k * NB + iB] + b * f.e[(long long)k * NB + iB + 1]; };
  if (e_input <= e_at(0)) return pre_query_thermo(f, f.s_axis[0], rhoB, out);
  if (e_input >= e_at(Ns - 1)) return pre_query_thermo(f, f.s_axis[Ns - 1], rhoB, out);
  int lo = 0, hi = Ns - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) / 2;
    if (e_at(mid) <= e_input) lo = mid; else hi = mid;
  }
  const double e_lo = e_at(lo), e_hi = e_at(hi);
  const double alpha = (e_hi > e_lo) ? (e_input - e_lo) / (e_hi - e_lo) : 0.0;
  const double L_lo = log10(f.s_axis[lo]), L_hi = log10(f.s_axis[hi]);
  return pre_query_thermo(f, pow(10.0, L_lo + alpha * (L_hi - L_lo)), rhoB, out);
}
// EoS_Interpolator::efreeze_at_T
CCB_HD double pre_efreeze_at_T(const PreFlat& f, double T_fm) {
  const int Ns = f.Ns, NB = f.NB;
  if (T_fm <= f.T[0]) return f.e[0];
  if (T_fm >= f.T[(long long)(Ns - 1) * NB]) return f.e[(long long)(Ns - 1) * NB];
  int lo = 0, hi = Ns - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) / 2;
    if (f.T[(long long)mid * NB] <= T_fm) lo = mid; else hi = mid;
  }
  const double T_l = f.T[(long long)lo * NB], T_h = f.T[(long long)hi * NB];
  const double alpha = (T_h > T_l) ? (T_fm - T_l) / (T_h - T_l) : 0.0;
  return (1.0 - alpha) * f.e[(long long)lo * NB] + alpha * f.e[(long long)hi * NB];
}

}  // namespace ccb
