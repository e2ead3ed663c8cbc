// Phases 03 and 09: the two SPH neighbour sums, as atomic-free warp-per-cell gathers.
//
// Reference: SPHWorkstation::smooth_all_particle_fields (src/sph_workstation.cpp:495-578)
// and smooth_all_particle_gradients (:797-986). There, one Kokkos team per particle
// walks a CSR neighbour list and issues 5 (fields) or 5D+D^2 [+6+3D] (gradients)
// atomic adds PER PAIR into a 4.5 KB/particle AoSoA. Here:
//   * particles are already sorted by cell, so the neighbours of a cell are the
//     contiguous particle ranges of its 3^D surrounding cells;
//   * a warp owns up to 32 target particles of ONE cell; every lane streams the same
//     source records (warp-uniform 128-bit loads of a packed record), so the loads are
//     broadcast and the cost is the fp64 pipe, not memory;
//   * sums accumulate in registers and are written once: no atomics, and results are
//     reproducible run to run (the reference's atomic ordering is not).
// The source records are packed and pre-scaled by the per-particle kernel that runs
// just before (p/sigma^2, bulk/sigma^2, pi/sigma^2 ...), which removes the divisions
// the reference re-evaluates for every pair.
#pragma once

#include "gather_params.cuh"

namespace ccb {

__device__ __forceinline__ void load_rec(const double* __restrict__ p, int n, double* r) {
  const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
  for (int k = 0; k < n / 2; ++k) { const double2 v = __ldg(q + k); r[2 * k] = v.x; r[2 * k + 1] = v.y; }
}

// Exactly the predicate of the neighbour list: dist_sqr = dx*dx + dy*dy + dz*dz, no FMA contraction.
__device__ __forceinline__ double dist_sqr_exact(double dx, double dy, double dz) {
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

template <int D, bool FACT>
__global__ void __launch_bounds__(128) gather_fields_kernel(DeviceView dv, CellGrid g, GatherParams gp) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int64_t nwarps = (int64_t)gridDim.x * warps_per_block;
  const int ntasks = *dv.ntasks;
  for (int64_t task = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); task < ntasks; task += nwarps) {
    const int first = dv.tasks[task];
    const uint32_t key = dv.key[first];
    const uint32_t cend = dv.cell_start[key + 1];
    const int64_t i = (int64_t)first + lane;
    const bool active = i < cend;
    const int64_t ii = active ? i : first;
    double ra[REC_F];
    load_rec(dv.rec_fields + ii * REC_F, REC_F, ra);
    const int cz = key % g.nc[2], cy = (key / g.nc[2]) % g.nc[1], cx = key / (g.nc[2] * g.nc[1]);
    double a_sig = 0., a_s = 0., a_B = 0., a_S = 0., a_Q = 0.;
    int cnt = 0;
    for (int ox = (cx > 0 ? -1 : 0); ox <= (cx < g.nc[0] - 1 ? 1 : 0); ++ox)
      for (int oy = (cy > 0 ? -1 : 0); oy <= (cy < g.nc[1] - 1 ? 1 : 0); ++oy)
        for (int oz = (cz > 0 ? -1 : 0); oz <= (cz < g.nc[2] - 1 ? 1 : 0); ++oz) {
          const uint32_t nb = (uint32_t)(((int64_t)(cx + ox) * g.nc[1] + (cy + oy)) * g.nc[2] + (cz + oz));
          const uint32_t js = dv.cell_start[nb], je = dv.cell_start[nb + 1];
          for (uint32_t j = js; j < je; ++j) {
            double rb[REC_F];
            load_rec(dv.rec_fields + (int64_t)j * REC_F, REC_F, rb);
            const double dx = ra[0] - rb[0], dy = ra[1] - rb[1], dz = ra[2] - rb[2];
            const double r2 = dist_sqr_exact(dx, dy, dz);
            if (r2 <= gp.rsqr) {
              ++cnt;
              double w;
              if (FACT) {
                const double rp = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                w = sph_w(gp.k2, rp) * sph_w(gp.k1, fabs(dz));
              } else {
                w = sph_w(gp.kD, sqrt(r2));
              }
              a_sig += rb[3] * w;
              a_s += rb[4] * w;
              a_B += rb[5] * w;
              a_S += rb[6] * w;
              a_Q += rb[7] * w;
            }
          }
        }
    if (active) {
      const double gamma = dv.f[CCB_F_GAMMA][i];
      double sigma, s, rB, rS, rQ;
      if (gp.milne) {
        Metric<true> m(gp.t);
        sigma = to_lrf(m, a_sig, gamma); s = to_lrf(m, a_s, gamma); rB = to_lrf(m, a_B, gamma);
        rS = to_lrf(m, a_S, gamma); rQ = to_lrf(m, a_Q, gamma);
      } else {
        Metric<false> m(gp.t);
        sigma = to_lrf(m, a_sig, gamma); s = to_lrf(m, a_s, gamma); rB = to_lrf(m, a_B, gamma);
        rS = to_lrf(m, a_S, gamma); rQ = to_lrf(m, a_Q, gamma);
      }
      dv.f[CCB_F_SIGMA_LAB][i] = a_sig;
      dv.f[CCB_F_SIGMA][i] = sigma;
      col(dv.f[CCB_F_SMOOTHED], dv.cap, 0, i) = a_s;
      col(dv.f[CCB_F_SMOOTHED], dv.cap, 1, i) = a_B;
      col(dv.f[CCB_F_SMOOTHED], dv.cap, 2, i) = a_S;
      col(dv.f[CCB_F_SMOOTHED], dv.cap, 3, i) = a_Q;
      col(dv.f[CCB_F_THERMO], dv.cap, 5, i) = s;
      col(dv.f[CCB_F_THERMO], dv.cap, 6, i) = rB;
      col(dv.f[CCB_F_THERMO], dv.cap, 7, i) = rS;
      col(dv.f[CCB_F_THERMO], dv.cap, 8, i) = rQ;
      // btrack = number of neighbours (self excluded), sticky at -1 (src/system_state.cpp:712-715)
      if (dv.btrack[i] != -1) dv.btrack[i] = cnt - 1;
    }
  }
}

// grad_a W(r_a - r_b): plain kernel (src/kernel.cpp:97-118) or factorised D=3 variant (:153-170)
template <int D, bool FACT>
__device__ __forceinline__ bool grad_kernel(const GatherParams& gp, const double* rel, double r2, double* gK) {
  if (FACT) {
    const double rp = sqrt(__dadd_rn(__dmul_rn(rel[0], rel[0]), __dmul_rn(rel[1], rel[1])));
    const double de = fabs(rel[2]);
    const double f2 = sph_gradw_factor(gp.k2, rp), f1 = sph_gradw_factor(gp.k1, de);
    const double W2 = sph_w(gp.k2, rp), W1 = sph_w(gp.k1, de);
    gK[0] = (f2 * rel[0]) * W1;
    gK[1] = (f2 * rel[1]) * W1;
    gK[2] = W2 * (f1 * rel[2]);
    return true;
  }
  if (r2 >= 4.0 * gp.kD.h * gp.kD.h * 1.0000001) return false;  // cheap reject; exact branch below
  const double r = sqrt(r2);
  const double f = sph_gradw_factor(gp.kD, r);
#pragma unroll
  for (int d = 0; d < D; ++d) gK[d] = f * rel[d];
  return f != 0.0;
}

template <int D, bool SHEAR, bool DIFF, bool FACT>
__global__ void __launch_bounds__(128) gather_gradients_kernel(DeviceView dv, CellGrid g, GatherParams gp) {
  constexpr int REC = DIFF ? REC_G1 : REC_G0;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int64_t nwarps = (int64_t)gridDim.x * warps_per_block;
  const int ntasks = *dv.ntasks;
  for (int64_t task = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); task < ntasks; task += nwarps) {
    const int first = dv.tasks[task];
    const uint32_t key = dv.key[first];
    const uint32_t cend = dv.cell_start[key + 1];
    const int64_t i = (int64_t)first + lane;
    const bool active = i < cend;
    const int64_t ii = active ? i : first;
    double ra[REC];
    load_rec(dv.rec_grad + ii * REC, REC, ra);
    const double sig_a = ra[RG_SIGLAB];
    const int cz = key % g.nc[2], cy = (key / g.nc[2]) % g.nc[1], cx = key / (g.nc[2] * g.nc[1]);
    double gP[D], gE[D], gB[D], gSh[D], dSh[D], gV[D][D], dq[3], gq[3], gal[D][3];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      gP[d] = gE[d] = gB[d] = gSh[d] = dSh[d] = 0.;
#pragma unroll
      for (int e = 0; e < D; ++e) gV[d][e] = 0.;
#pragma unroll
      for (int c = 0; c < 3; ++c) gal[d][c] = 0.;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) dq[c] = gq[c] = 0.;

    for (int ox = (cx > 0 ? -1 : 0); ox <= (cx < g.nc[0] - 1 ? 1 : 0); ++ox)
      for (int oy = (cy > 0 ? -1 : 0); oy <= (cy < g.nc[1] - 1 ? 1 : 0); ++oy)
        for (int oz = (cz > 0 ? -1 : 0); oz <= (cz < g.nc[2] - 1 ? 1 : 0); ++oz) {
          const uint32_t nb = (uint32_t)(((int64_t)(cx + ox) * g.nc[1] + (cy + oy)) * g.nc[2] + (cz + oz));
          const uint32_t js = dv.cell_start[nb], je = dv.cell_start[nb + 1];
          for (uint32_t j = js; j < je; ++j) {
            const double* pb = dv.rec_grad + (int64_t)j * REC;
            const double2 b01 = __ldg(reinterpret_cast<const double2*>(pb));
            const double2 b23 = __ldg(reinterpret_cast<const double2*>(pb) + 1);
            double rel[3] = {ra[0] - b01.x, ra[1] - b01.y, ra[2] - b23.x};
            const double r2 = dist_sqr_exact(rel[0], rel[1], rel[2]);
            if (!(r2 <= gp.rsqr) || (int64_t)j == i) continue;  // outside the neighbour list, or self (grad W(0) = 0)
            double gK[D];
            if (!grad_kernel<D, FACT>(gp, rel, r2, gK)) continue;
            double rb[REC];
            load_rec(pb, REC, rb);
            const double m_b = rb[RG_M];
            double sK[D];
#pragma unroll
            for (int d = 0; d < D; ++d) sK[d] = (m_b * sig_a) * gK[d];
            const double cP = rb[RG_PT] + ra[RG_PT], cB = rb[RG_BT] + ra[RG_BT];
            const double mos = m_b / sig_a;
            const double cE = mos * (rb[RG_E] - ra[RG_E]);
#pragma unroll
            for (int d = 0; d < D; ++d) {
              gP[d] += cP * sK[d];
              gE[d] += cE * gK[d];
              gB[d] += cB * sK[d];
#pragma unroll
              for (int e = 0; e < D; ++e) gV[d][e] += mos * (rb[RG_V + e] - ra[RG_V + e]) * gK[d];
            }
            double Kv = 0.;
#pragma unroll
            for (int d = 0; d < D; ++d) Kv += sK[d] * ra[RG_V + d];
            if (SHEAR) {
              // symmetric spatial block packed as (11,12,13,22,23,33)
              const int sm[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
#pragma unroll
              for (int d = 0; d < D; ++d) {
                gSh[d] += Kv * (rb[RG_S0 + d] + ra[RG_S0 + d]);
                double x = 0.;
#pragma unroll
                for (int e = 0; e < D; ++e) x += sK[e] * rb[RG_SM + sm[d][e]] + sK[e] * ra[RG_SM + sm[d][e]];
                dSh[d] += x;
              }
            }
            if (DIFF) {
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                gq[c] += Kv * (rb[RG_Q0 + c] + ra[RG_Q0 + c]);
                double x = 0.;
#pragma unroll
                for (int d = 0; d < D; ++d) {
                  x += sK[d] * rb[RG_QI + c * 3 + d] + sK[d] * ra[RG_QI + c * 3 + d];
                  gal[d][c] += mos * (rb[RG_ALPHA + c] - ra[RG_ALPHA + c]) * gK[d];
                }
                dq[c] += x;
              }
            }
          }
        }
    if (active) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        col(dv.f[CCB_F_GRADP], dv.cap, d, i) = d < D ? gP[d < D ? d : 0] : 0.;
        col(dv.f[CCB_F_GRADE], dv.cap, d, i) = d < D ? gE[d < D ? d : 0] : 0.;
        col(dv.f[CCB_F_GRADBULK], dv.cap, d, i) = d < D ? gB[d < D ? d : 0] : 0.;
        col(dv.f[CCB_F_GRADSHEAR], dv.cap, d, i) = d < D ? gSh[d < D ? d : 0] : 0.;
        col(dv.f[CCB_F_DIVSHEAR], dv.cap, d, i) = d < D ? dSh[d < D ? d : 0] : 0.;
#pragma unroll
        for (int e = 0; e < 3; ++e)
          col(dv.f[CCB_F_GRADV], dv.cap, d * 3 + e, i) = (d < D && e < D) ? gV[d < D ? d : 0][e < D ? e : 0] : 0.;
#pragma unroll
        for (int c = 0; c < 3; ++c) col(dv.f[CCB_F_GRAD_ALPHA], dv.cap, d * 3 + c, i) = d < D ? gal[d < D ? d : 0][c] : 0.;
        col(dv.f[CCB_F_DIV_QA], dv.cap, d, i) = dq[d];
        col(dv.f[CCB_F_GRAD_QA], dv.cap, d, i) = gq[d];
      }
    }
  }
}

}  // namespace ccb
