// Phases 03 and 09: the two SPH neighbour sums as atomic-free, shared-memory-tiled gathers.
//
// Reference: SPHWorkstation::smooth_all_particle_fields (src/sph_workstation.cpp:495-578) and
// smooth_all_particle_gradients (:797-986). There, one Kokkos team per particle walks a CSR
// neighbour list and issues 5 (fields) or 5D+D^2 [+6+3D] (gradients) atomic adds PER PAIR into
// a 4.5 KB/particle AoSoA. Here (particles are already sorted by cell, phase 01):
//
//   * a CTA takes a task = up to NT consecutive particles of ONE cell of the cell-sorted order (task list: cells.cuh), one
//     target per thread with its sums in registers. Their neighbours are the particles of the 3^D stencil around that cell:
//     3 (2-D) or 9 (3-D) CONTIGUOUS ranges of the sorted records (a row of cells along the fastest grid axis is contiguous
//     in memory); in a slab run a ghost column is one more such range, in the neighbouring rank's record array (peer
//     memory) or behind the own records. CTAs are persistent and pull tasks from a device-side counter;
//   * the candidate sources are streamed through shared memory in tiles of 64 records, double-buffered with cp.async
//     (tile k + 1 lands while tile k is processed). Tile k takes every ntiles-th candidate, so each tile is a uniform
//     sample of the neighbourhood and every target finds about the same number of neighbours in it;
//   * phase 1 (per tile): every thread tests its target against all staged sources in fp32, in the expanded form
//     |b|^2 - 2 a.b <= R^2 - |a|^2 (coordinates relative to a nearby cell corner; the cut-off carries a bound on the
//     rounding, so no true neighbour is missed), and keeps the hits as mask words. The reads are warp-uniform
//     (shared-memory broadcast) and run on the fp32 pipe, off the fp64 pipe;
//   * phase 2: every thread walks ITS OWN hit bits, re-tests exactly in fp64 (the reference's predicate
//     dx*dx + dy*dy + dz*dz <= (2h)^2, unfused) and evaluates the pair term from the staged record, 16 bytes at a time.
//     A warp iterates max-over-lanes(hits) times instead of once per candidate: only ~1/3 (2-D) to ~1/6 (3-D) of the
//     candidates in the 3^D cells are neighbours, which is what a lane-per-target loop over all candidates wastes;
//   * the gradient sums are accumulated in moment form (GradPolicy), so the target's own record is read once, in the
//     epilogue; parts of the records that are zero in every particle are skipped (DeviceView::qflag);
//   * sums stay in registers and are written once: no atomics, reproducible run to run (the reference's atomic ordering
//     is not).
#pragma once

#include "gather_params.cuh"

namespace ccb {

constexpr int kGatherTile = 64;  // source records per shared-memory tile (default)

// Exactly the predicate of the neighbour list: dist_sqr = dx*dx + dy*dy + dz*dz, no FMA contraction.
__device__ __forceinline__ double dist_sqr_exact(double dx, double dy, double dz) {
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// 1 / sqrt(x) for a positive NORMAL x (a squared pair distance inside the cut-off): the hardware seed (MUFU.RSQ64H, 2^-22) and
// one third-order correction -- the sequence rsqrt() itself runs, without its range check and slow-path call (2 ulp).
__device__ __forceinline__ double rsqrt_pos(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-(y * y), x, 1.0);
  return fma(fma(e, 0.375, 0.5), y * e, y);
}
// r and 1/r from r^2 with one rsqrt instead of sqrt + divisions; r2 == 0 (self) is handled by the callers
__device__ __forceinline__ void r_and_inv(double r2, double& r, double& ir) {
  ir = rsqrt_pos(r2);
  r = r2 * ir;
}
// W(r, h) of src/kernel.cpp:36-48 with q = r * (1/h)
__device__ __forceinline__ double sph_w_fast(const KernelNorm& k, double r) {
  const double q = r * k.inv_h;
  const double tq = 2.0 - q;
  const double qq = q * q;
  const double inner = k.norm * (1 - 1.5 * qq + 0.75 * q * qq);
  const double outer = k.norm2 * tq * tq * tq;
  return q >= 2.0 ? 0.0 : (q >= 1.0 ? outer : inner);
}
// f with grad_a W = f (r_a - r_b), src/kernel.cpp:97-118, from r and 1/r
__device__ __forceinline__ double sph_gradw_factor_fast(const KernelNorm& k, double r, double ir) {
  const double q = r * k.inv_h;
  const double tq = 2.0 - q;
  const double outer = (k.norm_grad * ir) * tq * tq;
  const double inner = k.norm_grad2 * (1.5 * q - 2) * k.inv_h;
  return q >= 2.0 ? 0.0 : (q >= 1.0 ? outer : inner);
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
  double r, ir;
  r_and_inv(r2, r, ir);
  const double f = sph_gradw_factor_fast(gp.kD, r, ir);
#pragma unroll
  for (int d = 0; d < D; ++d) gK[d] = f * rel[d];
  return f != 0.0;
}

// ------------------------------------------------------------------------------------------
// smooth_fields: sigma_lab, smoothed s, rho_B, rho_S, rho_Q, neighbour count
template <int D_, bool FACT, bool EMIT_ = false>
struct FieldsPolicy {
  static constexpr int D = D_, REC = REC_F, STRIDE = REC_F_STRIDE, CH_STAGE = REC_F / 2, VARIANT = -1;
  static constexpr int MASK_MODE = 1;  // stores the exact hit masks of its tiles (GatherParams::hitmask)
  // EMIT: debug twin that also writes the ids of the neighbours it visits (ccb_debug_neighbour_list): GatherParams::nbr_off / nbr_idx
  static constexpr bool EMIT = EMIT_;
  struct Target { double x[3]; };
  struct Acc { double sig, s, B, S, Q; int cnt; };
  __device__ static const double* records(const DeviceView& dv) { return dv.rec_fields; }
  __device__ static void init(Acc& a) { a.sig = a.s = a.B = a.S = a.Q = 0.; a.cnt = 0; }
  __device__ static void load_target(const double* r, Target& t) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(r)), b = __ldg(reinterpret_cast<const double2*>(r) + 1);
    t.x[0] = a.x; t.x[1] = a.y; t.x[2] = b.x;
  }
  // A pair term in two halves:
  //   prep : exact fp64 predicate + kernel value; returns whether the pair is in the neighbour list
  //   accum: the sums
  struct Prep { double w; int hit; };
  __device__ static bool prep(const GatherParams& gp, const Target& t, const double2* sb, bool valid, Prep& p) {
    const double2 b0 = sb[0], b1 = sb[1];
    const double dx = t.x[0] - b0.x, dy = t.x[1] - b0.y, dz = (D == 3) ? t.x[2] - b1.x : 0.0;
    // D = 2: z = 0 for every particle, and adding the exact zero dz*dz does not change the sum: skipped
    const double r2 = (D == 3) ? dist_sqr_exact(dx, dy, dz) : __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    // the self term W(0) is part of the sum (reset_fields, src/sph_workstation.cpp:514-523). The candidates of a task are exactly
    // the 3^D cells around the targets' cell, the cells the reference's list searches: a pair on the cut-off needs no special care
    const bool in = valid && r2 <= gp.rsqr;
    if (!in) { p.w = 0.0; p.hit = 0; return false; }
    double w;
    if (FACT) {
      const double rp = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
      w = sph_w(gp.k2, rp) * sph_w(gp.k1, fabs(dz));
    } else {
      const double r = r2 > 0.0 ? r2 * rsqrt_pos(r2) : 0.0;
      w = sph_w_fast(gp.kD, r);
    }
    p.w = in ? w : 0.0;
    p.hit = in ? 1 : 0;
    return in;
  }
  static constexpr int FLAG = 2;  // DeviceView::qflag[2..3]: some record carries a non-zero B, S or Q entry
  // charged = false: every B, S, Q entry of this launch's records is zero; their sums stay exact zeros unread
  __device__ static void accum(const double2* sb, const Prep& p, Acc& a, bool charged) {
    const double2 b1 = sb[1], b2 = sb[2];
    a.cnt += p.hit;
    a.sig += b1.y * p.w;
    a.s += b2.x * p.w;
    if (charged) {
      const double2 b3 = sb[3];
      a.B += b2.y * p.w;
      a.S += b3.x * p.w;
      a.Q += b3.y * p.w;
    }
  }
  __device__ static void store(const DeviceView& dv, const GatherParams& gp, int64_t i, const Acc& a) {
    const double gamma = dv.f[CCB_F_GAMMA][i];
    double sigma, s, rB, rS, rQ;
    if (gp.milne) {
      Metric<true> m(gp.t);
      sigma = to_lrf(m, a.sig, gamma); s = to_lrf(m, a.s, gamma); rB = to_lrf(m, a.B, gamma);
      rS = to_lrf(m, a.S, gamma); rQ = to_lrf(m, a.Q, gamma);
    } else {
      Metric<false> m(gp.t);
      sigma = to_lrf(m, a.sig, gamma); s = to_lrf(m, a.s, gamma); rB = to_lrf(m, a.B, gamma);
      rS = to_lrf(m, a.S, gamma); rQ = to_lrf(m, a.Q, gamma);
    }
    dv.f[CCB_F_SIGMA_LAB][i] = a.sig;
    dv.f[CCB_F_SIGMA][i] = sigma;
    col(dv.f[CCB_F_SMOOTHED], dv.cap, 0, i) = a.s;
    col(dv.f[CCB_F_SMOOTHED], dv.cap, 1, i) = a.B;
    col(dv.f[CCB_F_SMOOTHED], dv.cap, 2, i) = a.S;
    col(dv.f[CCB_F_SMOOTHED], dv.cap, 3, i) = a.Q;
    col(dv.f[CCB_F_THERMO], dv.cap, 5, i) = s;
    col(dv.f[CCB_F_THERMO], dv.cap, 6, i) = rB;
    col(dv.f[CCB_F_THERMO], dv.cap, 7, i) = rS;
    col(dv.f[CCB_F_THERMO], dv.cap, 8, i) = rQ;
    // btrack = number of neighbours (self excluded), sticky at -1 (src/system_state.cpp:712-715)
    if (dv.btrack[i] != -1) dv.btrack[i] = a.cnt - 1;
  }
};

// ------------------------------------------------------------------------------------------
// smooth_gradients: gradP, gradE, gradBulk, gradV, (gradshear, divshear), (div_qa, grad_qa, grad_alpha)
//
// Every sum of src/sph_workstation.cpp:833-984 has the form  sum_b (X_b +- X_a) * c_a * m_b grad_a W_ab  with a target-side
// factor c_a (sigma_a or 1/sigma_a) and, for gradshear / grad_qa, the extra target-side contraction with v_a. With
// A_b = m_b grad_a W_ab the target-side quantities factor out of the neighbour loop:
//     sum_b (X_b +- X_a) c_a A_b  =  c_a ( sum_b X_b A_b  +-  X_a sum_b A_b ).
// The loop therefore accumulates only MOMENTS of the source quantities against A_b (one FMA each) and sum_b A_b; the
// target's own record is not needed until the epilogue. Against the reference's per-pair expression this halves the
// fp64 instructions per pair and frees the 20..36 registers that held the target record. The results differ from the
// per-pair form by rounding only (the sums are re-associated), far inside the 1e-10 parity bar.
//
// QNZ (diffusion runs only): compile-time twin of the run-time sparsity flag DeviceView::qflag[0..1]. QNZ = false is the
// variant for "every record of this launch carries q^mu_a = 0" (kappa = 0 and no initial diffusion: the reference's own
// config_visc_gubser.yaml / input_ampt.yaml): the six q-moment accumulators do not exist, and only the chunks below KQ of a
// record are staged into shared memory. Both variants are launched; the one that does not match the flag exits at once.
template <int D_, bool SHEAR, bool DIFF, bool FACT, bool QNZ = DIFF>
struct GradPolicy {
  using R = GradRec<D_, DIFF>;
  static constexpr int D = D_, REC = R::REC;
  static constexpr int KQ = DIFF ? (R::SIGLAB + 1) / 2 : REC / 2;  // first chunk that holds nothing a source contributes when q^mu_a = 0
  static constexpr int CH_STAGE = (DIFF && !QNZ) ? KQ : REC / 2;            // 16-byte chunks of a record staged per source
  static constexpr int STRIDE = (CH_STAGE % 2 == 1) ? 2 * CH_STAGE : 2 * CH_STAGE + 2;  // shared-memory stride: odd number of chunks
  static constexpr int VARIANT = DIFF ? (QNZ ? 1 : 0) : -1;  // -1: runs whatever the flag says
  static constexpr int MASK_MODE = 2;  // reads the hit masks the field gather of the stage stored
  static constexpr bool EMIT = false;
  struct Target { double x[3]; double v[D]; };
  // fp64 accumulators live across the neighbour loop: decides the register budget of the launch (inst_gather.cu)
  static constexpr int NACC = 4 * D + D * D + ((SHEAR || DIFF) ? 1 : 0) + (SHEAR ? 2 * D : 0) + (DIFF ? 3 * D : 0) + ((DIFF && QNZ) ? 6 : 0);
  struct Acc {
    double sA[D], mP[D], mB[D], mE[D], mV[D][D];
    double sKv, kS0[D], cSh[D];        // shear (sKv also feeds grad_qa)
    double kQ0[3], cq[3], mAL[D][3];   // diffusion (kQ0, cq: QNZ only; otherwise never touched and optimised away)
  };
  __device__ static const double* records(const DeviceView& dv) { return dv.rec_grad; }
  __device__ static void init(Acc& a) {
    a.sKv = 0.;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      a.sA[d] = a.mP[d] = a.mB[d] = a.mE[d] = a.kS0[d] = a.cSh[d] = 0.;
#pragma unroll
      for (int e = 0; e < D; ++e) a.mV[d][e] = 0.;
#pragma unroll
      for (int c = 0; c < 3; ++c) a.mAL[d][c] = 0.;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) a.kQ0[c] = a.cq[c] = 0.;
  }
  __device__ static void load_target(const double* r, Target& t) {
#pragma unroll
    for (int d = 0; d < 3; ++d) t.x[d] = (d < D) ? __ldg(r + R::X + (d < D ? d : 0)) : 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) t.v[d] = __ldg(r + R::V + d);
  }
  // one entry `val` of the source record (index idx, a compile-time constant after unrolling) -> the moments it feeds
  __device__ static __forceinline__ void consume(int idx, double val, const double* A, double Kv, Acc& a) {
    if (idx == R::PT) {
#pragma unroll
      for (int d = 0; d < D; ++d) a.mP[d] = fma(A[d], val, a.mP[d]);
    } else if (idx == R::BT) {
#pragma unroll
      for (int d = 0; d < D; ++d) a.mB[d] = fma(A[d], val, a.mB[d]);
    } else if (idx == R::E) {
#pragma unroll
      for (int d = 0; d < D; ++d) a.mE[d] = fma(A[d], val, a.mE[d]);
    } else if (idx >= R::V && idx < R::V + D) {
#pragma unroll
      for (int e = 0; e < D; ++e)
        if (e == idx - R::V) {
#pragma unroll
          for (int d = 0; d < D; ++d) a.mV[d][e] = fma(A[d], val, a.mV[d][e]);
        }
    } else if (SHEAR && idx >= R::S0 && idx < R::S0 + D) {
#pragma unroll
      for (int d = 0; d < D; ++d)
        if (d == idx - R::S0) a.kS0[d] = fma(Kv, val, a.kS0[d]);
    } else if (SHEAR && idx >= R::SM && idx < R::SM + R::NSM) {
      // pi^{de} / sigma^2 with d <= e feeds divshear_d (with A_e) and, off the diagonal, divshear_e (with A_d)
#pragma unroll
      for (int d = 0; d < D; ++d)
#pragma unroll
        for (int e = d; e < D; ++e)
          if (R::sym(d, e) == idx - R::SM) {
            a.cSh[d] = fma(A[e], val, a.cSh[d]);
            if (e != d) a.cSh[e] = fma(A[d], val, a.cSh[e]);
          }
    } else if (DIFF && idx >= R::ALPHA && idx < R::ALPHA + 3) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (c == idx - R::ALPHA) {
#pragma unroll
          for (int d = 0; d < D; ++d) a.mAL[d][c] = fma(A[d], val, a.mAL[d][c]);
        }
    } else if (DIFF && QNZ && idx >= R::Q0 && idx < R::Q0 + 3) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (c == idx - R::Q0) a.kQ0[c] = fma(Kv, val, a.kQ0[c]);
    } else if (DIFF && QNZ && idx >= R::QI && idx < R::QI + 3 * D) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int d = 0; d < D; ++d)
          if (c * D + d == idx - R::QI) a.cq[c] = fma(A[d], val, a.cq[c]);
    }
  }
  struct Prep { double A[D], Kv; };
  // exact predicate + A_b = m_b grad_a W_ab; false for a pair that is not in the neighbour list (nothing else is computed then)
  __device__ static bool prep(const GatherParams& gp, const Target& t, const double2* sb, bool valid, Prep& p) {
    const double2 b0 = sb[0], b1 = sb[1];
    double rel[3];
    rel[0] = t.x[0] - b0.x; rel[1] = t.x[1] - b0.y;
    rel[2] = (D == 3) ? t.x[2] - b1.x : 0.0;
    const double r2 = (D == 3) ? dist_sqr_exact(rel[0], rel[1], rel[2]) : __dadd_rn(__dmul_rn(rel[0], rel[0]), __dmul_rn(rel[1], rel[1]));
    // outside the neighbour list, or coincident (self): grad W(0) = 0 and every pair term carries grad W
    if (!(valid && (r2 <= gp.rsqr) && r2 != 0.0)) return false;
    const double m_b = (R::M == 2) ? b1.x : b1.y;
    if (FACT) {
      double gK[D];
      grad_kernel<D, FACT>(gp, rel, r2, gK);
#pragma unroll
      for (int d = 0; d < D; ++d) p.A[d] = m_b * gK[d];
    } else {
      double r, ir;
      r_and_inv(r2, r, ir);
      const double g = m_b * sph_gradw_factor_fast(gp.kD, r, ir);
#pragma unroll
      for (int d = 0; d < D; ++d) p.A[d] = g * rel[d];
    }
    p.Kv = 0.;
#pragma unroll
    for (int d = 0; d < D; ++d) p.Kv = fma(p.A[d], t.v[d], p.Kv);
    return true;
  }
  // the record is consumed 16 bytes at a time, in layout order: no more than one chunk is live at any point
  // QNZ = false: every record of this launch carries q^mu_a = 0 (DeviceView::qflag), so the chunks from Q0 on (Q0, QI,
  // padding) contribute exact zeros and are neither staged, loaded nor accumulated.
  static constexpr int FLAG = 0;  // DeviceView::qflag[0..1]
  __device__ static void accum(const double2* sb, const Prep& p, Acc& a, bool) {
#pragma unroll
    for (int d = 0; d < D; ++d) a.sA[d] += p.A[d];
    if (SHEAR || DIFF) a.sKv += p.Kv;
    if (R::M == 2) consume(3, sb[1].y, p.A, p.Kv, a);
#pragma unroll
    for (int k = 2; k < KQ; ++k) {
      const double2 v = sb[k];
      consume(2 * k, v.x, p.A, p.Kv, a);
      consume(2 * k + 1, v.y, p.A, p.Kv, a);
    }
    if (DIFF && QNZ) {
#pragma unroll
      for (int k = KQ; k < REC / 2; ++k) {
        const double2 v = sb[k];
        consume(2 * k, v.x, p.A, p.Kv, a);
        consume(2 * k + 1, v.y, p.A, p.Kv, a);
      }
    }
  }
  __device__ static void store(const DeviceView& dv, const GatherParams&, int64_t i, const Acc& a) {
    // epilogue: the target's own record enters here, once
    double ra[REC];
    {
      const double2* q = reinterpret_cast<const double2*>(dv.rec_grad + i * REC);
#pragma unroll
      for (int k = 0; k < REC / 2; ++k) { const double2 v = __ldg(q + k); ra[2 * k] = v.x; ra[2 * k + 1] = v.y; }
    }
    const double sig = ra[R::SIGLAB], inv_sig = 1.0 / sig;
    double gP[3] = {0., 0., 0.}, gE[3] = {0., 0., 0.}, gB[3] = {0., 0., 0.}, gSh[3] = {0., 0., 0.}, dSh[3] = {0., 0., 0.};
    double gV[3][3] = {{0.}}, gal[3][3] = {{0.}}, dq[3] = {0., 0., 0.}, gq[3] = {0., 0., 0.};
#pragma unroll
    for (int d = 0; d < D; ++d) {
      gP[d] = sig * (a.mP[d] + ra[R::PT] * a.sA[d]);
      gB[d] = sig * (a.mB[d] + ra[R::BT] * a.sA[d]);
      gE[d] = inv_sig * (a.mE[d] - ra[R::E] * a.sA[d]);
#pragma unroll
      for (int e = 0; e < D; ++e) gV[d][e] = inv_sig * (a.mV[d][e] - ra[R::V + e] * a.sA[d]);
      if (SHEAR) {
        gSh[d] = sig * (a.kS0[d] + ra[R::S0 + d] * a.sKv);
        double x = a.cSh[d];
#pragma unroll
        for (int e = 0; e < D; ++e) x = fma(ra[R::SM + R::sym(d, e)], a.sA[e], x);
        dSh[d] = sig * x;
      }
      if (DIFF) {
#pragma unroll
        for (int c = 0; c < 3; ++c) gal[d][c] = inv_sig * (a.mAL[d][c] - ra[R::ALPHA + (DIFF ? c : 0)] * a.sA[d]);
      }
    }
    if (DIFF && QNZ) {  // !QNZ: every q^mu_a of the launch is zero, and so are these two sums
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        gq[c] = sig * (a.kQ0[c] + ra[R::Q0 + (DIFF ? c : 0)] * a.sKv);
        double x = a.cq[c];
#pragma unroll
        for (int d = 0; d < D; ++d) x = fma(ra[R::QI + (DIFF ? c * D + d : 0)], a.sA[d], x);
        dq[c] = sig * x;
      }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      col(dv.f[CCB_F_GRADP], dv.cap, d, i) = gP[d];
      col(dv.f[CCB_F_GRADE], dv.cap, d, i) = gE[d];
      col(dv.f[CCB_F_GRADBULK], dv.cap, d, i) = gB[d];
      col(dv.f[CCB_F_GRADSHEAR], dv.cap, d, i) = gSh[d];
      col(dv.f[CCB_F_DIVSHEAR], dv.cap, d, i) = dSh[d];
#pragma unroll
      for (int e = 0; e < 3; ++e) col(dv.f[CCB_F_GRADV], dv.cap, d * 3 + e, i) = gV[d][e];
#pragma unroll
      for (int c = 0; c < 3; ++c) col(dv.f[CCB_F_GRAD_ALPHA], dv.cap, d * 3 + c, i) = gal[d][c];
      col(dv.f[CCB_F_DIV_QA], dv.cap, d, i) = dq[d];
      col(dv.f[CCB_F_GRAD_QA], dv.cap, d, i) = gq[d];
    }
  }
};

// ------------------------------------------------------------------------------------------
// The tiled driver shared by both gathers. NT = threads (= targets) per task, MINB = CTAs per SM the
// register allocation is bounded for.
constexpr int kMaxSeg = 9;

// Ampere-style asynchronous copies (LDGSTS): 16 bytes global -> shared without a register round trip. The fall-back staging path
// (CCB_STAGE_BULK=0); the default path is the bulk copy below. Neither is a TMA *tensor* box: the records of a tile are an
// interleaved sample of the candidates (every ntiles-th), each a separate 64..304-byte row of the record array.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
// Bulk asynchronous copies (the TMA engine, UBLKCP): ONE instruction moves a whole staged record (64 .. 304 contiguous bytes)
// global -> shared and reports its bytes to an mbarrier; the tile is complete when the barrier's transaction count is met.
// Replaces 4 .. 19 LDGSTS per record and the cp.async group wait. Built with -DCCB_STAGE_BULK=1; NOT the default: measured on
// all five configs (profiles/r02u_bulk_staging_experiment.txt) the engine's per-operation cost loses against LDGSTS for records
// this small -- field gather (64-byte records) +9..13 %, gradient gather -1.6 % (192-byte records) to +4 % (112-byte records).
#ifndef CCB_STAGE_BULK
#define CCB_STAGE_BULK 0
#endif
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)),
               "l"(gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Packed fp32 FMA of sm_100 (FFMA2): two candidates of the pre-filter per instruction. Each half is an IEEE fp32 fma, so the masks
// are those of the scalar form bit for bit.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  ra = *reinterpret_cast<unsigned long long*>(&a); rb = *reinterpret_cast<unsigned long long*>(&b); rc = *reinterpret_cast<unsigned long long*>(&c);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

template <class P, int NT, int MINB, int TILE = kGatherTile>
__global__ void __launch_bounds__(NT, MINB) gather_kernel(DeviceView dv, CellGrid g, GatherParams gp) {
  constexpr int D = P::D, NW = TILE / 32, CH = P::CH_STAGE, RCH = P::REC / 2, RS2 = P::STRIDE / 2;
  // boundary launch of a slab run: a failed wait for the neighbours (p2p_wait -> CCB_ERR_COMM) must not be followed by reads of
  // ghost records that never became ready
  if (gp.task_sel == 2 && dv.err[0] != 0) return;
  // compile-time sparsity variants (GradPolicy QNZ): only the launch that matches the run-time flag does the work
  if (P::VARIANT >= 0) {
    const bool nz = (dv.qflag[P::FLAG] | dv.qflag[P::FLAG + 1]) != 0;
    if (nz != (P::VARIANT == 1)) return;
  }
  __shared__ __align__(16) double2 srec2[2][TILE * RS2];  // double-buffered tiles of source records
  __shared__ __align__(16) float spos[4][TILE];  // fp32 x, y, z relative to the task origin, and x^2 + y^2 + z^2
  __shared__ uint32_t seg_start[kMaxSeg], seg_pref[kMaxSeg + 1];
  __shared__ const double2* seg_base[kMaxSeg];  // record array of the segment: own, or a neighbour's (ghost column)
  __shared__ int s_nseg, s_task;
#if CCB_STAGE_BULK
  __shared__ __align__(8) unsigned long long tile_bar[2];  // one transaction barrier per tile buffer
  if (threadIdx.x == 0) {
    mbar_init(&tile_bar[0], 1); mbar_init(&tile_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  uint32_t staged = 0;  // tiles this CTA has staged so far: tile number g lives in buffer g & 1 and completes phase (g >> 1) & 1 of its barrier
  __syncthreads();
#endif
  const double* __restrict__ rec = P::records(dv);
  const double2* __restrict__ rec2 = reinterpret_cast<const double2*>(rec);
  const int nx = g.nc[0], ny = g.nc[1], nz = g.nc[2];
  const int F = (nz > 1) ? nz : ny;
  const bool qnz = (dv.qflag[P::FLAG] | dv.qflag[P::FLAG + 1]) != 0;  // run-time sparsity flag of the field records (FieldsPolicy::accum)

  // Tasks (cells.cuh: task list): a task = up to NT consecutive particles of ONE cell, so that every target of a CTA has the same
  // 3^D-cell stencil; the list is in cell order: [tasks of the slab's first cell column | interior | last column]. The CTAs of a
  // launch are persistent and pull tasks from a counter, neighbouring tasks running at the same time (their candidates share L2).
  const int* __restrict__ ti = gp.task_info;
  int a0 = 0, a1 = ti[0], b0 = 0, b1 = 0;
  if (gp.task_sel == 1) { a0 = ti[1]; a1 = ti[2]; }                      // interior
  else if (gp.task_sel == 2) { a1 = ti[1]; b0 = ti[2]; b1 = ti[0]; }     // the two boundary columns
  const int na = a1 - a0, count = na + (b1 - b0);

  for (;;) {
    __syncthreads();  // the previous task's tiles, segment table and s_task are no longer in use
    if (threadIdx.x == 0) s_task = atomicAdd(gp.task_counter, 1);
    __syncthreads();
    const int q = s_task;
    if (q >= count) break;
    const int64_t first = dv.tasks[q < na ? a0 + q : b0 + (q - na)];
    const uint32_t key_f = dv.key[first];
    const int64_t last = min(first + (int64_t)NT, (int64_t)dv.cell_start[key_f + 1]) - 1;
    const int64_t i = first + threadIdx.x;
    const bool active = i <= last;
    typename P::Target tg;
    P::load_target(rec + (active ? i : first) * P::REC, tg);
    const int fa = (int)(key_f % (uint32_t)F);
    const uint32_t rho = key_f / (uint32_t)F;  // row of cells along the fastest digit: contiguous in the sorted order
    typename P::Acc acc;
    P::init(acc);
    int emitted = 0;  // EMIT only
    const int f0 = max(fa - 1, 0), f1 = min(fa + 1, F - 1);
    // neighbouring rows: (x +- 1, y +- 1) for a 3-D grid (rho = x ny + y), x +- 1 for a 2-D one (rho = x)
    const int ry = (nz > 1) ? (int)(rho % ny) : 0, rx = (nz > 1) ? (int)(rho / ny) : (int)rho;
    const int nrow_y = (nz > 1) ? 3 : 1;
    // frame of the fp32 pre-filter: coordinates relative to the corner of the task's cell, so that every coordinate in
    // play is two cell edges at most
    double org[3];
    float tx[3];
    {
      // grid arrays are in KEY order (slowest digit first); CellGrid::axis maps a digit to its spatial coordinate
      const int c3[3] = {rx, (nz > 1) ? ry : fa, (nz > 1) ? fa : 0};
      double ok[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) ok[k] = g.min[k] + c3[k] / g.rdx[k];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        org[d] = (d < D) ? (g.axis[0] == d ? ok[0] : (g.axis[1] == d ? ok[1] : ok[2])) : 0.0;
        tx[d] = (float)(tg.x[d] - org[d]);
      }
    }
    // squared extent of the candidates and targets about `org` (+-1 cell along every axis in use, +1 for the cell itself)
    float ext2;
    {
      const float e0 = (float)(2.0 / g.rdx[0]), e1 = (float)(2.0 / g.rdx[1]), e2 = (D == 3) ? (float)(2.0 / g.rdx[2]) : 0.f;
      ext2 = e0 * e0 + e1 * e1 + e2 * e2;
    }
    if (threadIdx.x == 0) {
      int ns = 0;
      uint32_t C = 0;
      for (int dx = -1; dx <= 1; ++dx)
        for (int dy = 0; dy < nrow_y; ++dy) {
          const int x = rx + dx, y = (nz > 1) ? ry + dy - 1 : 0;
          if (x < 0 || x >= nx || y < 0 || y >= ((nz > 1) ? ny : 1)) continue;
          const uint32_t row = (nz > 1) ? (uint32_t)(x * ny + y) : (uint32_t)x;
          const uint32_t* cs = dv.cell_start;
          const double2* base = rec2;
          uint32_t koff = 0;  // ghost columns have their own table, indexed from the column's first cell
          if (x == dv.ghost_col[0]) { cs = dv.ghost_start[0]; koff = (uint32_t)x * ny * nz; base = reinterpret_cast<const double2*>(dv.ghost_rec[0]); }
          else if (x == dv.ghost_col[1]) { cs = dv.ghost_start[1]; koff = (uint32_t)x * ny * nz; base = reinterpret_cast<const double2*>(dv.ghost_rec[1]); }
          const uint32_t js = cs[row * F + f0 - koff], je = cs[row * F + f1 + 1 - koff];
          if (je > js) { seg_start[ns] = js; seg_pref[ns] = C; seg_base[ns] = base; C += je - js; ++ns; }
        }
      seg_pref[ns] = C;
      s_nseg = ns;
    }
    __syncthreads();
    const int nseg = s_nseg;
    const uint32_t C = seg_pref[nseg];
    const uint32_t ntiles = (C + TILE - 1) / TILE;
    // tile k = candidates k, k + ntiles, ... < C: a uniform sample of the neighbourhood, so that every lane finds about
    // the same number of hits in it (contiguous tiles [k TILE, (k + 1) TILE) measured 1.7x slower: the walk of phase 2
    // costs max-over-lanes(hits) per tile). Tile k + 1 is in flight (cp.async) while tile k is processed.
    // Staging: record r of tile k is candidate r * ntiles + k. NT / TILE threads share a record and copy alternating 16-byte
    // chunks of it, so that the segment search and the address arithmetic run once per thread and tile.
#if CCB_STAGE_BULK
    const uint32_t g0 = staged;   // global number of this task's tile 0
    auto stage = [&](uint32_t k) {
      const uint32_t slot = (g0 + k) & 1u;
      double2* buf = srec2[slot];
      const int nt = (int)((C - k + ntiles - 1) / ntiles);
      if (threadIdx.x == 0) mbar_expect_tx(&tile_bar[slot], (unsigned)nt * (unsigned)(CH * 16));
      for (int r = threadIdx.x; r < nt; r += NT) {
        const uint32_t cand = (uint32_t)r * ntiles + k;
        int sg = 0;
        while (cand >= seg_pref[sg + 1]) ++sg;
        const double2* src = seg_base[sg] + ((int64_t)seg_start[sg] + (cand - seg_pref[sg])) * RCH;
        bulk_g2s(buf + r * RS2, src, CH * 16, &tile_bar[slot]);
      }
    };
    if (ntiles) stage(0);
    staged += ntiles;
#else
    auto stage = [&](uint32_t k, double2* buf) {
      constexpr int TPR = (NT >= TILE) ? NT / TILE : 1;
      const int nt = (int)((C - k + ntiles - 1) / ntiles);
      for (int r = threadIdx.x / TPR; r < nt; r += NT / TPR) {
        const uint32_t cand = (uint32_t)r * ntiles + k;
        int sg = 0;
        while (cand >= seg_pref[sg + 1]) ++sg;
        const double2* src = seg_base[sg] + ((int64_t)seg_start[sg] + (cand - seg_pref[sg])) * RCH;
        double2* dst = buf + r * RS2;
#pragma unroll
        for (int c = threadIdx.x % TPR; c < CH; c += TPR) cp_async16(dst + c, src + c);
      }
      cp_async_commit();
    };
    if (ntiles) stage(0, srec2[0]);
#endif
    // hit masks of the stage (GatherParams::hitmask): written by the field gather, read back by the gradient gather
    const int kmask = gp.hitmask ? gp.hit_kmax : 0;
    unsigned long long* hm = gp.hitmask + (active ? i : first);
    unsigned long long mnext = 0ull;  // MASK_MODE 2: the mask of the next tile, loaded one tile ahead
    if (P::MASK_MODE == 2 && kmask > 0 && ntiles) mnext = active ? __ldg(hm) : 0ull;
    for (uint32_t k = 0; k < ntiles; ++k) {
      const int nt = (int)((C - k + ntiles - 1) / ntiles);
      const bool reuse = P::MASK_MODE == 2 && (int)k < kmask;  // CTA-uniform
#if CCB_STAGE_BULK
      const double2* srec = srec2[(g0 + k) & 1u];
      mbar_wait(&tile_bar[(g0 + k) & 1u], ((g0 + k) >> 1) & 1u);   // every byte of tile k has landed
      __syncthreads();  // tile k - 1 (the other buffer, spos) is fully consumed by every thread
      if (k + 1 < ntiles) stage(k + 1);
#else
      const double2* srec = srec2[k & 1];
      cp_async_wait_all();
      __syncthreads();  // tile k has landed for every thread, and tile k - 1 (buffer (k + 1) & 1, spos) is fully consumed
      if (k + 1 < ntiles) stage(k + 1, srec2[(k + 1) & 1]);
#endif
      uint32_t m[NW];
      if (reuse) {
        m[0] = (uint32_t)mnext;
        if (NW > 1) m[NW - 1] = (uint32_t)(mnext >> 32);
        mnext = 0ull;
        if ((int)(k + 1) < kmask && k + 1 < ntiles && active) mnext = __ldg(hm + (int64_t)(k + 1) * gp.hit_stride);
      } else {
        for (int r = threadIdx.x; r < TILE; r += NT) {  // fp32 relative positions of tile k
          if (r < nt) {
            const double2 v0 = srec[r * RS2];
            const float xf = (float)(v0.x - org[0]), yf = (float)(v0.y - org[1]);
            float zf = 0.f;
            if (D == 3) zf = (float)(srec[r * RS2 + 1].x - org[2]);
            spos[0][r] = xf; spos[1][r] = yf;
            if (D == 3) spos[2][r] = zf;
            spos[3][r] = fmaf(xf, xf, fmaf(yf, yf, zf * zf));
          } else {
            spos[0][r] = 0.f; spos[1][r] = 0.f; spos[2][r] = 0.f; spos[3][r] = 3.0e38f;
          }
        }
        __syncthreads();
        // ---- phase 1: fp32 pre-filter, hit masks. |a - b|^2 <= R^2 is tested in the expanded form
        //   |b|^2 - 2 a.b <= R^2 - |a|^2   (|b|^2 staged per source, the right-hand side formed once per lane):
        // D FMAs + 1 compare per candidate instead of D subtractions, D multiply-adds and a compare. The expansion
        // cancels, so the cut-off carries an absolute margin for the fp32 rounding of coordinates up to `ext` (the
        // extent of the stencil around its origin): no true neighbour is missed; a false positive costs one exact test.
        const float ax = -2.f * tx[0], ay = -2.f * tx[1], az = -2.f * tx[2];
        const float2 ax2 = make_float2(ax, ax), ay2 = make_float2(ay, ay), az2 = make_float2(az, az);
        const float thr = (gp.rsqr_f + 4.0e-6f * ext2) - fmaf(tx[0], tx[0], fmaf(tx[1], tx[1], tx[2] * tx[2]));
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          uint32_t mw = 0u;
          if (w * 32 < nt) {
#pragma unroll
            for (int b4 = 0; b4 < 8; ++b4) {
              const float4 X = reinterpret_cast<const float4*>(&spos[0][w * 32])[b4];
              const float4 Y = reinterpret_cast<const float4*>(&spos[1][w * 32])[b4];
              const float4 S = reinterpret_cast<const float4*>(&spos[3][w * 32])[b4];
              float4 Z = make_float4(0.f, 0.f, 0.f, 0.f);
              if (D == 3) Z = reinterpret_cast<const float4*>(&spos[2][w * 32])[b4];
              // v = |b|^2 + az z + ay y + ax x for four candidates, two per FFMA2, in the order of the scalar expression
              float2 lo = make_float2(S.x, S.y), hi = make_float2(S.z, S.w);
              if (D == 3) { lo = ffma2(az2, make_float2(Z.x, Z.y), lo); hi = ffma2(az2, make_float2(Z.z, Z.w), hi); }
              lo = ffma2(ax2, make_float2(X.x, X.y), ffma2(ay2, make_float2(Y.x, Y.y), lo));
              hi = ffma2(ax2, make_float2(X.z, X.w), ffma2(ay2, make_float2(Y.z, Y.w), hi));
              if (lo.x <= thr) mw |= (1u << (b4 * 4 + 0));
              if (lo.y <= thr) mw |= (1u << (b4 * 4 + 1));
              if (hi.x <= thr) mw |= (1u << (b4 * 4 + 2));
              if (hi.y <= thr) mw |= (1u << (b4 * 4 + 3));
            }
          }
          m[w] = active ? mw : 0u;
        }
      }
      // MASK_MODE 1: the exact mask of this tile = the pre-filter's hits minus those the fp64 predicate rejects below
      uint32_t mex[NW];
#pragma unroll
      for (int w = 0; w < NW; ++w) mex[w] = m[w];
      // ---- phase 2: every lane walks its own hits (a software-pipelined, branch-free variant of this loop measured
      // 5-10 % slower: the loop is bound by shared-memory gather bandwidth, not by per-warp latency)
      int mine = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) mine += __popc(m[w]);
      const int maxc = __reduce_max_sync(0xffffffffu, mine);
      for (int it = 0; it < maxc; ++it) {
        if (it < mine) {
          uint32_t wsel = m[NW - 1];
          int base = (NW - 1) * 32;
#pragma unroll
          for (int w = NW - 2; w >= 0; --w)
            if (m[w]) { wsel = m[w]; base = w * 32; }
          const uint32_t low = wsel & (0u - wsel);
          const int j = base + __ffs(wsel) - 1;
#pragma unroll
          for (int w = 0; w < NW; ++w)
            if (base == w * 32) m[w] ^= low;
          typename P::Prep pp;
          if (P::prep(gp, tg, srec + j * RS2, true, pp)) {
            P::accum(srec + j * RS2, pp, acc, qnz);
            if (P::EMIT) {  // debug twin: the id of every neighbour visited (self excluded), in visiting order
              const uint32_t cand = (uint32_t)j * ntiles + k;
              int sg = 0;
              while (cand >= seg_pref[sg + 1]) ++sg;
              const int64_t gidx = (int64_t)seg_start[sg] + (cand - seg_pref[sg]);
              if (gidx != i) gp.nbr_idx[gp.nbr_off[i] + emitted++] = dv.id[gidx];
            }
          } else if (P::MASK_MODE == 1) {
#pragma unroll
            for (int w = 0; w < NW; ++w)
              if (base == w * 32) mex[w] &= ~low;
          }
        }
      }
      if (P::MASK_MODE == 1 && (int)k < kmask && active)
        hm[(int64_t)k * gp.hit_stride] = (unsigned long long)mex[0] | (NW > 1 ? ((unsigned long long)mex[NW - 1] << 32) : 0ull);
    }
    if (active) P::store(dv, gp, i, acc);
  }
}

}  // namespace ccb
