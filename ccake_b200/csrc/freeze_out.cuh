// Freeze-out: the 0 -> 1 -> 3 -> 4 status machine and the hypersurface interpolation, on the device.
// Reference: FreezeOut<D> (include/freeze_out.h:165-933), driven by SPHWorkstation::freeze_out_particles
// (src/sph_workstation.cpp:1475-1485) once per time step when Settings::particlization_enabled, and by the initial
// check_freeze_out_status of SPHWorkstation::initial_smoothing (:450-452).
//
// The reference keeps six per-particle snapshots (frz1, frz2, fback .. fback4: 57 doubles each) and a results record as Cabana
// AoSoAs in particle order. Here the particles are re-sorted by cell every RK stage, so the snapshots are rows of an
// array-of-structures indexed by PARTICLE ID (456-byte rows: a scattered access moves whole sectors); the kernels walk the
// particles in their current order and address the rows through DeviceView::id. Everything is per particle: the reference's
// ~10 masked copy launches per step collapse into three kernels.
#pragma once

#include "local.cuh"

namespace ccb {

// one snapshot row (FRZ, include/freeze_out.h:27-34)
enum {
  FZ_SHEAR = 0,          // [4][4]
  FZ_DIFF = 16,          // [3][4]
  FZ_R = 28, FZ_U = 31, FZ_GRADP = 34, FZ_GRADE = 37,
  FZ_T_ = 40, FZ_S, FZ_E, FZ_RHOB, FZ_RHOS, FZ_RHOQ, FZ_TEMP, FZ_MUB, FZ_MUS, FZ_MUQ, FZ_THETA, FZ_BULK, FZ_SIGLAB, FZ_SHEAR33,
  FZ_SHV_NABLA_U, FZ_WFZ, FZ_CS2FZ,
  FZ_ROW = 57
};
// one results row (ResultsTypes, include/freeze_out.h:35-49); CCB_FO_* in include/ccake_b200.h mirrors these offsets
enum {
  FR_SHEARSUB = 0, FR_DIFFOUT = 16, FR_DIVT = 28, FR_DIVE = 31, FR_DIVP = 34, FR_RSUB = 37, FR_UOUT = 40,
  FR_DIVTTEMP = 43, FR_DIVPPRESS, FR_DIVEENER, FR_GSUB, FR_BULKSUB, FR_THETASUB, FR_SWSUB, FR_SHEAR33SUB, FR_TLIST,
  FR_SFO, FR_EFLUC, FR_TFLUC, FR_MUBFLUC, FR_MUSFLUC, FR_MUQFLUC, FR_CS2FZFLUC, FR_WFZFLUC, FR_RHOBFLUC, FR_RHOSFLUC, FR_RHOQFLUC,
  FR_EOS_TYPE,
  FR_ROW = 64
};

struct FreezeArrays {
  double *frz1, *frz2, *fback, *fback2, *fback3, *fback4;  // [id][FZ_ROW]; nullptr when particlization is off
  double* results;                                           // [id][FR_ROW]
  int32_t* print;                                            // [id]: frozen out in the last call (results row is new)
  unsigned long long* count;                                 // particles that reached status 3 in the last check
};

// the current state of particle i (sorted order) -> snapshot row (freezeout_first_time_step / setup_next_step, :229-262, 601-633);
// the `t` slot is only ever written by check_freeze_out_status
template <int D>
__device__ __forceinline__ void snapshot_row(const DeviceView& dv, int64_t i, double* __restrict__ row) {
#pragma unroll
  for (int a = 0; a < D + 1; ++a)
#pragma unroll
    for (int b = 0; b < D + 1; ++b) row[FZ_SHEAR + a * 4 + b] = col(dv.f[CCB_F_SHV], dv.cap, a * 4 + b, i);
#pragma unroll
  for (int c = 0; c < 12; ++c) row[FZ_DIFF + c] = col(dv.f[CCB_F_DIFFUSION], dv.cap, c, i);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    row[FZ_R + d] = col(dv.f[CCB_F_POSITION], dv.cap, d, i);
    row[FZ_U + d] = col(dv.f[CCB_F_U], dv.cap, d, i);
    row[FZ_GRADP + d] = col(dv.f[CCB_F_GRADP], dv.cap, d, i);
    row[FZ_GRADE + d] = col(dv.f[CCB_F_GRADE], dv.cap, d, i);
  }
  const double* T = dv.f[CCB_F_THERMO];
  row[FZ_S] = T[5 * dv.cap + i]; row[FZ_E] = T[4 * dv.cap + i];
  row[FZ_RHOB] = T[6 * dv.cap + i]; row[FZ_RHOS] = T[7 * dv.cap + i]; row[FZ_RHOQ] = T[8 * dv.cap + i];
  row[FZ_TEMP] = T[i]; row[FZ_MUB] = T[dv.cap + i]; row[FZ_MUS] = T[2 * dv.cap + i]; row[FZ_MUQ] = T[3 * dv.cap + i];
  row[FZ_THETA] = dv.f[CCB_F_THETA][i]; row[FZ_BULK] = dv.f[CCB_F_BULK][i]; row[FZ_SIGLAB] = dv.f[CCB_F_SIGMA_LAB][i];
  row[FZ_SHEAR33] = col(dv.f[CCB_F_SHV], dv.cap, 15, i); row[FZ_SHV_NABLA_U] = dv.f[CCB_F_SHV_NABLA_U][i];
  row[FZ_WFZ] = T[11 * dv.cap + i]; row[FZ_CS2FZ] = T[10 * dv.cap + i];
}
// shift_snapshot (:398-440): every member but the components beyond D of the D-dimensional blocks
template <int D>
__device__ __forceinline__ void shift_row(double* __restrict__ dst, const double* __restrict__ src) {
#pragma unroll
  for (int a = 0; a < D + 1; ++a)
#pragma unroll
    for (int b = 0; b < D + 1; ++b) dst[FZ_SHEAR + a * 4 + b] = src[FZ_SHEAR + a * 4 + b];
#pragma unroll
  for (int c = 0; c < 12; ++c) dst[FZ_DIFF + c] = src[FZ_DIFF + c];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    dst[FZ_R + d] = src[FZ_R + d]; dst[FZ_U + d] = src[FZ_U + d];
    dst[FZ_GRADP + d] = src[FZ_GRADP + d]; dst[FZ_GRADE + d] = src[FZ_GRADE + d];
  }
#pragma unroll
  for (int k = FZ_T_; k < FZ_ROW; ++k) dst[k] = src[k];
}

// FreezeOut::check_freeze_out_status (:462-541). fa.frz1 == nullptr (particlization off: only the initial check runs): the
// snapshot reads / writes are skipped; status 1 cannot occur before the first call, so nothing is lost.
template <int UNUSED>  // (a template only so that the header can be included by several translation units)
__global__ void freeze_check_status_kernel(DeviceView dv, FreezeArrays fa, double time_in, double efcheck) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int mine = 0;
  if (i < dv.n) {
    const int id = dv.id[i];
    const int freeze = dv.freeze[i];
    const double e = col(dv.f[CCB_F_THERMO], dv.cap, 4, i);
    if (freeze == 0) {
      if (e <= efcheck) { dv.freeze[i] = 1; if (fa.frz2) fa.frz2[(int64_t)id * FZ_ROW + FZ_T_] = time_in; }
    } else if (freeze == 1) {
      const double e1 = fa.frz1 ? fa.frz1[(int64_t)id * FZ_ROW + FZ_E] : 0.0;
      if (dv.btrack[i] == -1) { dv.freeze[i] = 3; if (fa.frz1) fa.frz1[(int64_t)id * FZ_ROW + FZ_T_] = time_in; }
      else if (e > e1) { dv.freeze[i] = 1; if (fa.frz2) fa.frz2[(int64_t)id * FZ_ROW + FZ_T_] = time_in; }
      else if (e <= efcheck) { dv.freeze[i] = 3; if (fa.frz1) fa.frz1[(int64_t)id * FZ_ROW + FZ_T_] = time_in; }
      else dv.freeze[i] = 0;
    }
    mine = dv.freeze[i] == 3 ? 1u : 0u;
    if (fa.print) fa.print[id] = 0;
  }
  const unsigned int b = __ballot_sync(0xffffffffu, mine != 0);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(fa.count, (unsigned long long)__popc(b));
}

// which = 2: freezeout_first_time_step (current -> frz2); which = 1: freezeout_second_time_step (current -> frz1)
template <int D>
__global__ void freeze_snapshot_kernel(DeviceView dv, FreezeArrays fa, int which) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dv.n) return;
  snapshot_row<D>(dv, i, (which == 2 ? fa.frz2 : fa.frz1) + (int64_t)dv.id[i] * FZ_ROW);
}

// FreezeOut::freeze_out_evo (:316-392): the back-up snapshots of particles that are losing their neighbours. The reference's
// masked shifts, per particle and in its order.
template <int D>
__global__ void freeze_evo_kernel(DeviceView dv, FreezeArrays fa) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dv.n) return;
  const int freeze = dv.freeze[i], bt = dv.btrack[i];
  if (freeze >= 4) return;
  const int64_t o = (int64_t)dv.id[i] * FZ_ROW;
  if (bt <= 3 && bt > 0) {
    shift_row<D>(fa.fback4 + o, fa.fback2 + o);
    shift_row<D>(fa.fback3 + o, fa.fback + o);
    shift_row<D>(fa.fback2 + o, fa.frz2 + o);
    shift_row<D>(fa.fback + o, fa.frz1 + o);
  } else if (bt == 0) {
    if (fa.fback[o + FZ_GRADP] != 0) {
      shift_row<D>(fa.frz2 + o, fa.fback2 + o);
      shift_row<D>(fa.frz1 + o, fa.fback + o);
    } else {
      shift_row<D>(fa.frz2 + o, fa.fback4 + o);
      shift_row<D>(fa.frz1 + o, fa.fback3 + o);
    }
    dv.freeze[i] = 4;
    dv.btrack[i] = -1;
  }
}

// bsqsvfreezeout, third time step onwards (:585-638): frz2 <- frz1 for particles that are not frozen out, frz1 <- current for all
template <int D>
__global__ void freeze_advance_kernel(DeviceView dv, FreezeArrays fa) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dv.n) return;
  const int64_t o = (int64_t)dv.id[i] * FZ_ROW;
  if (dv.freeze[i] < 4) shift_row<D>(fa.frz2 + o, fa.frz1 + o);
  snapshot_row<D>(dv, i, fa.frz1 + o);
}

// FreezeOut::bsqsvinterpolate (:647-931): the hypersurface element of every particle that reached status 3
template <int D, bool MILNE>
__global__ void freeze_interpolate_kernel(DeviceView dv, FreezeArrays fa, double efcheck, double taup, double taupp, double dt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dv.n) return;
  if (dv.freeze[i] != 3) return;
  const int id = dv.id[i];
  const double* f1 = fa.frz1 + (int64_t)id * FZ_ROW;
  const double* f2 = fa.frz2 + (int64_t)id * FZ_ROW;
  double* R = fa.results + (int64_t)id * FR_ROW;
  // the snapshot closer to the freeze-out energy density
  const double d1 = fabs(f1[FZ_E] - efcheck), d2 = fabs(f2[FZ_E] - efcheck);
  const bool first = d1 < d2;
  const double* f = first ? f1 : f2;
  const double tau_s = first ? taup : taupp;
  const double tlist = dv.btrack[i] != -1 ? tau_s : tau_s - dt;
  R[FR_TLIST] = tlist;
#pragma unroll
  for (int a = 0; a < D + 1; ++a)
#pragma unroll
    for (int b = 0; b < D + 1; ++b) R[FR_SHEARSUB + a * 4 + b] = f[FZ_SHEAR + a * 4 + b];
#pragma unroll
  for (int c = 0; c < 12; ++c) R[FR_DIFFOUT + c] = f[FZ_DIFF + c];
  double uout[D], gradP[D], gradE[D], u_cov[D], gradP_con[D], gradE_con[D], divT_con[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    R[FR_RSUB + d] = f[FZ_R + d];
    uout[d] = f[FZ_U + d]; R[FR_UOUT + d] = uout[d];
    gradP[d] = f[FZ_GRADP + d]; gradE[d] = f[FZ_GRADE + d];
  }
  const Metric<MILNE> m(tlist);
  lower_vec<D, MILNE>(m, uout, u_cov);
  raise_vec<D, MILNE>(m, gradP, gradP_con);
  raise_vec<D, MILNE>(m, gradE, gradE_con);
  R[FR_BULKSUB] = f[FZ_BULK]; R[FR_THETASUB] = f[FZ_THETA];
  const double shv_nabla_u = f[FZ_SHV_NABLA_U];
  double sigsub = f[FZ_SIGLAB];
  R[FR_SHEAR33SUB] = f[FZ_SHEAR33];
  R[FR_EFLUC] = f[FZ_E]; R[FR_MUBFLUC] = f[FZ_MUB]; R[FR_MUSFLUC] = f[FZ_MUS]; R[FR_MUQFLUC] = f[FZ_MUQ];
  double sFO = f[FZ_S], Tfluc = f[FZ_TEMP];
  R[FR_CS2FZFLUC] = f[FZ_CS2FZ]; R[FR_WFZFLUC] = f[FZ_WFZ];
  R[FR_RHOBFLUC] = f[FZ_RHOB]; R[FR_RHOSFLUC] = f[FZ_RHOS]; R[FR_RHOQFLUC] = f[FZ_RHOQ];
  // normals
  double norm2 = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) norm2 += uout[d] * u_cov[d];
  const double gsub = sqrt(-norm2 + 1);
  R[FR_GSUB] = gsub;
  if (MILNE) sigsub /= gsub * tlist; else sigsub /= gsub;
  R[FR_SWSUB] = col(dv.f[CCB_F_SPH_MASS], dv.cap, 0, i) / sigsub;
  double divT[D], divE[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    divT[d] = (1.0 / sFO) * gradP[d];
    R[FR_DIVP + d] = gradP[d];
    divE[d] = gradE[d];
    divT_con[d] = (1.0 / sFO) * gradP_con[d];
  }
  const double cs2 = f[FZ_CS2FZ], w = f[FZ_WFZ], bulk = f[FZ_BULK], theta = f[FZ_THETA];
  double inner = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) inner += uout[d] * gradP[d];
  const double divPpress = -(1.0 / gsub) * (cs2 * (w + bulk) * theta - cs2 * shv_nabla_u + inner);
  R[FR_DIVPPRESS] = divPpress;
  double divTtemp = divPpress / sFO;
  inner = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) inner += uout[d] * gradE[d];
  double divEener = -(1.0 / gsub) * ((w + bulk) * theta - shv_nabla_u + inner);
  norm2 = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) norm2 += divT[d] * divT_con[d];
  double norm = -sqrt(fabs(divTtemp * divTtemp + norm2));
  norm = fabs(norm) > 1.E-14 ? norm : 1.E-14;
  divTtemp /= norm;
#pragma unroll
  for (int d = 0; d < D; ++d) R[FR_DIVT + d] = divT[d] / norm;
  R[FR_DIVTTEMP] = divTtemp;
  norm2 = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) norm2 += divE[d] * gradE_con[d];
  double normE = -sqrt(fabs(divEener * divEener + norm2));
  normE = fabs(normE) > 1.E-14 ? normE : 1.E-14;
  R[FR_DIVEENER] = divEener / normE;
#pragma unroll
  for (int d = 0; d < D; ++d) R[FR_DIVE + d] = divE[d] / normE;
  R[FR_SFO] = sFO * pow(Tfluc * kHbarcGeVfm, 3);
  R[FR_TFLUC] = Tfluc * kHbarcGeVfm;
  R[FR_EOS_TYPE] = dv.eos_type_is_table ? (dv.eos_id[i] == 0 ? 1.0 : 0.0) : 0.0;
  dv.freeze[i] = 4;
  fa.print[id] = 1;
}

}  // namespace ccb
