// Record layouts and the parameter block of the two neighbour gathers (gather.cuh).
//
// A "record" is everything one SOURCE particle contributes to a pair term, packed
// contiguously (array of structures, 16-byte aligned) by the per-particle kernel that runs
// just before the gather, and pre-scaled where the pair formula allows (p / sigma^2 ...), so
// that the gather reads one contiguous block per source instead of ~20 strided columns.
#pragma once

#include "state.cuh"

namespace ccb {

// smooth_fields record: x y z m_s | m_s*ext_s m_B*ext_B m_S*ext_S m_Q*ext_Q
constexpr int REC_F = 8;
constexpr int REC_F_STRIDE = 10;  // shared-memory stride in doubles: an odd number of 16-byte units

// smooth_gradients record, compacted per dimension D (and with/without the diffusion block).
template <int D, bool DIFF> struct GradRec {
  static constexpr int X = 0;                 // position, D doubles
  static constexpr int M = D;                 // sph_mass.s
  static constexpr int PT = D + 1;            // p / sigma_lab^2
  static constexpr int BT = D + 2;            // Pi / sigma_lab^2
  static constexpr int E = D + 3;             // e
  static constexpr int V = D + 4;             // v^i, D doubles
  static constexpr int S0 = 2 * D + 4;        // pi^{0i} / sigma_lab^2, D doubles
  static constexpr int SM = 3 * D + 4;        // pi^{ij} / sigma_lab^2, i <= j, D(D+1)/2 doubles
  static constexpr int NSM = D * (D + 1) / 2;
  // everything a SOURCE contributes comes first (the gather stages only that part when the diffusion currents are zero);
  // sigma_lab is used of the TARGET only and sits behind it
  static constexpr int ALPHA = SM + NSM;                // mu_a / T, 3 doubles (DIFF only)
  static constexpr int SIGLAB = DIFF ? ALPHA + 3 : SM + NSM;   // sigma_lab: 14 | 20 without diffusion, 16 | 22 with it (D = 2 | 3)
  static constexpr int BASE = SIGLAB + 1;
  static constexpr int Q0 = BASE;                       // q^0_a / sigma_lab^2, 3 doubles
  static constexpr int QI = BASE + 3;                   // q^i_a / sigma_lab^2, [a][i], 3 D doubles
  static constexpr int FULL = BASE + 3 + 3 * D;
  static constexpr int REC = DIFF ? ((FULL + 1) / 2) * 2 : ((BASE + 1) / 2) * 2;   // doubles per record in global memory (even)
  static constexpr int STRIDE = ((REC / 2) % 2 == 1) ? REC : REC + 2;      // shared-memory stride: odd number of 16-byte units
  __host__ __device__ static constexpr int sym(int d, int e) { return d <= e ? d * D - d * (d - 1) / 2 + (e - d) : e * D - e * (e - 1) / 2 + (d - e); }
};
constexpr int REC_G_MAX = 36;  // GradRec<3, true>::REC: the record buffer is sized for it

struct GatherParams {
  KernelNorm kD;     // D-dimensional kernel with hT
  KernelNorm k2, k1; // factorised variant: W_2D(r_perp, hT) * W_1D(|d eta|, hEta)
  double rsqr;       // (2 hT)^2: the neighbour cut-off of the reference (src/system_state.cpp:689)
  float rsqr_f;      // the same with a safety margin, for the fp32 pre-filter
  // which tasks a launch covers (DeviceView::tasks is in cell order): task_sel 0 = all, 1 = interior, 2 = the slab's first and
  // last cell column (the targets that need ghost records; launched once the halo has arrived / the neighbours are ready).
  // task_info (device): [0] number of tasks, [1] first interior task, [2] first task of the last column.
  // task_counter (device): the next task of this launch; zeroed before the launch.
  int task_sel;
  const int* task_info;
  int* task_counter;
  double t;
  int milne;
  // The neighbour list of the stage as tile bitmasks: the field gather (phase 03) stores, per target and tile, the mask of the
  // sources that passed the EXACT predicate; the gradient gather (phase 09) of the same stage runs over the same tasks and tiles and
  // reads the masks back instead of repeating the fp32 pre-filter (216 B per particle at 27 tiles of 64, against ~1 KB for a CSR
  // list of 245 neighbours). hitmask[k * hit_stride + i] for tile k < hit_kmax of target i; nullptr = off.
  unsigned long long* hitmask;
  int hit_kmax;
  int64_t hit_stride;
  // debug twin of the field gather (FieldsPolicy EMIT): CSR offsets per target (sorted order) and the id buffer it fills
  const int64_t* nbr_off;
  int32_t* nbr_idx;
};

}  // namespace ccb
