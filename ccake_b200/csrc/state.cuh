// Device-side state of one context: a slim structure-of-arrays of fp64 columns
// (one cudaMalloc'd slab), physically re-sorted by neighbour cell every RK stage.
// Replaces the reference's 22-member, ~4.5 KB/particle Cabana AoSoA
// (include/particle.h:54-100).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/ccake_b200.h"
#include "physics_eom.cuh"

namespace ccb {

// Width (doubles per particle) of every public field, indexed by ccb_field.
static const int kFieldWidth[CCB_F_COUNT] = {
    3, 3, 4, 4, 1, 5, 9, 1, 4,        // position .. tseed (persistent, permuted every stage)
    3, 1, 1, 1, 4, 16, 12, 1, 16, 12, 13,  // v .. transport
    3, 3, 3, 3, 3, 9, 3, 3, 9,        // gradients
    3, 4, 1, 5, 9, 1, 1, 1,           // time derivatives
    2};                               // chiBB, chiBB0
static const int kNumPersistentFields = 9;  // CCB_F_POSITION .. CCB_F_TSEED
// evolved by the Runge-Kutta update; each keeps a y0 snapshot (reference EvolverCache, include/evolver.h:21-47)
static const int kEvolvedFields[] = {CCB_F_POSITION, CCB_F_U, CCB_F_EXTENSIVE, CCB_F_EXTENSIVE_BULK,
                                     CCB_F_EXTENSIVE_SHV, CCB_F_EXTENSIVE_DIFFUSION, CCB_F_EZ};
static const int kNumEvolvedFields = 7;

// Neighbour grid of the reference (src/system_state.cpp:651-700) with its arrays in KEY order: digit 0 is the slowest digit
// of the cell key, digit 2 the fastest; axis[k] is the spatial coordinate (0 x, 1 y, 2 z / eta) digit k indexes. The slab
// axis (ccb_config::slab_axis) is digit 0, so that a slab's edge column is one contiguous range of the cell-sorted order;
// the other coordinates follow in ascending order (a 2-D run's unused third coordinate is always digit 2, one cell).
struct CellGrid {
  double min[3], rdx[3];
  int nc[3];
  int axis[3];
  int64_t ncells;
};

// Everything a kernel needs, passed by value.
struct DeviceView {
  int64_t n;
  double* f[CCB_F_COUNT];      // f[field] -> first component column; component c at f[field] + c * cap
  double* y0[CCB_F_COUNT];     // snapshot columns for evolved fields (nullptr otherwise)
  int64_t cap;
  int32_t *id, *btrack, *freeze, *eos_id;
  uint32_t* key;               // cell id per particle (current order)
  uint32_t* cell_start;        // ncells + 1
  double* rec_fields;          // 8 doubles / particle:  x y z m_s  m_s*ext_s m_B*ext_B m_S*ext_S m_Q*ext_Q
  double* rec_grad;            // REC_GRAD doubles / particle, see gather kernels
  int* qflag;                  // [0] some own gradient record carries a non-zero diffusion current q^mu_a, [1] some ghost record does;
                               // [2], [3] likewise for a non-zero B, S or Q entry of the field records
  int32_t* tasks;              // gather tasks in cell order: first particle of every chunk of <= NT particles of one cell
  int32_t* err;                // [0] status code, [1] first offending particle id
  int eos_type_is_table;       // the run's EoS chain starts with the table (thermodynamic_info::eos_type_is_table of member 0)
  // slab decomposition: the cell columns (index along x) just outside this rank's slab hold GHOST records,
  // appended to rec_fields / rec_grad behind the n own records; ghost_start[s][c] .. [c + 1] is the record range
  // of cell c of that column (c = y * nz + z). ghost_col = -1: no neighbour on that side.
  int32_t ghost_col[2];
  const uint32_t* ghost_start[2];
  // where the ghost records of each side live: behind the own records (halo exchanged by message: the own array again), or
  // the NEIGHBOUR's record array itself, mapped over NVLink peer memory (ghost_start then holds the neighbour's own offsets)
  const double* ghost_rec[2];
};

CCB_HD double& col(double* base, int64_t cap, int c, int64_t i) { return base[(int64_t)c * cap + i]; }

}  // namespace ccb
