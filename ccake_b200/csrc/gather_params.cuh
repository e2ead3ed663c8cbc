// Record layouts and the parameter block of the two neighbour gathers (gather.cuh).
#pragma once

#include "state.cuh"

namespace ccb {

constexpr int REC_F = 8;    // x y z m_s | m_s*ext_s m_B*ext_B m_S*ext_S m_Q*ext_Q
constexpr int REC_G0 = 20;  // without diffusion
constexpr int REC_G1 = 36;  // with diffusion
// gradient record slots
enum { RG_X = 0, RG_M = 3, RG_PT = 4, RG_BT = 5, RG_E = 6, RG_V = 7, RG_S0 = 10, RG_SM = 13, RG_SIGLAB = 19,
       RG_ALPHA = 20, RG_Q0 = 23, RG_QI = 26 };

struct GatherParams {
  KernelNorm kD;     // D-dimensional kernel with hT
  KernelNorm k2, k1; // factorised variant: W_2D(r_perp, hT) * W_1D(|d eta|, hEta)
  double rsqr;       // (2 hT)^2: the neighbour cut-off of the reference (src/system_state.cpp:689)
  double t;
  int milne;
};

}  // namespace ccb
