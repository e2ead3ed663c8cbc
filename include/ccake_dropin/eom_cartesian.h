// Drop-in layer, part 5 of 8: the TEOM static interface for Cartesian coordinates (reference include/eom_cartesian.h,
// src/eom_cartesian.cpp:18-202).
#ifndef EOM_CARTESIAN_H
#define EOM_CARTESIAN_H

#include <cmath>
#include <memory>
#include <string>

#include "system_state.h"

namespace ccake {

template <unsigned int D>
class EoM_cartesian {
 public:
  EoM_cartesian() = delete;
  EoM_cartesian(std::shared_ptr<Settings> settingsPtr_in) : settingsPtr(settingsPtr_in) {}
  ~EoM_cartesian() {}

  static double gamma_calc(double u[D], const double&) {
    double c = 0.0;
    for (unsigned int i = 0; i < D; ++i) c += u[i] * u[i];
    return std::sqrt(1.0 + c);
  }
  static double get_LRF(const double& lab, const double& gamma, const double&) { return lab / gamma; }   // src/eom_cartesian.cpp:197-202
  static void update_velocity(std::shared_ptr<SystemState<D>> sysPtr) { sysPtr->device->check(ccb_smooth_fields(sysPtr->device->ctx)); }
  static void reset_pi_tensor(std::shared_ptr<SystemState<D>> sysPtr) { sysPtr->device->check(ccb_local_update(sysPtr->device->ctx)); }
  static void evaluate_time_derivatives(std::shared_ptr<SystemState<D>> sysPtr, std::shared_ptr<Settings>) {
    sysPtr->device->check(ccb_time_derivatives(sysPtr->device->ctx));
  }
  static void check_causality(std::shared_ptr<SystemState<D>>) {}
  static constexpr int coordinate_system = CCB_COORD_CARTESIAN;
  std::string name = "Israel-Stewart (Cartesian)";

 private:
  std::shared_ptr<Settings> settingsPtr;
};

}  // namespace ccake
#endif
