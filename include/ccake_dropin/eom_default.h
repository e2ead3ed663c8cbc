// Drop-in layer, part 4 of 8: the TEOM static interface (reference include/eom_default.h:27-50, src/eom_default.cpp) for
// hyperbolic (Milne) coordinates. gamma_calc and get_LRF are host arithmetic (the reference calls them on the host while it
// reads the initial conditions); the three phase functions forward to the C ABI.
#ifndef EOM_DEFAULT_H
#define EOM_DEFAULT_H

#include <cmath>
#include <memory>
#include <string>

#include "system_state.h"

namespace ccake {

template <unsigned int D>
class EoM_default {
 public:
  EoM_default() = delete;
  EoM_default(std::shared_ptr<Settings> settingsPtr_in) : settingsPtr(settingsPtr_in) {}
  ~EoM_default() {}

  // gamma = sqrt(1 - u_i u^i) with the Milne lowering (-, -, -t^2); D = 2 never applies t^2 (src/eom_default.cpp:24-37, milne.hpp:353-354)
  static double gamma_calc(double u[D], const double& time_squared) {
    double c = 0.0;
    for (unsigned int i = 0; i < D; ++i) c += u[i] * u[i] * ((D == 3 && i == D - 1) || D == 1 ? -time_squared : -1.0);
    return std::sqrt(1.0 - c);
  }
  // lab frame -> local rest frame: lab / gamma / t. The header names the third argument time_squared, but the implementation divides by
  // it as it stands and every caller passes t (src/eom_default.cpp:199-203, src/sph_workstation.cpp:564-568): kept that way
  static double get_LRF(const double& lab, const double& gamma, const double& t) { return lab / gamma / t; }
  static void update_velocity(std::shared_ptr<SystemState<D>> sysPtr) { sysPtr->device->check(ccb_smooth_fields(sysPtr->device->ctx)); }
  static void reset_pi_tensor(std::shared_ptr<SystemState<D>> sysPtr) { sysPtr->device->check(ccb_local_update(sysPtr->device->ctx)); }
  static void evaluate_time_derivatives(std::shared_ptr<SystemState<D>> sysPtr, std::shared_ptr<Settings>) {
    sysPtr->device->check(ccb_time_derivatives(sysPtr->device->ctx));
  }
  static void check_causality(std::shared_ptr<SystemState<D>>) {}   // out of scope (GSL eigen solver on the host)
  static constexpr int coordinate_system = CCB_COORD_HYPERBOLIC;
  std::string name = "Israel-Stewart";

 private:
  std::shared_ptr<Settings> settingsPtr;
};

}  // namespace ccake
#endif
