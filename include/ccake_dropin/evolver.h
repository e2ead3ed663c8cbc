// Drop-in layer, part 3 of 8: Evolver<D> (reference include/evolver.h:58-116, src/evolver.cpp:98-561).
// The Runge-Kutta stage updates are fused into the device step (ccb_step: RK2 midpoint, or the reference's five-evaluation RK4
// sequence), so the time-derivative functional the reference hands over is not called: the library evaluates the same phases
// itself. It is only checked for being callable.
#ifndef EVOLVER_H
#define EVOLVER_H

#include <functional>
#include <memory>

#include "system_state.h"

namespace ccake {

template <unsigned int D>
class Evolver {
 public:
  Evolver() = delete;
  Evolver(std::shared_ptr<Settings> settingsPtr_in, std::shared_ptr<SystemState<D>> systemPtr_in)
      : settingsPtr(settingsPtr_in), systemPtr(systemPtr_in) {}

  // src/evolver.cpp:98-116: rk_order 2 or 4, anything else is an error there too (exit(8))
  void execute_timestep(double dt, int rk_order, std::function<void(void)> time_derivatives_functional) {
    if (!time_derivatives_functional) throw std::invalid_argument("Evolver::execute_timestep: empty time-derivative functional");
    if (rk_order != 2 && rk_order != 4) throw std::invalid_argument("Invalid Runge-Kutta order!");
    if (rk_order != settingsPtr->rk_order)
      throw std::invalid_argument("Evolver::execute_timestep: rk_order differs from Settings::rk_order, which sized the device caches");
    b200::Device& dev = *systemPtr->device;
    dev.check(ccb_step(dev.ctx, dt));
    systemPtr->t = ccb_time(dev.ctx);
  }
  void allocate_cache() {}      // the y0 snapshot and the RK4 accumulator live in the library (allocated with the particles)
  void allocate_k_values() {}
  void set_current_timestep_quantities() {}

 private:
  std::shared_ptr<Settings> settingsPtr;
  std::shared_ptr<SystemState<D>> systemPtr;
};

}  // namespace ccake
#endif
