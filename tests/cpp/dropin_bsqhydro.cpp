// TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// The reference's UNMODIFIED src/BSQHydro.cpp (constructor, read_in_initial_conditions -> read_ccake / read_ICCING,
// initialize_hydrodynamics, run) compiled against the drop-in headers of include/ccake_dropin/ instead of the reference's
// system_state.h / sph_workstation.h / evolver.h / eom_*.h / eos.h / eos_interpolator.h / freeze_out.h / output.h: the proof that
// the boundary drops in. Built by __graft_entry__.build() when /root/reference is present, into
// tests/libdropin_bsqhydro.so (which travels to the GPU box); the include path is an overlay directory holding the drop-in
// headers plus links to the reference headers that stay the maintainer's own (settings.h, particle.h, kernel.h, ...). Cabana_Core.hpp,
// which particle.h includes, comes from oracle/shim (Cabana is absent from this image). Settings are filled from key=value lines
// (oracle/settings_kv.h; yaml-cpp is absent) -- the same parser the reference build of the oracle uses.
#include <array>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

// BSQHydro keeps its SystemState private and has no accessor; the test reads the particles after run(). Every standard header
// the reference pulls in is included above, so the macro touches the reference's class declarations only.
#define private public
#include "BSQHydro.h"
#undef private
#include "BSQHydro.cpp"   // the reference's translation unit, found on the include path ($(REF)/src)
#include "settings_kv.h"

using namespace ccake;

namespace {
constexpr int ROW = 24;   // doubles per particle handed back (see dropin_bsqhydro_run)

template <unsigned int D, template <unsigned int> class TEOM>
int run_case(std::shared_ptr<Settings> st, long cap, long* n_out, double* t_out, double* rows, int* steps_out) {
  BSQHydro<D, TEOM> sim(st);
  sim.read_in_initial_conditions();
  sim.initialize_hydrodynamics();
  sim.run();
  sim.systemPtr->copy_device_to_host();
  const long n = (long)sim.systemPtr->particles.size();
  *n_out = n;
  *t_out = sim.systemPtr->t;
  *steps_out = sim.systemPtr->number_of_elapsed_timesteps;
  if (n > cap) return 2;
  for (long i = 0; i < n; ++i) {
    const Particle<D>& p = sim.systemPtr->particles[i];
    double* o = rows + (size_t)i * ROW;
    for (unsigned d = 0; d < 3; ++d) { o[d] = d < D ? p.r(d) : 0.0; o[3 + d] = d < D ? p.hydro.u(d) : 0.0; }
    o[6] = p.thermo.T; o[7] = p.thermo.e; o[8] = p.thermo.p; o[9] = p.thermo.s; o[10] = p.thermo.rhoB; o[11] = p.thermo.rhoS; o[12] = p.thermo.rhoQ;
    o[13] = p.hydro.bulk; o[14] = p.hydro.shv(0, 0); o[15] = p.hydro.shv(1, 1); o[16] = p.hydro.shv(2, 2); o[17] = p.hydro.shv(1, 2); o[18] = p.hydro.shv(3, 3);
    o[19] = p.hydro.sigma; o[20] = p.hydro.gamma; o[21] = (double)p.Freeze; o[22] = p.extensive.s; o[23] = p.smoothed.s;
  }
  return 0;
}
}  // namespace

extern "C" {
int dropin_row_width() { return ROW; }
// Runs the whole of ccake.cpp's sequence after Input (src/ccake.cpp:60-120) on the given settings. Returns 0, or 1 on an exception
// (message in err), 2 when `cap` rows are too few. rows: [n][ROW] in particle-id order.
int dropin_bsqhydro_run(const char* kv, long cap, long* n_out, double* t_out, double* rows, int* steps_out, char* err, int errlen) {
  try {
    auto st = settings_kv::parse_settings(kv);
    if (!st) throw std::runtime_error("unknown setting");
    const bool cart = st->coordinate_system == "cartesian";
    if (st->dim == 2) return cart ? run_case<2, EoM_cartesian>(st, cap, n_out, t_out, rows, steps_out) : run_case<2, EoM_default>(st, cap, n_out, t_out, rows, steps_out);
    if (st->dim == 3) return cart ? run_case<3, EoM_cartesian>(st, cap, n_out, t_out, rows, steps_out) : run_case<3, EoM_default>(st, cap, n_out, t_out, rows, steps_out);
    throw std::runtime_error("dim must be 2 or 3");
  } catch (const std::exception& e) {
    std::snprintf(err, (size_t)errlen, "%s", e.what());
    return 1;
  }
}
}
