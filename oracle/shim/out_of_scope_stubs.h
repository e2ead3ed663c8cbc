// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Force-included (-include) ahead of the reference sources when building
// oracle/_ref/. It pre-defines the include guards of three reference headers that
// are OUT OF SCOPE for the hot path (SURVEY.md section 2 rows 14, 15:
// freeze_out.h, jets.h, source.h) and declares inert
// stand-ins with just the members src/sph_workstation.cpp names, so the hot-path
// translation units compile unchanged without dragging HDF5 / jets / sources in.
// Every stand-in aborts if it is ever reached.
#ifndef CCAKE_ORACLE_OUT_OF_SCOPE_STUBS_H
#define CCAKE_ORACLE_OUT_OF_SCOPE_STUBS_H

#define BBMG_H
#define SOURCE_H

#include <Cabana_Core.hpp>

#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <memory>

#include "settings.h"
#include "system_state.h"
#include "thermodynamic_info.h"

namespace ccake {
[[noreturn]] inline void oracle_out_of_scope(const char* what) {
  std::fprintf(stderr, "oracle/_ref: %s is out of scope and was stubbed out\n", what);
  std::abort();
}

template <unsigned int D> class BBMG {
 public:
  BBMG() {}
  BBMG(std::shared_ptr<Settings>, std::shared_ptr<SystemState<D>>) {}
  void initial() { oracle_out_of_scope("BBMG::initial"); }
  void propagate() { oracle_out_of_scope("BBMG::propagate"); }
};

template <unsigned int D> class Source {
 public:
  Source(std::shared_ptr<Settings>, std::shared_ptr<SystemState<D>>) {}
  void set_bbmg(BBMG<D>*) {}
  void deposit_sources() { oracle_out_of_scope("Source::deposit_sources"); }
  void advance_sources() { oracle_out_of_scope("Source::advance_sources"); }
};

}  // namespace ccake
#endif
