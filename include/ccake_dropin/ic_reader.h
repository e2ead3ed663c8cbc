// Drop-in layer: the initial-condition readers of BSQHydro (src/BSQHydro.cpp:84-145 read_ICCING, :207-425 read_ccake) on the
// library's parallel text parser (ccb_ic_read, csrc/ic_reader.cpp). Two ways to use it:
//   * read_ccake / read_iccing(settings, system): fill SystemState::particles as the reference does (one Particle<D> per row; the
//     reference's own memory footprint, fine up to ~1e6 rows);
//   * ccb_ic_read + ccb_upload_ic directly: the rows never become Particle objects (8-64 M rows).
#ifndef CCAKE_DROPIN_IC_READER_H
#define CCAKE_DROPIN_IC_READER_H

#include <memory>
#include <stdexcept>
#include <string>

#include "constants.h"
#include "system_state.h"

namespace ccake {
namespace b200 {

template <unsigned int D>
void read_ic_file(std::shared_ptr<Settings> s, std::shared_ptr<SystemState<D>> sys, int format) {
  ccb_ic_file* f = nullptr;
  const std::string path = s->IC_file.string();
  if (ccb_ic_read(path.c_str(), format, s->input_as_entropy ? 1 : 0, s->e_cutoff, s->baryon_charge_enabled ? 1 : 0, s->strange_charge_enabled ? 1 : 0,
                  s->electric_charge_enabled ? 1 : 0, s->input_initial_diffusion ? 1 : 0, &f))
    throw std::runtime_error("File " + path + " could not be opened or parsed!");
  double h[6];
  ccb_ic_file_header(f, h);
  s->stepx = h[0]; s->stepy = h[1]; s->xmin = h[3]; s->ymin = h[4];
  if (format == 0) { s->stepEta = h[2]; s->etamin = h[5]; }
  const ccb_ic* ic = ccb_ic_file_view(f);
  const double hc = constants::hbarc_GeVfm;
  for (int64_t i = 0; i < ic->n; ++i) {
    Particle<D> p;
    if (D == 1) { p.r(0) = ic->r[3 * i + 2]; p.hydro.u(0) = ic->u[3 * i + 2]; }   // 1+1D evolves eta (src/BSQHydro.cpp:291-304)
    else for (unsigned int d = 0; d < D; ++d) { p.r(d) = ic->r[3 * i + d]; p.hydro.u(d) = ic->u[3 * i + d]; }
    const double* sh = ic->shv6 + 6 * i;
    p.hydro.shv(1, 1) = sh[0] / hc; p.hydro.shv(1, 2) = p.hydro.shv(2, 1) = sh[1] / hc; p.hydro.shv(1, 3) = p.hydro.shv(3, 1) = sh[2] / hc;
    p.hydro.shv(2, 2) = sh[3] / hc; p.hydro.shv(2, 3) = p.hydro.shv(3, 2) = sh[4] / hc; p.hydro.shv(3, 3) = sh[5] / hc;
    if (s->input_as_entropy) p.input.s = ic->e_or_s[i]; else p.input.e = ic->e_or_s[i] / hc;
    p.input.rhoB = ic->rhoB[i]; p.input.rhoS = ic->rhoS[i]; p.input.rhoQ = ic->rhoQ[i];
    p.hydro.bulk = ic->bulk[i] / hc;
    for (int c = 0; c < 3; ++c) for (int m = 0; m < 4; ++m) p.hydro.diffusion(c, m) = ic->diffusion ? ic->diffusion[12 * i + 4 * c + m] : 0.0;
    sys->add_particle(p);
  }
  ccb_ic_file_free(f);
}
template <unsigned int D> void read_ccake(std::shared_ptr<Settings> s, std::shared_ptr<SystemState<D>> sys) { read_ic_file<D>(s, sys, 0); }
inline void read_iccing(std::shared_ptr<Settings> s, std::shared_ptr<SystemState<2>> sys) { read_ic_file<2>(s, sys, 1); }

}  // namespace b200
}  // namespace ccake
#endif
