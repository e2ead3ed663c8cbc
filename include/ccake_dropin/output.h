// Drop-in layer: Output<D, TEOM> (reference include/output.h:28-95, src/output.cpp) fed from the device state.
// print_system_state pulls the fields it prints with SystemState::copy_device_to_host and writes the reference's text format
// (src/output.cpp:86-167: one line with t, then one line per particle with the 46 columns listed there); the conserved totals
// come from ONE device reduction (ccb_reduce) instead of a full copy; the freeze-out records come from the device's results
// rows. HDF5 output is not available in this image (no libhdf5): hdf_evolution is rejected.
// The inverse Reynolds numbers and the bulk / diffusion numbers of EoM::compute_hydro_numbers (src/eom_default.cpp:1831-1996) are
// evaluated here from the downloaded fields; the shear Knudsen number needs the sigma tensor, which the fused equations of motion
// never store: it is written as 0.
#ifndef OUTPUT_H
#define OUTPUT_H

#include <cmath>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>

#include "constants.h"
#include "sph_workstation.h"

namespace ccake {

template <unsigned int D, template <unsigned int> class TEOM>
class Output {
 public:
  Output(std::shared_ptr<Settings> settingsPtr_in, std::shared_ptr<SystemState<D>> sys_in, std::shared_ptr<SPHWorkstation<D, TEOM>> ws_in)
      : settingsPtr(settingsPtr_in), systemPtr(sys_in), wsPtr(ws_in), output_directory(settingsPtr_in->results_directory.string()) {}
  ~Output() {}
  void set_results_directory(std::string path) { output_directory = path; }
  std::string get_freeze_out_filename() const { return output_directory + "/freeze_out.dat"; }

  void print_system_state() {
    systemPtr->copy_device_to_host();
    hydro_numbers();
    if (settingsPtr->txt_evolution) print_system_state_to_txt();
    if (settingsPtr->hdf_evolution) throw std::runtime_error("hdf_evolution: HDF5 is not available in this build; use txt_evolution");
    n_timesteps_output++;
  }
  void print_conservation_status() {
    const double Bdiff = (systemPtr->Btotal - systemPtr->Btotal0) / systemPtr->Btotal0 * 100.0;
    const double Sdiff = (systemPtr->Stotal - systemPtr->Stotal0) / systemPtr->Stotal0 * 100.0;
    const double Qdiff = (systemPtr->Qtotal - systemPtr->Qtotal0) / systemPtr->Qtotal0 * 100.0;
    std::cout << "t = " << systemPtr->t << ": " << std::scientific << std::setprecision(6) << systemPtr->Eloss << " " << systemPtr->S << " " << Bdiff << " "
              << Sdiff << " " << Qdiff << " " << std::defaultfloat << "\n"
              << "Btotal = " << systemPtr->Btotal << " Stotal = " << systemPtr->Stotal << " Qtotal = " << systemPtr->Qtotal << " " << std::endl;
  }
  void buffer_conservation_line() {
    cons_buffer << systemPtr->t << " " << systemPtr->Eloss << " " << systemPtr->S << " " << systemPtr->Btotal << " " << systemPtr->Stotal << " "
                << systemPtr->Qtotal << "\n";
  }
  void flush_conservation(const std::string& path) {
    std::ofstream outfile(path.c_str(), std::ios::out | std::ios::trunc);
    if (!outfile) throw std::runtime_error("Could not open " + path);
    outfile << "t Eloss S Btotal Stotal Qtotal\n" << cons_buffer.str();
  }
  void buffer_causality_line() {}                 // the causality check (GSL eigen solver) is out of scope
  void flush_causality(const std::string&) {}
  void print_causality_minimal_to_txt() {}

  // src/output.cpp:417-600: the hypersurface elements frozen out in the last step, buffered; flush_freeze_out writes the file
  void print_freeze_out(std::shared_ptr<FreezeOut<D>> freeze_out, double B, double Q, double S) {
    if (fo_header.empty()) { std::ostringstream hdr; hdr << "# " << B << " " << S << " " << Q << "\n"; fo_header = hdr.str(); }
    freeze_out->fetch_results();
    for (const auto& r : freeze_out->results) {
      if (!r.print) continue;
      fo_buffer << r.divEener << " ";
      for (unsigned int d = 0; d < D; d++) fo_buffer << r.divE[d] << " ";
      fo_buffer << r.gsub << " ";
      for (unsigned int d = 0; d < D; d++) fo_buffer << r.uout[d] << " ";
      fo_buffer << r.swsub << " " << r.bulksub << " " << r.shearsub[0][0] << " " << r.shearsub[1][1] << " " << r.shearsub[2][2] << " " << r.shear33sub << " "
                << r.shearsub[1][2] << " ";
      if (D == 3) fo_buffer << r.shearsub[1][3] << " " << r.shearsub[2][3] << " ";
      fo_buffer << r.tlist << " ";
      for (unsigned int d = 0; d < D; d++) fo_buffer << r.rsub[d] << " ";
      fo_buffer << r.sFO << " " << r.Efluc << " " << r.Tfluc << " " << r.muBfluc << " " << r.muSfluc << " " << r.muQfluc << " " << r.wfzfluc << " "
                << r.cs2fzfluc << " " << r.rhoBfluc << " " << r.rhoSfluc << " " << r.rhoQfluc << " ";
      for (int c = 0; c < 3; ++c) for (int m = 0; m < 4; ++m) fo_buffer << r.diffout[c][m] << " ";
      fo_buffer << r.eos_type << "\n";
      fo_particles_buffered++;
    }
  }
  void flush_freeze_out() {
    std::ofstream FO(get_freeze_out_filename(), std::ios::out | std::ios::trunc);
    if (!FO) throw std::runtime_error("Could not open " + get_freeze_out_filename());
    FO << fo_header << fo_buffer.str();
  }
  long freeze_out_particles_buffered() const { return fo_particles_buffered; }

 private:
  // inverse Reynolds / Knudsen numbers from the downloaded fields (src/eom_default.cpp:1937-1985)
  void hydro_numbers() {
    const double t2 = systemPtr->t * systemPtr->t;
    const bool milne = settingsPtr->coordinate_system != "cartesian";
    const double g33 = milne ? t2 : 1.0;
    for (auto& p : systemPtr->particles) {
      double norm = 0.0;   // pi^{mu nu} pi_{mu nu}
      for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) {
        const double ga = a == 0 ? 1.0 : (a == 3 ? -g33 : -1.0), gb = b == 0 ? 1.0 : (b == 3 ? -g33 : -1.0);
        norm += p.hydro.shv(a, b) * p.hydro.shv(a, b) * ga * gb;
      }
      p.hydro.inverse_reynolds_shear = std::sqrt(std::fabs(norm)) / p.p();
      p.hydro.inverse_reynolds_bulk = std::fabs(p.hydro.bulk) / p.p();
      p.hydro.bulk_knudsen = p.hydro.tau_Pi * std::fabs(p.hydro.theta);
      p.hydro.shear_knudsen = 0.0;
      const double rho[3] = {p.rhoB(), p.rhoS(), p.rhoQ()};
      double re[3];
      for (int c = 0; c < 3; ++c) {
        const double q0 = p.hydro.diffusion(c, 0), q1 = p.hydro.diffusion(c, 1), q2 = p.hydro.diffusion(c, 2), q3 = p.hydro.diffusion(c, 3);
        const double q_sq = q1 * q1 + q2 * q2 + g33 * q3 * q3 - q0 * q0;
        const double n_abs = std::fabs(rho[c]);
        re[c] = (n_abs > 1e-15 && q_sq > 0.0) ? std::sqrt(q_sq) / n_abs : 0.0;
      }
      p.hydro.inverse_reynolds_diffusion_B = re[0]; p.hydro.inverse_reynolds_diffusion_S = re[1]; p.hydro.inverse_reynolds_diffusion_Q = re[2];
    }
  }
  void print_system_state_to_txt() {
    const double hc = constants::hbarc_GeVfm;
    std::ofstream out((output_directory + "/system_state_" + std::to_string(n_timesteps_output) + ".dat").c_str());
    out << systemPtr->t << "\n";
    int iParticle = 0;
    for (auto& p : systemPtr->particles) {
      out << iParticle++ << " " << systemPtr->t << " " << std::setprecision(6) << std::scientific;
      for (unsigned int d = 0; d < D; d++) out << p.r(d) << " ";
      for (unsigned int d = D; d < 3; d++) out << 0.0 << " ";
      out << p.p() * hc << " " << p.T() * hc << " " << p.muB() * hc << " " << p.muS() * hc << " " << p.muQ() * hc << " " << p.e() * hc << " " << p.rhoB() << " "
          << p.rhoS() << " " << p.rhoQ() << " " << p.s() << " " << p.hydro.eta_pi << " " << p.hydro.zeta_Pi << " " << p.hydro.tau_Pi << " " << p.hydro.tau_pi
          << " " << p.hydro.theta << " " << p.hydro.inverse_reynolds_shear << " " << p.hydro.inverse_reynolds_bulk << " " << p.hydro.shear_knudsen << " "
          << p.hydro.bulk_knudsen << " " << p.hydro.shv(0, 0) << " " << p.hydro.shv(1, 1) << " " << p.hydro.shv(2, 2) << " " << p.hydro.shv(0, 1) << " "
          << p.hydro.shv(0, 2) << " " << p.hydro.shv(1, 2) << " " << p.hydro.shv(1, 3) << " " << p.hydro.shv(2, 3) << " " << p.hydro.shv(3, 3) << " ";
      for (unsigned int d = 0; d < D; d++) out << p.hydro.u(d) << " ";
      for (unsigned int d = D; d < 3; d++) out << 0.0 << " ";
      out << p.hydro.gamma << " " << p.Freeze << " " << p.get_current_eos_name() << " " << p.hydro.causality << " " << p.hydro.knudsen_diffusion_B << " "
          << p.hydro.knudsen_diffusion_S << " " << p.hydro.knudsen_diffusion_Q << " " << p.hydro.inverse_reynolds_diffusion_B << " "
          << p.hydro.inverse_reynolds_diffusion_S << " " << p.hydro.inverse_reynolds_diffusion_Q << " " << "\n";
    }
  }

  std::shared_ptr<Settings> settingsPtr;
  std::shared_ptr<SystemState<D>> systemPtr;
  std::shared_ptr<SPHWorkstation<D, TEOM>> wsPtr;
  std::string output_directory;
  int n_timesteps_output = 0;
  std::ostringstream cons_buffer, fo_buffer;
  std::string fo_header;
  long fo_particles_buffered = 0;
};

}  // namespace ccake
#endif
