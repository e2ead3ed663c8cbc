// Drop-in layer, part 8 of 8: SPHWorkstation<D, TEOM> with the public surface of the reference's class
// (include/sph_workstation.h:53-163, src/sph_workstation.cpp), on top of the C ABI (include/ccake_b200.h).
//
// BSQHydro<D,TEOM>::initialize_hydrodynamics (src/BSQHydro.cpp:569-611) and ::run (:657-814) call these methods in the
// reference's order; the whole of the initialisation runs on the device inside initialize_entropy_and_charge_densities()
// (ccb_upload_ic = process_initial_conditions' device half + initial sigma + e|s -> (T, mu) location + initial smoothing), so the
// calls that follow it there (calculate_gamma_and_velocities, initial_smoothing) find their work done.
#ifndef SPH_WORKSTATION_H
#define SPH_WORKSTATION_H

#include <algorithm>
#include <cmath>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "constants.h"
#include "eom_cartesian.h"
#include "eom_default.h"
#include "eos.h"
#include "eos_interpolator.h"
#include "evolver.h"
#include "freeze_out.h"
#include "phase_profiler.h"
#include "system_state.h"

namespace ccake {

// jets and external sources are out of scope: inert members so that host code naming them compiles
template <unsigned int D> class BBMG {
 public:
  BBMG() {}
  BBMG(std::shared_ptr<Settings>, std::shared_ptr<SystemState<D>>) {}
  void initial() {}
  void propagate() {}
};
template <unsigned int D> class Source {
 public:
  Source(std::shared_ptr<Settings>, std::shared_ptr<SystemState<D>>) {}
  void set_bbmg(BBMG<D>*) {}
  void deposit_sources() {}
  void advance_sources() {}
};

template <unsigned int D, template <unsigned int> class TEOM>
class SPHWorkstation {
 private:
  std::shared_ptr<Settings> settingsPtr;
  std::shared_ptr<SystemState<D>> systemPtr;
  Evolver<D> evolver;
  EquationOfState eos;
  std::shared_ptr<EoS_Interpolator> eos_interpolatorPtr;
  b200::Device& dev() { return *systemPtr->device; }

 public:
  std::shared_ptr<FreezeOut<D>> freezePtr;
  void setup_freeze_out() { freezePtr = std::make_shared<FreezeOut<D>>(settingsPtr, systemPtr); }
  // SPHWorkstation::freeze_out_particles (src/sph_workstation.cpp:1475-1485). advance_timestep already ran it on the device when
  // particlization is enabled (the reference calls it from there too, :2063); an explicit call runs it once more
  void freeze_out_particles() { int k = 0; freezePtr->freeze_out_particles(k); }
  BBMG<D> bbmg;
  std::shared_ptr<Source<D>> sourcePtr;
  void setup_source_terms() { sourcePtr = std::make_shared<Source<D>>(settingsPtr, systemPtr); }

  SPHWorkstation() = delete;
  SPHWorkstation(std::shared_ptr<Settings> settingsPtr_in, std::shared_ptr<SystemState<D>> systemPtr_in)
      : settingsPtr(settingsPtr_in), systemPtr(systemPtr_in), evolver(Evolver<D>(settingsPtr_in, systemPtr_in)), bbmg(BBMG<D>(settingsPtr_in, systemPtr_in)) {}
  ~SPHWorkstation() {}

  // SPHWorkstation::initialize (src/sph_workstation.cpp:38-84): system state, equation of state (forward table or pre-inverted
  // table), transport parameters (inside the config), freeze-out energy density
  void initialize() {
    systemPtr->initialize();
    ccb_config cfg = b200::to_config(*settingsPtr);
    cfg.dim = (int)D;
    cfg.coordinate_system = TEOM<D>::coordinate_system;
    dev().create(cfg);
    eos.set_SettingsPtr(settingsPtr);
    eos.attach(systemPtr->device);
    if (settingsPtr->online_inverter_enabled) {
      if (settingsPtr->eos_type == "table")
        dev().check(ccb_load_eos_table_dat(dev().ctx, (settingsPtr->eos_path + "/thermo.dat").c_str(), settingsPtr->normalize_by_T ? 1 : 0));
      eos.init();
    } else {
      eos_interpolatorPtr = std::make_shared<EoS_Interpolator>(settingsPtr->preinverted_eos_path);
      eos_interpolatorPtr->attach(systemPtr->device);
    }
    dev().check(ccb_efcheck(dev().ctx, &systemPtr->efcheck));
  }

  // the host half of process_initial_conditions (src/sph_workstation.cpp:1229-1340): gamma, unit extensive densities and SPH
  // masses, the root-finder's seed, particle ids
  void process_initial_conditions() {
    if (settingsPtr->IC_type == "freezein") freeze_in_match();   // src/sph_workstation.cpp:1236-1237
    systemPtr->n_particles = (int)systemPtr->particles.size();
    const double t = settingsPtr->t0, t2 = t * t;
    for (auto& p : systemPtr->particles) {
      double u[D];
      for (unsigned int i = 0; i < D; ++i) u[i] = p.hydro.u(i);
      p.hydro.gamma = TEOM<D>::gamma_calc(u, t2);
      p.extensive.s = p.extensive.rhoB = p.extensive.rhoS = p.extensive.rhoQ = 1.0;
      p.sph_mass.s = p.sph_mass.rhoB = p.sph_mass.rhoS = p.sph_mass.rhoQ = 1.0;
      p.thermo.T = 580.0 / constants::hbarc_MeVfm;
      p.thermo.muB = p.thermo.muS = p.thermo.muQ = 0.0;
      p.thermo.eos_name = "default";
      p.efcheck = systemPtr->efcheck;
    }
    for (int i = 0; i < systemPtr->n_particles; i++) { systemPtr->particles[i].ID = i; systemPtr->particles[i].hydro.ID = i; }
  }
  // IC_type "freezein" (SPHWorkstation::freeze_in_match, src/sph_workstation.cpp:1024-1221): the raw T^{tau mu}, j_a^tau that
  // BSQHydro::read_freezein stored on every particle -> (epsilon, u^mu, rho_a) by the fixed-point iteration v <- |M| / (M^0 + P(eps, n))
  // with the pressure closed by the EoS. The reference iterates particle by particle with one host root search per iterate; here
  // all unconverged cells advance together and each iterate is ONE batched device query (the per-particle arithmetic and the
  // convergence test are the reference's, so every cell takes the same iterates).
  void freeze_in_match() {
    if (D < 2) throw std::runtime_error("Freeze-in IC requires D >= 2 (raw T^{mu nu} grids are 2D/3D).");
    const double hc = constants::hbarc_GeVfm;
    const double tau = settingsPtr->t0;
    const bool hyper = settingsPtr->coordinate_system == "hyperbolic";
    const double t2 = (D == 3 && hyper) ? tau * tau : 1.0;
    const double tol = settingsPtr->freezein_tol, v_cap = 1.0 - 1e-12, e_floor = settingsPtr->e_cutoff / hc;
    const int maxit = settingsPtr->freezein_max_iter;
    const bool use[3] = {settingsPtr->baryon_charge_enabled, settingsPtr->strange_charge_enabled, settingsPtr->electric_charge_enabled};
    auto& ps = systemPtr->particles;
    const size_t n = ps.size();
    std::vector<double> M0(n), Mx(n), My(n), Mn(n), Mmag(n), v(n), P(n, 0.0), j0(3 * n);
    std::vector<char> done(n, 0);
    for (size_t i = 0; i < n; ++i) {
      const Particle<D>& p = ps[i];
      M0[i] = p.freezein_Ttau[0] / hc; Mx[i] = p.freezein_Ttau[1] / hc; My[i] = p.freezein_Ttau[2] / hc;
      Mn[i] = (D == 3) ? p.freezein_Ttau[3] / hc : 0.0;
      Mmag[i] = std::sqrt(Mx[i] * Mx[i] + My[i] * My[i] + t2 * Mn[i] * Mn[i]);
      j0[3 * i] = p.freezein_jB[0]; j0[3 * i + 1] = p.freezein_jS[0]; j0[3 * i + 2] = p.freezein_jQ[0];
      v[i] = (M0[i] > 0.0) ? std::min(Mmag[i] / M0[i], v_cap) : 0.0;
    }
    std::vector<size_t> idx;
    std::vector<double> dens, th;
    std::vector<int32_t> member;
    for (int it = 0; it < maxit; ++it) {
      idx.clear(); dens.clear();
      for (size_t i = 0; i < n; ++i) {
        if (done[i]) continue;
        const double eps = std::max(M0[i] - v[i] * Mmag[i], 0.0), gamma = 1.0 / std::sqrt(1.0 - v[i] * v[i]);
        if (eps > e_floor) {
          idx.push_back(i);
          dens.push_back(eps);
          for (int a = 0; a < 3; ++a) dens.push_back(use[a] ? j0[3 * i + a] / gamma : 0.0);
        } else P[i] = 0.0;   // near-vacuum cell
      }
      th.assign(16 * idx.size(), 0.0); member.assign(idx.size(), -1);
      eos.s_out_batch((int64_t)idx.size(), dens.data(), th.data(), member.data());
      // locate_phase_diagram_point_eBSQ: a cell without a solution is located again without its charges (src/sph_workstation.cpp:1808-1832)
      std::vector<size_t> again;
      std::vector<double> dens0;
      for (size_t k = 0; k < idx.size(); ++k)
        if (member[k] < 0) { again.push_back(k); dens0.insert(dens0.end(), {dens[4 * k], 0.0, 0.0, 0.0}); }
      if (!again.empty()) {
        std::vector<double> th0(16 * again.size());
        std::vector<int32_t> m0(again.size());
        eos.s_out_batch((int64_t)again.size(), dens0.data(), th0.data(), m0.data());
        for (size_t q = 0; q < again.size(); ++q)
          if (m0[q] >= 0) { member[again[q]] = m0[q]; for (int c = 0; c < 16; ++c) th[16 * again[q] + c] = th0[16 * q + c]; }
      }
      for (size_t k = 0; k < idx.size(); ++k) if (member[k] >= 0) P[idx[k]] = th[16 * k + 9];   // else: the pressure of the last successful location
      bool all = true;
      for (size_t i = 0; i < n; ++i) {
        if (done[i]) continue;
        const double denom = M0[i] + P[i];
        const double v_new = (denom > 0.0) ? std::min(Mmag[i] / denom, v_cap) : 0.0;
        if (std::abs(v_new - v[i]) < tol) done[i] = 1; else all = false;
        v[i] = v_new;
      }
      if (all) break;
    }
    std::vector<Particle<D>> kept;
    kept.reserve(n);
    for (size_t i = 0; i < n; ++i) {
      const double gamma = 1.0 / std::sqrt(1.0 - v[i] * v[i]);
      const double eps = std::max(M0[i] - v[i] * Mmag[i], 0.0);
      if (eps <= e_floor) continue;   // the decisive cut, on the rest-frame energy density
      Particle<D> p = ps[i];
      const double s = Mmag[i] > 0.0 ? gamma * v[i] / Mmag[i] : 0.0;
      p.input.e = eps;
      p.input.rhoB = use[0] ? j0[3 * i] / gamma : 0.0; p.input.rhoS = use[1] ? j0[3 * i + 1] / gamma : 0.0; p.input.rhoQ = use[2] ? j0[3 * i + 2] / gamma : 0.0;
      p.hydro.u(0) = s * Mx[i];
      if (D >= 2) p.hydro.u(1) = s * My[i];
      if (D == 3) p.hydro.u(2) = s * Mn[i];
      p.hydro.bulk = 0.0; p.hydro.shv = 0.0; p.hydro.diffusion = 0.0;   // ideal IC: the residual stress is discarded
      p.thermo.T = 580.0 / constants::hbarc_MeVfm; p.thermo.muB = p.thermo.muS = p.thermo.muQ = 0.0; p.thermo.eos_name = "default";
      p.Freeze = 0;
      kept.push_back(p);
    }
    systemPtr->particles = std::move(kept);
  }

  // The rows BSQHydro::read_ccake stored in `particles` (src/BSQHydro.cpp:262-403, fm units) go to the device in the units of the
  // IC file, and the library runs initialize_entropy_and_charge_densities + calculate_intial_sigma + initial_smoothing
  // (src/sph_workstation.cpp:253-457) there
  void initialize_entropy_and_charge_densities() {
    const double hc = constants::hbarc_GeVfm;
    const size_t n = systemPtr->particles.size();
    std::vector<double> r(3 * n, 0.0), e(n), rB(n), rS(n), rQ(n), u(3 * n, 0.0), bulk(n), shv6(6 * n), q(12 * n);
    for (size_t i = 0; i < n; ++i) {
      const Particle<D>& p = systemPtr->particles[i];
      for (unsigned int d = 0; d < D; ++d) { r[3 * i + d] = p.r(d); u[3 * i + d] = p.hydro.u(d); }
      e[i] = settingsPtr->input_as_entropy ? p.input.s : p.input.e * hc;
      rB[i] = p.input.rhoB; rS[i] = p.input.rhoS; rQ[i] = p.input.rhoQ;
      bulk[i] = p.hydro.bulk * hc;
      const double s6[6] = {p.hydro.shv(1, 1), p.hydro.shv(1, 2), p.hydro.shv(1, 3), p.hydro.shv(2, 2), p.hydro.shv(2, 3), p.hydro.shv(3, 3)};
      for (int k = 0; k < 6; ++k) shv6[6 * i + k] = s6[k] * hc;
      for (int c = 0; c < 3; ++c) for (int m = 0; m < 4; ++m) q[12 * i + 4 * c + m] = p.hydro.diffusion(c, m);
    }
    ccb_ic ic;
    std::memset(&ic, 0, sizeof(ic));
    ic.n = (int64_t)n; ic.r = r.data(); ic.e_or_s = e.data(); ic.rhoB = rB.data(); ic.rhoS = rS.data(); ic.rhoQ = rQ.data(); ic.u = u.data();
    ic.bulk = bulk.data(); ic.shv6 = shv6.data();
    if (settingsPtr->input_initial_diffusion) ic.diffusion = q.data();
    dev().check(ccb_upload_ic(dev().ctx, &ic));
    systemPtr->on_device = true;
    systemPtr->t = ccb_time(dev().ctx);
    evolver.allocate_cache();
  }
  void calculate_gamma_and_velocities() { if (systemPtr->on_device) dev().check(ccb_smooth_fields(dev().ctx)); }   // phases 02 (+ 03, idempotent)
  void initial_smoothing() {}                       // done by initialize_entropy_and_charge_densities (ccb_upload_ic), incl. the initial freeze-out check
  void initialize_jets_bbmg() {}
  void add_toy_jet() {}
  void calculate_intial_sigma(double) { dev().check(ccb_smooth_fields(dev().ctx)); }

  // ---- the phases of get_time_derivatives (src/sph_workstation.cpp:1543-1583) -------------------------------------------------
  void reset_pi_tensor(double) { TEOM<D>::reset_pi_tensor(systemPtr); }                                  // 05 (+ 06, 08: one kernel)
  void regulate_diffusion() {}                                                                         // 06: inside ccb_local_update
  void smooth_all_particle_fields(double) { dev().check(ccb_smooth_fields(dev().ctx)); }                // 02 + 03
  void smooth_all_particle_gradients(double) { dev().check(ccb_smooth_gradients(dev().ctx)); }          // 09
  void update_all_particle_thermodynamics() { dev().check(ccb_update_thermodynamics(dev().ctx)); }      // 04
  void update_all_particle_viscosities() { dev().check(ccb_local_update(dev().ctx)); }                  // 08 (with 05, 06)
  void get_time_derivatives(double) { dev().check(ccb_get_time_derivatives(dev().ctx)); }

  // ---- host-side EoS look-ups on one particle (src/sph_workstation.cpp:659-741, 1791-1848) --------------------------------------
  void locate_phase_diagram_point_sBSQ(Particle<D>& p, double s_In, double rhoB_In, double rhoS_In, double rhoQ_In) {
    eos.tbqs(p.T(), p.muB(), p.muQ(), p.muS(), p.get_current_eos_name());
    if (eos.update_s(s_In, rhoB_In, rhoS_In, rhoQ_In, p.print_this_particle)) eos.set_thermo(p.thermo);
  }
  void locate_phase_diagram_point_sBSQ(Particle<D>& p, double s_In) { locate_phase_diagram_point_sBSQ(p, s_In, 0.0, 0.0, 0.0); }
  double locate_phase_diagram_point_eBSQ(Particle<D>& p, double e_In, double rhoB_In, double rhoS_In, double rhoQ_In) {
    bool found = false;
    double sVal = eos.s_out(e_In, rhoB_In, rhoS_In, rhoQ_In, found, p.print_this_particle);
    if (!found) {   // freeze the particle out and try again without charge densities (src/sph_workstation.cpp:1808-1832)
      p.Freeze = 4;
      sVal = eos.s_out(e_In, 0.0, 0.0, 0.0, found, p.print_this_particle);
    }
    if (found) eos.set_thermo(p.thermo);
    return sVal;
  }
  double locate_phase_diagram_point_eBSQ(Particle<D>& p, double e_In) { return locate_phase_diagram_point_eBSQ(p, e_In, 0.0, 0.0, 0.0); }

  void set_bulk_Pi() {}               // folded into the device initialisation (Settings::bulk_from_trace)
  void calculate_extensive_shv() {}   // likewise (Settings::input_initial_shear)
  void set_diffusion() {}             // likewise (Settings::input_initial_diffusion)

  // src/sph_workstation.cpp:1993-2024
  bool continue_evolution() {
    std::cout << "t = " << systemPtr->t << std::endl;
    bool keep_going = true;
    if (settingsPtr->particlization_enabled) {
      ccb_totals tt;
      dev().check(ccb_reduce(dev().ctx, &tt));
      systemPtr->number_part_fo = (int)tt.n_frozen;
      keep_going = (systemPtr->number_part_fo < systemPtr->n_particles);
    }
    keep_going = keep_going && (systemPtr->t <= settingsPtr->max_tau);
    return keep_going;
  }
  // src/sph_workstation.cpp:2043-2073: the RK step (Evolver) and, when particlization is enabled, the freeze-out of the step, which
  // the library runs at the end of ccb_step
  void advance_timestep(double dt, int rk_order) {
    PhaseProfiler& prof = PhaseProfiler::instance();   // CCAKE_PROFILE / CCAKE_PROFILE_EVERY (include/phase_profiler.h:35-140)
    if (prof.enabled() && !profiling) { dev().check(ccb_profile_enable(dev().ctx, 1)); profiling = true; }
    evolver.execute_timestep(dt, rk_order, [this, dt] { this->get_time_derivatives(dt); });
    systemPtr->number_of_elapsed_timesteps++;
    if (profiling) {   // device time of the step per reference phase label -> the reference's own profiler table
      for (int ph = 0; ph < ccb_profile_phase_count(); ++ph) {
        double ms = 0.0;
        int64_t calls = 0;
        dev().check(ccb_profile_get(dev().ctx, ph, &ms, &calls));
        if (calls > prof_calls[ph]) prof.add(ccb_profile_phase_name(ph), 1e-3 * (ms - prof_ms[ph]));
        prof_ms[ph] = ms; prof_calls[ph] = calls;
      }
    }
  }

 private:
  bool profiling = false;
  double prof_ms[16] = {0};
  int64_t prof_calls[16] = {0};

 public:
};

}  // namespace ccake
#endif
