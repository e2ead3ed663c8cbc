// Drop-in layer, part 2 of 8: SystemState<D> with the public surface of the reference's class
// (include/system_state.h:40-127, src/system_state.cpp), on top of the C ABI (include/ccake_b200.h).
//
// `particles` stays the host-side std::vector<Particle<D>> the reference's readers, Output and BSQHydro work on;
// copy_host_to_device / copy_device_to_host marshal it field by field in PARTICLE-ID order (the library scatters by id, as
// src/system_state.cpp:344 does). The Cabana AoSoA and the Verlet list do not exist here: the device state is the library's.
#ifndef SYSTEM_STATE_H
#define SYSTEM_STATE_H

#include <array>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "device_context.h"
#include "particle.h"

namespace ccake {

template <unsigned int D>
class SystemState {
 public:
  SystemState() = delete;
  SystemState(std::shared_ptr<Settings> settingsPtr_in) : device(std::make_shared<b200::Device>()), settingsPtr(settingsPtr_in) {}
  ~SystemState() {}

  double t = 0.0;
  double S = 0.0, S0 = 0.0, Btotal = 0.0, Btotal0 = 0.0, Stotal = 0.0, Stotal0 = 0.0, Qtotal = 0.0, Qtotal0 = 0.0;
  double E = 0.0, Ez = 0.0, E0 = 0.0, Eloss = 0.0, dEz = 0.0, Etot = 0.0, Ejet = 0.0, Ejet0 = 0.0;
  double efcheck = 0.0, sfcheck = 0.0;
  int number_part_fo = 0;
  int n_particles = 0;
  int number_of_elapsed_timesteps = 0;
  std::vector<Particle<D>> particles;
  // eccentricity histories (include/system_state.h:93-96): the members exist so that BSQHydro::run compiles; the observable itself is
  // out of scope, compute_eccentricities() records nothing and the output loop over eta_slices is empty
  std::vector<double> timesteps;
  std::vector<std::vector<double>> e_2_P_history_by_slice, e_2_X_history_by_slice;
  std::vector<std::vector<int>> count_P_history_by_slice, count_X_history_by_slice;
  std::vector<double> eta_slices;

  // the device context of the run (created by SPHWorkstation::initialize); `on_device`: the particle system lives there
  std::shared_ptr<b200::Device> device;
  bool on_device = false;

  void add_particle(Particle<D> p) { particles.push_back(p); }                 // src/system_state.cpp:557-560
  void initialize() { number_of_elapsed_timesteps = 0; t = settingsPtr->t0; }  // :37-52
  // the device arrays are allocated with the upload of the initial conditions (SPHWorkstation::initialize_entropy_and_charge_densities)
  void allocate_cabana_particles() { n_particles = (int)particles.size(); if (on_device) copy_host_to_device(); }
  void initialize_linklist() {}                                                // the neighbour grid is fixed at ccb_create
  void reset_neighbour_list() { device->check(ccb_build_cells(device->ctx)); } // :651-719
  int n() { return n_particles; }
  double get_particle_T(int id) { return particles[id].T(); }
  double get_particle_Freeze(int id) { return particles[id].Freeze; }
  void set_EquationOfStatePtr(void*) {}
  void set_SettingsPtr(Settings*) {}

  // Settings::get_neighbors diagnostics of BSQHydro::run (include/system_state.h:136-205). The reference indexes its device arrays,
  // which are in particle-id order; here the id-ordered getters of the C ABI do the same. The neighbour count is btrack, the number
  // of particles within 2h that the field gather of the last stage visited (the length of the reference's Verlet list entry).
  std::vector<std::array<double, 4>> get_particle_data(int idx) {
    const size_t n = (size_t)ccb_n(device->ctx);
    std::vector<double> pos(3 * n);
    std::vector<int32_t> nb(n);
    device->check(ccb_get_field(device->ctx, CCB_F_POSITION, pos.data()));
    device->check(ccb_get_int_field(device->ctx, CCB_I_BTRACK, nb.data()));
    return {{(double)nb[idx], pos[3 * idx], pos[3 * idx + 1], pos[3 * idx + 2]}};
  }
  std::vector<int> entropy_density_based_idx_search() {   // ids of the particles with the smallest and the largest smoothed entropy density
    const size_t n = (size_t)ccb_n(device->ctx);
    std::vector<double> sm(4 * n);
    device->check(ccb_get_field(device->ctx, CCB_F_SMOOTHED, sm.data()));
    int idx_min = 0, idx_max = 0;
    for (size_t i = 1; i < n; ++i) {
      if (sm[4 * i] < sm[4 * idx_min]) idx_min = (int)i;
      if (sm[4 * i] > sm[4 * idx_max]) idx_max = (int)i;
    }
    return {idx_min, idx_max};
  }
  void print_neighbors(int idx) {
    auto d = get_particle_data(idx);
    std::cout << "System has " << ccb_n(device->ctx) << " particles" << std::endl << "Particle " << idx << " # neighbor = " << (int)d[0][0] << std::endl;
  }

  // ---- particles -> device: everything the device state holds between steps (src/system_state.cpp:72-326) -------------------
  void copy_host_to_device() {
    if (!on_device) return;
    const size_t n = particles.size();
    std::vector<double> b(n * 16);
    auto put = [&](int field, int w, auto&& get) {
      for (size_t i = 0; i < n; ++i) for (int c = 0; c < w; ++c) b[i * w + c] = get(particles[i], c);
      device->check(ccb_set_field(device->ctx, field, b.data()));
    };
    put(CCB_F_POSITION, 3, [](const Particle<D>& p, int c) { return c < (int)D ? p.r(c) : 0.0; });
    put(CCB_F_U, 3, [](const Particle<D>& p, int c) { return c < (int)D ? p.hydro.u(c) : 0.0; });
    put(CCB_F_EXTENSIVE, 4, [](const Particle<D>& p, int c) { return c == 0 ? p.extensive.s : c == 1 ? p.extensive.rhoB : c == 2 ? p.extensive.rhoS : p.extensive.rhoQ; });
    put(CCB_F_SPH_MASS, 4, [](const Particle<D>& p, int c) { return c == 0 ? p.sph_mass.s : c == 1 ? p.sph_mass.rhoB : c == 2 ? p.sph_mass.rhoS : p.sph_mass.rhoQ; });
    put(CCB_F_EXTENSIVE_BULK, 1, [](const Particle<D>& p, int) { return p.hydro.extensive_bulk; });
    put(CCB_F_EXTENSIVE_SHV, 5, [](const Particle<D>& p, int c) { return c < 3 ? p.hydro.extensive_shv(0, c) : p.hydro.extensive_shv(1, c - 2); });
    put(CCB_F_EXTENSIVE_DIFFUSION, 9, [](const Particle<D>& p, int c) { return p.hydro.extensive_diffusion(c / 3, c % 3); });
    put(CCB_F_EZ, 1, [](const Particle<D>& p, int) { return p.contribution_to_total_Ez; });
    put(CCB_F_TSEED, 4, [](const Particle<D>& p, int c) { return c == 0 ? p.thermo.T : c == 1 ? p.thermo.muB : c == 2 ? p.thermo.muS : p.thermo.muQ; });
    std::vector<int32_t> ib(n);
    for (size_t i = 0; i < n; ++i) ib[i] = particles[i].Freeze;
    device->check(ccb_set_int_field(device->ctx, CCB_I_FREEZE, ib.data()));
    for (size_t i = 0; i < n; ++i) ib[i] = particles[i].btrack;
    device->check(ccb_set_int_field(device->ctx, CCB_I_BTRACK, ib.data()));
    device->check(ccb_set_time(device->ctx, t));
  }

  // ---- device -> particles (src/system_state.cpp:335-555): every field Output and the readers of `particles` consume --------
  void copy_device_to_host() {
    if (!on_device) return;
    const size_t n = particles.size();
    std::vector<double> b(n * 16);
    auto get = [&](int field, int w, auto&& set) {
      device->check(ccb_get_field(device->ctx, field, b.data()));
      for (size_t i = 0; i < n; ++i) for (int c = 0; c < w; ++c) set(particles[i], c, b[i * w + c]);
    };
    const int Di = (int)D;
    get(CCB_F_POSITION, 3, [&](Particle<D>& p, int c, double v) { if (c < Di) p.r(c) = v; });
    get(CCB_F_U, 3, [&](Particle<D>& p, int c, double v) { if (c < Di) p.hydro.u(c) = v; });
    get(CCB_F_V, 3, [&](Particle<D>& p, int c, double v) { if (c < Di) p.hydro.v(c) = v; });
    get(CCB_F_EXTENSIVE, 4, [](Particle<D>& p, int c, double v) { (c == 0 ? p.extensive.s : c == 1 ? p.extensive.rhoB : c == 2 ? p.extensive.rhoS : p.extensive.rhoQ) = v; });
    get(CCB_F_SPH_MASS, 4, [](Particle<D>& p, int c, double v) { (c == 0 ? p.sph_mass.s : c == 1 ? p.sph_mass.rhoB : c == 2 ? p.sph_mass.rhoS : p.sph_mass.rhoQ) = v; });
    get(CCB_F_SMOOTHED, 4, [](Particle<D>& p, int c, double v) { (c == 0 ? p.smoothed.s : c == 1 ? p.smoothed.rhoB : c == 2 ? p.smoothed.rhoS : p.smoothed.rhoQ) = v; });
    get(CCB_F_D_EXTENSIVE_DT, 4, [](Particle<D>& p, int c, double v) { (c == 0 ? p.d_dt_extensive.s : c == 1 ? p.d_dt_extensive.rhoB : c == 2 ? p.d_dt_extensive.rhoS : p.d_dt_extensive.rhoQ) = v; });
    get(CCB_F_EXTENSIVE_BULK, 1, [](Particle<D>& p, int, double v) { p.hydro.extensive_bulk = v; });
    get(CCB_F_EXTENSIVE_SHV, 5, [](Particle<D>& p, int c, double v) { if (c < 3) p.hydro.extensive_shv(0, c) = v; else p.hydro.extensive_shv(1, c - 2) = v; });
    get(CCB_F_EXTENSIVE_DIFFUSION, 9, [](Particle<D>& p, int c, double v) { p.hydro.extensive_diffusion(c / 3, c % 3) = v; });
    get(CCB_F_EZ, 1, [](Particle<D>& p, int, double v) { p.contribution_to_total_Ez = v; });
    get(CCB_F_DEZ, 1, [](Particle<D>& p, int, double v) { p.contribution_to_total_dEz = v; });
    get(CCB_F_GAMMA, 1, [](Particle<D>& p, int, double v) { p.hydro.gamma = v; });
    get(CCB_F_SIGMA_LAB, 1, [](Particle<D>& p, int, double v) { p.hydro.sigma_lab = v; });
    get(CCB_F_SIGMA, 1, [](Particle<D>& p, int, double v) { p.hydro.sigma = v; });
    get(CCB_F_BULK, 1, [](Particle<D>& p, int, double v) { p.hydro.bulk = v; });
    get(CCB_F_THETA, 1, [](Particle<D>& p, int, double v) { p.hydro.theta = v; });
    get(CCB_F_SHV_NABLA_U, 1, [](Particle<D>& p, int, double v) { p.hydro.shv_nabla_u = v; });
    get(CCB_F_D_BULK_DT, 1, [](Particle<D>& p, int, double v) { p.hydro.d_extensive_bulk_dt = v; });
    get(CCB_F_SHV, 16, [](Particle<D>& p, int c, double v) { p.hydro.shv(c / 4, c % 4) = v; });
    get(CCB_F_DIFFUSION, 12, [](Particle<D>& p, int c, double v) { p.hydro.diffusion(c / 4, c % 4) = v; });
    get(CCB_F_THERMO, 16, [](Particle<D>& p, int c, double v) {
      thermodynamic_info& t = p.thermo;
      double* slot[16] = {&t.T, &t.muB, &t.muS, &t.muQ, &t.e, &t.s, &t.rhoB, &t.rhoS, &t.rhoQ, &t.p, &t.cs2, &t.w, &t.dwds, &t.dwdB, &t.dwdS, &t.dwdQ};
      *slot[c] = v;
    });
    get(CCB_F_DALPHA, 12, [](Particle<D>& p, int c, double v) {
      thermodynamic_info& t = p.thermo;
      double* slot[12] = {&t.dalpha_Bds, &t.dalpha_BdB, &t.dalpha_BdS, &t.dalpha_BdQ, &t.dalpha_Sds, &t.dalpha_SdB, &t.dalpha_SdS, &t.dalpha_SdQ,
                          &t.dalpha_Qds, &t.dalpha_QdB, &t.dalpha_QdS, &t.dalpha_QdQ};
      *slot[c] = v;
    });
    get(CCB_F_CHIBB, 2, [](Particle<D>& p, int c, double v) { (c == 0 ? p.thermo.chiBB : p.thermo.chiBB0) = v; });
    get(CCB_F_TRANSPORT, 13, [](Particle<D>& p, int c, double v) {
      auto& h = p.hydro;
      double* slot[13] = {&h.eta_pi, &h.tau_pi, &h.zeta_Pi, &h.tau_Pi, &h.delta_PiPi, &h.delta_pipi, &h.lambda_piPi, &h.lambda_Pipi, &h.phi1, &h.phi3,
                          &h.phi6, &h.phi7, &h.tau_pipi};
      *slot[c] = v;
    });
    get(CCB_F_GRADP, 3, [&](Particle<D>& p, int c, double v) { if (c < Di) p.hydro.gradP(c) = v; });
    get(CCB_F_GRADE, 3, [&](Particle<D>& p, int c, double v) { if (c < Di) p.hydro.gradE(c) = v; });
    get(CCB_F_GRADBULK, 3, [&](Particle<D>& p, int c, double v) { if (c < Di) p.hydro.gradBulk(c) = v; });
    get(CCB_F_GRADSHEAR, 3, [&](Particle<D>& p, int c, double v) { if (c < Di) p.hydro.gradshear(c) = v; });
    get(CCB_F_DIVSHEAR, 3, [&](Particle<D>& p, int c, double v) { if (c < Di) p.hydro.divshear(c) = v; });
    get(CCB_F_GRADV, 9, [&](Particle<D>& p, int c, double v) { if (c / 3 < Di && c % 3 < Di) p.hydro.gradV(c / 3, c % 3) = v; });
    get(CCB_F_DIV_QA, 3, [](Particle<D>& p, int c, double v) { p.hydro.div_qa(c) = v; });
    get(CCB_F_GRAD_QA, 3, [](Particle<D>& p, int c, double v) { p.hydro.grad_qa(c) = v; });
    get(CCB_F_GRAD_ALPHA, 9, [&](Particle<D>& p, int c, double v) { if (c / 3 < Di) p.hydro.grad_alpha_a(c / 3, c % 3) = v; });
    get(CCB_F_DU_DT, 3, [&](Particle<D>& p, int c, double v) { if (c < Di) p.hydro.du_dt(c) = v; });
    get(CCB_F_D_SHV_DT, 5, [](Particle<D>& p, int c, double v) { if (c < 3) p.hydro.d_extensive_shv_dt(0, c) = v; else p.hydro.d_extensive_shv_dt(1, c - 2) = v; });
    get(CCB_F_D_DIFFUSION_DT, 9, [](Particle<D>& p, int c, double v) { p.hydro.d_extensive_q_dt(c / 3, c % 3) = v; });
    std::vector<int32_t> ib(n);
    device->check(ccb_get_int_field(device->ctx, CCB_I_FREEZE, ib.data()));
    for (size_t i = 0; i < n; ++i) particles[i].Freeze = ib[i];
    device->check(ccb_get_int_field(device->ctx, CCB_I_BTRACK, ib.data()));
    for (size_t i = 0; i < n; ++i) particles[i].btrack = ib[i];
    device->check(ccb_get_int_field(device->ctx, CCB_I_EOS, ib.data()));
    const bool table = settingsPtr->eos_type == "table";
    static const char* chain[4] = {"table", "tanh_conformal", "conformal", "conformal_diagonal"};
    for (size_t i = 0; i < n; ++i) {
      particles[i].thermo.eos_name = table ? chain[ib[i] < 0 || ib[i] > 3 ? 0 : ib[i]] : settingsPtr->eos_type;
      particles[i].thermo.eos_type_is_table = (table && ib[i] == 0) ? 1.0 : 0.0;
      particles[i].hydro.t = t;
    }
    t = ccb_time(device->ctx);
  }
  // the online inverter's partial copies (:562-650) have no role here: phase 04 runs on the device
  void copy_thermo_device_to_host() { copy_device_to_host(); }
  void copy_thermo_host_to_device() {}

  // ---- conserved totals (:739-885): one device reduction ------------------------------------------------------------------
  void conservation_entropy(bool first_iteration = false) {
    ccb_totals tt = totals();
    S = tt.S;
    if (first_iteration) S0 = S;
  }
  void conservation_BSQ(bool first_iteration = false) {
    ccb_totals tt = totals();
    Btotal = tt.B; Stotal = tt.Sq; Qtotal = tt.Q;
    if (first_iteration) { Btotal0 = Btotal; Stotal0 = Stotal; Qtotal0 = Qtotal; }
  }
  void conservation_energy(bool first_iteration = false, double /*t*/ = 1.0) {
    ccb_totals tt = totals();
    E = tt.E;
    if (first_iteration) E0 = E;
    Ez = tt.Ez;
    Etot = E + Ez;
    Eloss = ((E0 - Etot) / E0) * 100;
  }
  void compute_eccentricities() {}   // observables: out of scope
  void compute_e_2_P(double, int) {}
  void compute_e_2_X(double, int) {}

 private:
  std::shared_ptr<Settings> settingsPtr;
  ccb_totals totals() {
    ccb_totals tt;
    device->check(ccb_reduce(device->ctx, &tt));
    return tt;
  }
};

}  // namespace ccake
#endif
