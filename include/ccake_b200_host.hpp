// ccake_b200_host.hpp -- C++ host side of the B200 hot path: the reference's class surface
// (SystemState<D>, SPHWorkstation<D>, Evolver<D>; reference include/system_state.h:40-127,
// include/sph_workstation.h:113-156, include/evolver.h:107-111) over the C ABI of ccake_b200.h.
//
// Header-only, C++17, no CUDA or Kokkos types: a maintainer includes this instead of the Kokkos/Cabana
// implementations of the three classes and links libccake_b200.so. Method names,
// argument meaning and error behaviour follow the reference; where the reference calls exit() the
// mirror throws ccake_b200::Error carrying the ccb_status and the library's message.
#ifndef CCAKE_B200_HOST_HPP
#define CCAKE_B200_HOST_HPP

#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "ccake_b200.h"

namespace ccake_b200 {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error("ccake_b200 error " + std::to_string(c) + ": " + what), code(c) {}
};

// One GPU context, shared by the three mirrors below (the reference shares a SystemState the same way).
class Context {
 public:
  explicit Context(const ccb_config& cfg) {
    const int rc = ccb_create(&ctx_, &cfg);
    if (rc) {
      const std::string msg = ctx_ ? ccb_last_error(ctx_) : "ccb_create failed";
      if (ctx_) ccb_destroy(ctx_);
      ctx_ = nullptr;
      throw Error(rc, msg);
    }
  }
  ~Context() { if (ctx_) ccb_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  ccb_ctx* get() const { return ctx_; }
  void check(int rc) const { if (rc) throw Error(rc, ccb_last_error(ctx_)); }

  // EquationOfState::init for eos_type "table": EoS_table(eos_path, normalize_by_T) (src/eos_initialization.cpp:273-279).
  // Call right after construction, before any particle is allocated.
  void load_eos_table(const std::string& eos_path, bool normalize_by_T = true) {
    check(ccb_load_eos_table_dat(ctx_, (eos_path + "/thermo.dat").c_str(), normalize_by_T ? 1 : 0));
  }
  void set_eos_table(const ccb_eos_table& table) { check(ccb_set_eos_table(ctx_, &table)); }

  // EoS_base::get_full_thermo / EquationOfState::update_s + set_thermo as batched device queries (include/ccake_b200.h).
  std::vector<double> eos_full_thermo(const std::vector<double>& points_TBQS, int member = 0) {
    const int64_t n = (int64_t)points_TBQS.size() / 4;
    std::vector<double> out((size_t)n * 17);
    check(ccb_eos_full_thermo(ctx_, member, n, points_TBQS.data(), out.data()));
    return out;
  }
  void eos_locate(const std::vector<double>& densities, const std::vector<double>& seed, std::vector<double>& thermo16, std::vector<int32_t>& eos,
                  bool energy = false) {
    const int64_t n = (int64_t)densities.size() / 4;
    thermo16.resize((size_t)n * 16); eos.resize((size_t)n);
    check(ccb_eos_locate(ctx_, energy ? 1 : 0, n, densities.data(), seed.empty() ? nullptr : seed.data(), nullptr, thermo16.data(), eos.data()));
  }

  // Slab decomposition (one Context per GPU / process; new: the reference is single-process). `unique_id` is the 128-byte
  // id rank 0 obtained from ccb_comm_unique_id and broadcast to the other ranks (MPI_Bcast, a file, a socket ...).
  void comm_init(const char unique_id[128]) { check(ccb_comm_init(ctx_, unique_id)); }
  void set_slab(int32_t col_lo, int32_t col_hi) { check(ccb_set_slab(ctx_, col_lo, col_hi)); }

 private:
  ccb_ctx* ctx_ = nullptr;
};

// SystemState<D>: particle storage and the host <-> device copies (reference src/system_state.cpp:58-555, 739-885).
template <unsigned int D>
class SystemState {
 public:
  explicit SystemState(std::shared_ptr<Context> c) : ctx(std::move(c)) {}

  double t() const { return ccb_time(ctx->get()); }
  void set_t(double v) { ctx->check(ccb_set_time(ctx->get(), v)); }
  int64_t n() const { return ccb_n(ctx->get()); }

  // BSQHydro::read_ccake rows + initialize_hydrodynamics (src/BSQHydro.cpp:207-425, 569-611)
  void initialize(const ccb_ic& ic) { ctx->check(ccb_upload_ic(ctx->get(), &ic)); }
  // allocate_cabana_particles for n blank particles (src/system_state.cpp:58-63)
  void allocate_cabana_particles(int64_t n_particles) { ctx->check(ccb_allocate(ctx->get(), n_particles)); }

  int state_width() const { return ccb_state_width(ctx->get()); }
  // evolved state as [column][particle id]; see ccb_upload_state in ccake_b200.h
  void copy_host_to_device(const std::vector<double>& state) {
    if ((int64_t)state.size() != (int64_t)state_width() * n()) throw Error(CCB_ERR_INVALID, "state block has the wrong size");
    ctx->check(ccb_upload_state(ctx->get(), state.data()));
  }
  void copy_device_to_host(std::vector<double>& state) {
    state.resize((size_t)state_width() * (size_t)n());
    ctx->check(ccb_download_state(ctx->get(), state.data()));
  }
  std::vector<double> get_field(ccb_field f) {
    std::vector<double> out((size_t)ccb_field_width(f) * (size_t)n());
    ctx->check(ccb_get_field(ctx->get(), f, out.data()));
    return out;
  }
  void set_field(ccb_field f, const std::vector<double>& v) { ctx->check(ccb_set_field(ctx->get(), f, v.data())); }

  void reset_neighbour_list() { ctx->check(ccb_build_cells(ctx->get())); }
  void initialize_linklist() { reset_neighbour_list(); }

  // conservation_entropy / conservation_BSQ / conservation_energy in one device pass; fills the public totals
  void conservation() {
    ccb_totals tot;
    ctx->check(ccb_reduce(ctx->get(), &tot));
    S = tot.S; Btotal = tot.B; Stotal = tot.Sq; Qtotal = tot.Q; E = tot.E; Ez = tot.Ez; number_part_fo = tot.n_frozen;
  }
  void conservation_entropy(bool /*first_iteration*/ = false) { conservation(); }
  void conservation_BSQ(bool /*first_iteration*/ = false) { conservation(); }
  void conservation_energy(bool /*first_iteration*/ = false, double /*t*/ = 0.0) { conservation(); }

  double S = 0, Btotal = 0, Stotal = 0, Qtotal = 0, E = 0, Ez = 0;
  int64_t number_part_fo = 0;
  std::shared_ptr<Context> ctx;
};

// Evolver<D>: the reference takes the derivative callback; the fused device step ignores it
// (the derivative evaluation and the RK update are one kernel sequence inside ccb_step).
template <unsigned int D>
class Evolver {
 public:
  explicit Evolver(std::shared_ptr<Context> c) : ctx(std::move(c)) {}
  void allocate_cache() {}
  void execute_timestep(double dt, int rk_order, std::function<void()> /*time_derivatives_functional*/ = nullptr) {
    if (rk_order != 2) throw Error(CCB_ERR_INVALID, "only rk_order 2 is supported");
    ctx->check(ccb_step(ctx->get(), dt));
  }
  std::shared_ptr<Context> ctx;
};

// SPHWorkstation<D>: the per-stage pipeline (reference src/sph_workstation.cpp:1510-1612, 2043-2073).
template <unsigned int D>
class SPHWorkstation {
 public:
  SPHWorkstation(std::shared_ptr<Context> c, std::shared_ptr<SystemState<D>> sys) : ctx(std::move(c)), systemPtr(std::move(sys)), evolver(ctx) {}

  void calculate_gamma_and_velocities() {}  // fused into smooth_all_particle_fields
  void smooth_all_particle_fields(double /*time_squared*/ = 0.0) { ctx->check(ccb_smooth_fields(ctx->get())); }
  void update_all_particle_thermodynamics() { ctx->check(ccb_update_thermodynamics(ctx->get())); }
  void reset_pi_tensor(double /*time_squared*/ = 0.0) { ctx->check(ccb_local_update(ctx->get())); }
  void update_all_particle_viscosities() {}  // fused into reset_pi_tensor
  void smooth_all_particle_gradients(double /*time_squared*/ = 0.0) { ctx->check(ccb_smooth_gradients(ctx->get())); }
  void evaluate_time_derivatives() { ctx->check(ccb_time_derivatives(ctx->get())); }
  void get_time_derivatives(double /*dt*/ = 0.0) { ctx->check(ccb_get_time_derivatives(ctx->get())); }
  void advance_timestep(double dt, int rk_order) { evolver.execute_timestep(dt, rk_order); }
  // waits for the stream; throws where the reference exits (negative entropy, no EoS solution)
  void synchronize() { ctx->check(ccb_synchronize(ctx->get())); }
  bool continue_evolution(double max_tau) {
    systemPtr->conservation();
    return systemPtr->t() < max_tau && systemPtr->number_part_fo < systemPtr->n();
  }

  std::shared_ptr<Context> ctx;
  std::shared_ptr<SystemState<D>> systemPtr;
  Evolver<D> evolver;
};

}  // namespace ccake_b200
#endif
