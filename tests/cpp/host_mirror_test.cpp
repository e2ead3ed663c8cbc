// Drives the hot path through the C++ host mirror (include/ccake_b200_host.hpp) the way BSQHydro::run
// drives the reference classes (src/BSQHydro.cpp:745-814). Input: a file holding the ccb_config bytes and the
// IC arrays written by tests/test_cpp_host.py; output: the evolved state after `steps` RK2 steps.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ccake_b200_host.hpp"

using namespace ccake_b200;

static std::vector<double> read_vec(FILE* f, size_t n) {
  std::vector<double> v(n);
  if (n && fread(v.data(), sizeof(double), n, f) != n) { std::fprintf(stderr, "short read\n"); std::exit(2); }
  return v;
}

int main(int argc, char** argv) {
  if (argc != 3) { std::fprintf(stderr, "usage: %s <input.bin> <output.bin>\n", argv[0]); return 2; }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  ccb_config cfg;
  int64_t n = 0, steps = 0;
  double dt = 0;
  if (fread(&cfg, sizeof cfg, 1, f) != 1 || fread(&n, 8, 1, f) != 1 || fread(&steps, 8, 1, f) != 1 || fread(&dt, 8, 1, f) != 1) return 2;
  auto r = read_vec(f, 3 * n), e = read_vec(f, n), u = read_vec(f, 3 * n), shv6 = read_vec(f, 6 * n);
  std::fclose(f);
  try {
    auto ctx = std::make_shared<Context>(cfg);
    auto sys = std::make_shared<SystemState<2>>(ctx);
    SPHWorkstation<2> ws(ctx, sys);
    ccb_ic ic{};
    ic.n = n; ic.r = r.data(); ic.e_or_s = e.data(); ic.u = u.data(); ic.shv6 = shv6.data();
    sys->initialize(ic);
    for (int64_t k = 0; k < steps && ws.continue_evolution(1e9); ++k) ws.advance_timestep(dt, 2);
    ws.synchronize();
    std::vector<double> state;
    sys->copy_device_to_host(state);
    sys->conservation();
    FILE* o = std::fopen(argv[2], "wb");
    const double t = sys->t(), S = sys->S;
    const int64_t w = sys->state_width();
    fwrite(&t, 8, 1, o); fwrite(&S, 8, 1, o); fwrite(&w, 8, 1, o);
    fwrite(state.data(), sizeof(double), state.size(), o);
    std::fclose(o);
  } catch (const Error& err) {
    std::fprintf(stderr, "%s\n", err.what());
    return 1;
  }
  return 0;
}
