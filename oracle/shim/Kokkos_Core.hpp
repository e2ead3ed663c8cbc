// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Minimal host-only stand-in for the slice of the Kokkos API that the CCAKE
// reference translation units use on the per-timestep SPH path (SURVEY.md
// section 8c). It exists so that the reference sources can be compiled
// *unchanged* from /root/reference into oracle/_ref/ and used as the parity
// oracle and CPU baseline. It carries no numerics of its own: loops run in index
// order (OpenMP-parallel over the outer index when built with -fopenmp),
// atomics are omp atomics, math functions forward to <cmath>.
#ifndef CCAKE_ORACLE_SHIM_KOKKOS_CORE_HPP
#define CCAKE_ORACLE_SHIM_KOKKOS_CORE_HPP

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <initializer_list>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

#define KOKKOS_FUNCTION
#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_FORCEINLINE_FUNCTION inline
#define KOKKOS_LAMBDA [=]
#define KOKKOS_CLASS_LAMBDA [=, *this]
#define KOKKOS_IF_ON_HOST(x) { KOKKOS_SHIM_STRIP x }
#define KOKKOS_IF_ON_DEVICE(x) {}
#define KOKKOS_SHIM_STRIP(...) __VA_ARGS__

namespace Kokkos {

struct HostSpace { HostSpace() {} };
struct CudaUVMSpace {};
struct Serial {};
struct OpenMP {};
struct Cuda {};
template <class E, class M> struct Device { using execution_space = E; using memory_space = M; };
struct RandomAccess {};
template <class T> struct MemoryTraits {};

inline void initialize(int&, char**) {}
inline void initialize() {}
inline void finalize() {}
inline void fence() {}
inline void fence(const std::string&) {}
[[noreturn]] inline void abort(const char* msg) { std::fprintf(stderr, "Kokkos::abort: %s\n", msg); std::abort(); }

// math
using std::sqrt; using std::pow; using std::fabs; using std::abs; using std::exp;
using std::log; using std::log10; using std::floor; using std::tanh; using std::isfinite;
inline float sqrtf(float x) { return std::sqrt(x); }
template <class T> inline T max(const T& a, const T& b) { return a < b ? b : a; }
template <class T> inline T min(const T& a, const T& b) { return b < a ? b : a; }

template <class T> inline void atomic_add(T* p, T v) {
#pragma omp atomic
  *p += v;
}
template <class T> inline void atomic_store(T* p, T v) {
#pragma omp atomic write
  *p = v;
}

template <class... Args> struct RangePolicy {
  long b, e;
  RangePolicy(long b_, long e_) : b(b_), e(e_) {}
};

template <class F> inline void parallel_for(const std::string&, long n, const F& f) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) f((int)i);
}
template <class... A, class F> inline void parallel_for(const std::string&, const RangePolicy<A...>& p, const F& f) {
#pragma omp parallel for schedule(static)
  for (long i = p.b; i < p.e; ++i) f((int)i);
}
template <class... A, class F> inline void parallel_for(const RangePolicy<A...>& p, const F& f) {
#pragma omp parallel for schedule(static)
  for (long i = p.b; i < p.e; ++i) f((int)i);
}

template <class T> struct Sum {
  T& ref;
  explicit Sum(T& r) : ref(r) {}
};

// Sum reductions. Serial in index order so that the oracle's totals are
// reproducible run to run.
template <class F, class T> inline void parallel_reduce(const std::string&, long n, const F& f, T& out) {
  T acc = T(0);
  for (long i = 0; i < n; ++i) f((int)i, acc);
  out = acc;
}
template <class F, class T> inline void parallel_reduce(const std::string&, long n, const F& f, Sum<T> out) {
  T acc = T(0);
  for (long i = 0; i < n; ++i) f((int)i, acc);
  out.ref = acc;
}
template <class... A, class F, class T> inline void parallel_reduce(const std::string&, const RangePolicy<A...>& p, const F& f, T& out) {
  T acc = T(0);
  for (long i = p.b; i < p.e; ++i) f((int)i, acc);
  out = acc;
}

namespace Profiling {
inline void pushRegion(const std::string&) {}
inline void popRegion() {}
}

// Rank-0 / rank-1 View, enough for the (out-of-scope) eccentricity code to compile.
template <class T, class... P> class View {
  std::shared_ptr<T> d;
 public:
  View() : d(std::make_shared<T>()) {}
  explicit View(const std::string&) : d(std::make_shared<T>()) { *d = T(); }
  T& operator()() const { return *d; }
};
template <class T, class... P> class View<T*, P...> {
  std::shared_ptr<std::vector<T>> d;
 public:
  View() : d(std::make_shared<std::vector<T>>()) {}
  View(const std::string&, size_t n) : d(std::make_shared<std::vector<T>>(n)) {}
  T& operator()(size_t i) const { return (*d)[i]; }
  size_t extent(int) const { return d->size(); }
  T* data() const { return d->data(); }
};
// rank-4 View (EoS_Interpolator's tables, include/eos_interpolator.h:32-34): LayoutRight, shared storage
template <class T, class... P> class View<T****, P...> {
  std::shared_ptr<std::vector<T>> d;
  size_t n[4] = {0, 0, 0, 0};
 public:
  View() : d(std::make_shared<std::vector<T>>()) {}
  View(const std::string&, size_t n0, size_t n1, size_t n2, size_t n3) : d(std::make_shared<std::vector<T>>(n0 * n1 * n2 * n3)) {
    n[0] = n0; n[1] = n1; n[2] = n2; n[3] = n3;
  }
  T& operator()(size_t i, size_t j, size_t k, size_t l) const { return (*d)[((i * n[1] + j) * n[2] + k) * n[3] + l]; }
  size_t extent(int a) const { return n[a]; }
  T* data() const { return d->data(); }
};
template <class V> inline V create_mirror_view(const V& v) { return v; }
template <class S, class V> inline V create_mirror_view_and_copy(const S&, const V& v) { return v; }
template <class T, class... P> inline void deep_copy(const View<T*, P...>& v, const T& x) {
  for (size_t i = 0; i < v.extent(0); ++i) v(i) = x;
}
template <class V> inline void deep_copy(V& a, const V& b) { a = b; }

}  // namespace Kokkos

#endif
