// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Minimal host-only stand-in for the slice of the Cabana API the CCAKE reference
// uses (AoSoA + slices, SimdPolicy loops, VerletList + neighbor_parallel_for),
// so that the reference sources compile unchanged into oracle/_ref/.
//
// Neighbour semantics follow upstream Cabana (LinkedCellList + VerletList with
// FullNeighborTag), which is NOT vendored in /root/reference:
//   * Cartesian grid: n_d = floor((max_d - min_d) / (ratio * r_cut)),
//     dx_d = (max_d - min_d) / n_d, cell(x) = floor((x - min_d) * (1 / dx_d)),
//     clamped into [0, n_d - 1];
//   * j is a neighbour of i  iff  i != j and |r_i - r_j|^2 <= r_cut^2.
// Because W(q >= 2) = 0 and grad W(q >= 2) = 0 (src/kernel.cpp:41-42, 103-108),
// the `<=` at the cut-off affects only the neighbour COUNT (btrack), never a
// smoothed field.
#ifndef CCAKE_ORACLE_SHIM_CABANA_CORE_HPP
#define CCAKE_ORACLE_SHIM_CABANA_CORE_HPP

#include <Kokkos_Core.hpp>

#include <cstring>
#include <tuple>

namespace Cabana {

template <class... Ts> struct MemberTypes {
  static constexpr std::size_t size = sizeof...(Ts);
  using tuple = std::tuple<Ts*...>;
};

namespace shim {
template <class T> struct member_traits {
  using value_type = T;
  static constexpr int rank = 0;
  static constexpr std::size_t e0 = 1, e1 = 1, e2 = 1;
};
template <class T, std::size_t A> struct member_traits<T[A]> {
  using value_type = T;
  static constexpr int rank = 1;
  static constexpr std::size_t e0 = A, e1 = 1, e2 = 1;
};
template <class T, std::size_t A, std::size_t B> struct member_traits<T[A][B]> {
  using value_type = T;
  static constexpr int rank = 2;
  static constexpr std::size_t e0 = A, e1 = B, e2 = 1;
};
template <class T, std::size_t A, std::size_t B, std::size_t C> struct member_traits<T[A][B][C]> {
  using value_type = T;
  static constexpr int rank = 3;
  static constexpr std::size_t e0 = A, e1 = B, e2 = C;
};
template <class T> struct identity { using type = T; };
}  // namespace shim

// One contiguous array per member, particle-major: data[i][e0][e1][e2].
template <class T, int VLEN> class Slice {
 public:
  using value_type = T;
  T* p = nullptr;
  std::size_t n = 0, s0 = 1, s1 = 1, s2 = 1, stride = 1;  // s*: extents
  Slice() {}
  Slice(T* p_, std::size_t n_, std::size_t e0, std::size_t e1, std::size_t e2)
      : p(p_), n(n_), s0(e0), s1(e1), s2(e2), stride(e0 * e1 * e2) {}
  std::size_t size() const { return n; }
  T& operator()(std::size_t i) const { return p[i * stride]; }
  T& operator()(std::size_t i, std::size_t a) const { return p[i * stride + a * s1 * s2]; }
  T& operator()(std::size_t i, std::size_t a, std::size_t b) const { return p[i * stride + (a * s1 + b) * s2]; }
  T& operator()(std::size_t i, std::size_t a, std::size_t b, std::size_t c) const { return p[i * stride + (a * s1 + b) * s2 + c]; }
  T& access(std::size_t s, std::size_t l) const { return (*this)(s * VLEN + l); }
  T& access(std::size_t s, std::size_t l, std::size_t a) const { return (*this)(s * VLEN + l, a); }
  T& access(std::size_t s, std::size_t l, std::size_t a, std::size_t b) const { return (*this)(s * VLEN + l, a, b); }
  T& access(std::size_t s, std::size_t l, std::size_t a, std::size_t b, std::size_t c) const { return (*this)(s * VLEN + l, a, b, c); }
};

template <class MT, class Dev, int VLEN> class AoSoA;
template <class... Ts, class Dev, int VLEN> class AoSoA<MemberTypes<Ts...>, Dev, VLEN> {
 public:
  using member_types = MemberTypes<Ts...>;
  static constexpr int vector_length = VLEN;
  std::size_t n = 0;
  // raw byte buffers, one per member, shared so that copies alias like Kokkos views do
  std::shared_ptr<std::array<std::vector<unsigned char>, sizeof...(Ts)>> buf;

  AoSoA() : buf(std::make_shared<std::array<std::vector<unsigned char>, sizeof...(Ts)>>()) {}
  AoSoA(const std::string&, std::size_t n_) : AoSoA() { resize(n_); }
  std::size_t size() const { return n; }
  void resize(std::size_t n_) {
    n = n_;
    std::size_t bytes[] = {sizeof(Ts)...};
    for (std::size_t m = 0; m < sizeof...(Ts); ++m) (*buf)[m].assign(bytes[m] * n, 0);
  }
};

template <std::size_t I, class... Ts, class Dev, int VLEN>
inline auto slice(const AoSoA<MemberTypes<Ts...>, Dev, VLEN>& a, const std::string& = "") {
  using M = std::tuple_element_t<I, std::tuple<shim::identity<Ts>...>>;
  using member = typename M::type;
  using tr = shim::member_traits<member>;
  using T = typename tr::value_type;
  return Slice<T, VLEN>(reinterpret_cast<T*>((*a.buf)[I].data()), a.n, tr::e0, tr::e1, tr::e2);
}

template <class A, class B> inline void deep_copy(A& dst, const B& src) {
  assert(dst.size() == src.size());
  for (std::size_t m = 0; m < src.buf->size(); ++m) {
    if ((*dst.buf)[m].data() != (*src.buf)[m].data())
      std::memcpy((*dst.buf)[m].data(), (*src.buf)[m].data(), (*src.buf)[m].size());
  }
}
template <class T, int V> inline void deep_copy(Slice<T, V>& dst, const Slice<T, V>& src) {
  if (dst.p != src.p) std::memcpy(dst.p, src.p, sizeof(T) * src.n * src.stride);
}
template <class S, class A> inline A create_mirror_view_and_copy(const S&, const A& a) { return a; }
template <class L, class A> inline void permute(const L&, A&) {}

template <int VLEN, class Exec> struct SimdPolicy {
  long b, e;
  SimdPolicy(long b_, long e_) : b(b_), e(e_) {}
};
template <int VLEN, class Exec, class F>
inline void simd_parallel_for(const SimdPolicy<VLEN, Exec>& p, const F& f, const std::string& = "") {
#pragma omp parallel for schedule(static)
  for (long i = p.b; i < p.e; ++i) f((int)(i / VLEN), (int)(i % VLEN));
}

struct FullNeighborTag {};
struct HalfNeighborTag {};
struct VerletLayoutCSR {};
struct VerletLayout2D {};
struct TeamOpTag {};
struct SerialOpTag {};
struct FirstNeighborsTag {};

// CSR full neighbour list built through a linked-cell binning (see file header).
template <class Mem, class Algo, class Layout, class Build> class VerletList {
 public:
  std::shared_ptr<std::vector<long>> offsets;
  std::shared_ptr<std::vector<int>> neighbors;
  std::shared_ptr<std::vector<int>> cell_of;  // linear cell id per particle (for parity tests)
  int ncell[3] = {0, 0, 0};

  VerletList() {}
  template <class PosSlice>
  VerletList(PosSlice x, std::size_t begin, std::size_t end, double radius, double cell_ratio,
             const double grid_min[3], const double grid_max[3]) {
    const std::size_t np = x.size();
    double rdx[3];
    for (int d = 0; d < 3; ++d) {
      const double delta = cell_ratio * radius;
      int nc = (int)std::floor((grid_max[d] - grid_min[d]) * (1.0 / delta));
      if (nc < 1) nc = 1;
      ncell[d] = nc;
      const double dx = (grid_max[d] - grid_min[d]) / nc;
      rdx[d] = 1.0 / dx;
    }
    auto locate = [&](double xp, int d) {
      int c = (int)std::floor((xp - grid_min[d]) * rdx[d]);
      if (c < 0) c = 0;
      if (c > ncell[d] - 1) c = ncell[d] - 1;
      return c;
    };
    const long ncells = (long)ncell[0] * ncell[1] * ncell[2];
    cell_of = std::make_shared<std::vector<int>>(np);
    std::vector<long> cstart(ncells + 1, 0);
    for (std::size_t i = 0; i < np; ++i) {
      const int ci = locate(x(i, 0), 0), cj = locate(x(i, 1), 1), ck = locate(x(i, 2), 2);
      const long c = ((long)ci * ncell[1] + cj) * ncell[2] + ck;
      (*cell_of)[i] = (int)c;
      cstart[c + 1]++;
    }
    for (long c = 0; c < ncells; ++c) cstart[c + 1] += cstart[c];
    std::vector<int> order(np);
    {
      std::vector<long> fill(cstart.begin(), cstart.end() - 1);
      for (std::size_t i = 0; i < np; ++i) order[fill[(*cell_of)[i]]++] = (int)i;
    }
    const double rsqr = radius * radius;
    offsets = std::make_shared<std::vector<long>>(np + 1, 0);
    auto visit = [&](std::size_t i, auto&& fn) {
      const int c = (*cell_of)[i];
      const int ck = c % ncell[2], cj = (c / ncell[2]) % ncell[1], ci = c / (ncell[2] * ncell[1]);
      const double xi = x(i, 0), yi = x(i, 1), zi = x(i, 2);
      for (int a = std::max(ci - 1, 0); a <= std::min(ci + 1, ncell[0] - 1); ++a)
        for (int b = std::max(cj - 1, 0); b <= std::min(cj + 1, ncell[1] - 1); ++b)
          for (int g = std::max(ck - 1, 0); g <= std::min(ck + 1, ncell[2] - 1); ++g) {
            const long cc = ((long)a * ncell[1] + b) * ncell[2] + g;
            for (long k = cstart[cc]; k < cstart[cc + 1]; ++k) {
              const int j = order[k];
              if ((std::size_t)j == i) continue;
              const double dx = xi - x(j, 0), dy = yi - x(j, 1), dz = zi - x(j, 2);
              if (dx * dx + dy * dy + dz * dz <= rsqr) fn(j);
            }
          }
    };
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = (long)begin; i < (long)end; ++i) {
      long cnt = 0;
      visit((std::size_t)i, [&](int) { ++cnt; });
      (*offsets)[i + 1] = cnt;
    }
    for (std::size_t i = 0; i < np; ++i) (*offsets)[i + 1] += (*offsets)[i];
    neighbors = std::make_shared<std::vector<int>>((*offsets)[np]);
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = (long)begin; i < (long)end; ++i) {
      long w = (*offsets)[i];
      visit((std::size_t)i, [&](int j) { (*neighbors)[w++] = j; });
    }
  }
};

template <class L> struct NeighborList {
  static std::size_t numNeighbor(const L& l, std::size_t i) { return (std::size_t)((*l.offsets)[i + 1] - (*l.offsets)[i]); }
  static std::size_t getNeighbor(const L& l, std::size_t i, std::size_t k) { return (std::size_t)(*l.neighbors)[(*l.offsets)[i] + k]; }
  static std::size_t totalNeighbor(const L& l) { return l.neighbors ? l.neighbors->size() : 0; }
};

template <class Policy, class F, class L>
inline void neighbor_parallel_for(const Policy& p, const F& f, const L& l, FirstNeighborsTag, TeamOpTag, const std::string& = "") {
#pragma omp parallel for schedule(dynamic, 64)
  for (long i = p.b; i < p.e; ++i) {
    const long b = (*l.offsets)[i], e = (*l.offsets)[i + 1];
    for (long k = b; k < e; ++k) f((int)i, (*l.neighbors)[k]);
  }
}
template <class Policy, class F, class L>
inline void neighbor_parallel_for(const Policy& p, const F& f, const L& l, FirstNeighborsTag, SerialOpTag, const std::string& s = "") {
  neighbor_parallel_for(p, f, l, FirstNeighborsTag(), TeamOpTag(), s);
}

}  // namespace Cabana

#endif
