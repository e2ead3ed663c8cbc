// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Stand-in for the handful of GSL entry points the CCAKE reference calls on the
// EoS path (src/rootfinder.cpp:206-359, src/eos_derivatives.cpp:457-480) so the
// reference translation units compile unchanged into oracle/_ref/. GSL itself
// (apt libgsl-dev, 2.7.x) is absent from this image.
//
//  * gsl_multiroot_fsolver_hybrids: restated as Powell's hybrid method (MINPACK
//    HYBRD as published: forward-difference Jacobian with eps = sqrt(DBL_EPSILON),
//    column scaling, dogleg step inside a trust region, Broyden rank-1 updates,
//    Jacobian refresh after two consecutive failures, the nslow1/nslow2
//    no-progress exits). Iterates are NOT bit-identical to GSL's QR-update
//    implementation; the accepted root satisfies the same stopping rule
//    sum|f_i| < 1e-10 that rootfinder.cpp applies.
//  * gsl_linalg_LU_decomp / LU_invert: Gaussian elimination with partial
//    pivoting (Doolittle, same pivot rule as GSL).
//  * gsl_eigen_*: declared only; check_causality is out of scope and aborts.
#ifndef CCAKE_ORACLE_GSL_SHIM_H
#define CCAKE_ORACLE_GSL_SHIM_H

#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <vector>

enum {
  GSL_SUCCESS = 0, GSL_FAILURE = -1, GSL_CONTINUE = -2, GSL_EDOM = 1, GSL_EINVAL = 4,
  GSL_EBADFUNC = 9, GSL_ESING = 21, GSL_ENOPROG = 27, GSL_ENOPROGJ = 28
};
enum CBLAS_TRANSPOSE_t { CblasNoTrans = 111, CblasTrans = 112 };

struct gsl_vector { size_t size; double* data; };
struct gsl_matrix { size_t size1, size2; double* data; };
struct gsl_permutation { size_t size; size_t* data; };
struct gsl_matrix_view { gsl_matrix matrix; };

inline gsl_vector* gsl_vector_alloc(size_t n) { auto* v = new gsl_vector{n, new double[n]()}; return v; }
inline void gsl_vector_free(gsl_vector* v) { if (v) { delete[] v->data; delete v; } }
inline double gsl_vector_get(const gsl_vector* v, size_t i) { return v->data[i]; }
inline void gsl_vector_set(gsl_vector* v, size_t i, double x) { v->data[i] = x; }
inline gsl_matrix* gsl_matrix_alloc(size_t n, size_t m) { return new gsl_matrix{n, m, new double[n * m]()}; }
inline void gsl_matrix_free(gsl_matrix* a) { if (a) { delete[] a->data; delete a; } }
inline double gsl_matrix_get(const gsl_matrix* a, size_t i, size_t j) { return a->data[i * a->size2 + j]; }
inline void gsl_matrix_set(gsl_matrix* a, size_t i, size_t j, double x) { a->data[i * a->size2 + j] = x; }
inline void gsl_matrix_set_zero(gsl_matrix* a) { for (size_t i = 0; i < a->size1 * a->size2; ++i) a->data[i] = 0.0; }
inline gsl_matrix_view gsl_matrix_view_array(double* base, size_t n1, size_t n2) { return gsl_matrix_view{gsl_matrix{n1, n2, base}}; }
inline gsl_permutation* gsl_permutation_alloc(size_t n) { return new gsl_permutation{n, new size_t[n]()}; }
inline void gsl_permutation_init(gsl_permutation* p) { for (size_t i = 0; i < p->size; ++i) p->data[i] = i; }
inline void gsl_permutation_free(gsl_permutation* p) { if (p) { delete[] p->data; delete p; } }

typedef void gsl_error_handler_t(const char*, const char*, int, int);
inline gsl_error_handler_t* gsl_set_error_handler_off() { return nullptr; }
inline gsl_error_handler_t* gsl_set_error_handler(gsl_error_handler_t*) { return nullptr; }

int gsl_linalg_LU_decomp(gsl_matrix* A, gsl_permutation* p, int* signum);
int gsl_linalg_LU_invert(const gsl_matrix* LU, const gsl_permutation* p, gsl_matrix* inverse);
int gsl_blas_dgemv(CBLAS_TRANSPOSE_t t, double alpha, const gsl_matrix* A, const gsl_vector* x, double beta, gsl_vector* y);
int gsl_blas_ddot(const gsl_vector* x, const gsl_vector* y, double* result);

struct gsl_multiroot_function {
  int (*f)(const gsl_vector* x, void* params, gsl_vector* f);
  size_t n;
  void* params;
};
struct gsl_multiroot_fsolver_type { const char* name; int scaled; };
struct gsl_multiroot_fsolver {
  const gsl_multiroot_fsolver_type* type;
  gsl_multiroot_function* function;
  gsl_vector* x;
  gsl_vector* f;
  gsl_vector* dx;
  void* state;
};
extern const gsl_multiroot_fsolver_type* gsl_multiroot_fsolver_hybrids;
extern const gsl_multiroot_fsolver_type* gsl_multiroot_fsolver_hybrid;
extern const gsl_multiroot_fsolver_type* gsl_multiroot_fsolver_dnewton;
gsl_multiroot_fsolver* gsl_multiroot_fsolver_alloc(const gsl_multiroot_fsolver_type* T, size_t n);
void gsl_multiroot_fsolver_free(gsl_multiroot_fsolver* s);
int gsl_multiroot_fsolver_set(gsl_multiroot_fsolver* s, gsl_multiroot_function* f, const gsl_vector* x);
int gsl_multiroot_fsolver_iterate(gsl_multiroot_fsolver* s);
int gsl_multiroot_test_residual(const gsl_vector* f, double epsabs);
int gsl_multiroot_test_delta(const gsl_vector* dx, const gsl_vector* x, double epsabs, double epsrel);

// ---- eigen (declarations only; out of scope) --------------------------------
struct gsl_complex { double dat[2]; };
#define GSL_REAL(z) ((z).dat[0])
#define GSL_IMAG(z) ((z).dat[1])
struct gsl_vector_complex { size_t size; double* data; };
struct gsl_matrix_complex { size_t size1, size2; double* data; };
struct gsl_vector_complex_view { gsl_vector_complex vector; };
struct gsl_eigen_nonsymmv_workspace { size_t n; };
enum gsl_eigen_sort_t { GSL_EIGEN_SORT_VAL_ASC, GSL_EIGEN_SORT_VAL_DESC, GSL_EIGEN_SORT_ABS_ASC, GSL_EIGEN_SORT_ABS_DESC };
[[noreturn]] inline void gsl_shim_eigen_abort() { std::fprintf(stderr, "gsl_eigen_* is not available in the oracle shim (check_causality is out of scope)\n"); std::abort(); }
inline gsl_vector_complex* gsl_vector_complex_alloc(size_t) { gsl_shim_eigen_abort(); }
inline void gsl_vector_complex_free(gsl_vector_complex*) {}
inline gsl_matrix_complex* gsl_matrix_complex_alloc(size_t, size_t) { gsl_shim_eigen_abort(); }
inline void gsl_matrix_complex_free(gsl_matrix_complex*) {}
inline gsl_eigen_nonsymmv_workspace* gsl_eigen_nonsymmv_alloc(size_t) { gsl_shim_eigen_abort(); }
inline void gsl_eigen_nonsymmv_free(gsl_eigen_nonsymmv_workspace*) {}
inline int gsl_eigen_nonsymmv(gsl_matrix*, gsl_vector_complex*, gsl_matrix_complex*, gsl_eigen_nonsymmv_workspace*) { gsl_shim_eigen_abort(); }
inline int gsl_eigen_nonsymmv_sort(gsl_vector_complex*, gsl_matrix_complex*, gsl_eigen_sort_t) { gsl_shim_eigen_abort(); }
inline gsl_complex gsl_vector_complex_get(const gsl_vector_complex*, size_t) { gsl_shim_eigen_abort(); }
inline gsl_vector_complex_view gsl_matrix_complex_column(gsl_matrix_complex*, size_t) { gsl_shim_eigen_abort(); }

#endif
