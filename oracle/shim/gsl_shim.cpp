// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
// Implementation of the GSL stand-in declared in gsl/gsl_shim.h.
#include "gsl/gsl_shim.h"

#include <algorithm>
#include <cfloat>
#include <cstring>

// ----------------------------------------------------------------------------
// LU with partial pivoting (row swaps recorded in p), as gsl_linalg_LU_decomp.
int gsl_linalg_LU_decomp(gsl_matrix* A, gsl_permutation* p, int* signum) {
  const size_t n = A->size1;
  *signum = 1;
  gsl_permutation_init(p);
  for (size_t j = 0; j + 1 < n; ++j) {
    double amax = std::fabs(gsl_matrix_get(A, j, j));
    size_t ipiv = j;
    for (size_t i = j + 1; i < n; ++i) {
      const double a = std::fabs(gsl_matrix_get(A, i, j));
      if (a > amax) { amax = a; ipiv = i; }
    }
    if (ipiv != j) {
      for (size_t k = 0; k < n; ++k) std::swap(A->data[j * n + k], A->data[ipiv * n + k]);
      std::swap(p->data[j], p->data[ipiv]);
      *signum = -*signum;
    }
    const double ajj = gsl_matrix_get(A, j, j);
    if (ajj != 0.0) {
      for (size_t i = j + 1; i < n; ++i) {
        const double aij = gsl_matrix_get(A, i, j) / ajj;
        gsl_matrix_set(A, i, j, aij);
        for (size_t k = j + 1; k < n; ++k)
          gsl_matrix_set(A, i, k, gsl_matrix_get(A, i, k) - aij * gsl_matrix_get(A, j, k));
      }
    }
  }
  return GSL_SUCCESS;
}

int gsl_linalg_LU_invert(const gsl_matrix* LU, const gsl_permutation* p, gsl_matrix* inv) {
  const size_t n = LU->size1;
  for (size_t i = 0; i < n; ++i)
    if (gsl_matrix_get(LU, i, i) == 0.0) return GSL_EDOM;
  for (size_t c = 0; c < n; ++c) {
    std::vector<double> x(n);
    // b = P e_c
    for (size_t i = 0; i < n; ++i) x[i] = (p->data[i] == c) ? 1.0 : 0.0;
    for (size_t i = 1; i < n; ++i)
      for (size_t k = 0; k < i; ++k) x[i] -= gsl_matrix_get(LU, i, k) * x[k];
    for (size_t ii = n; ii-- > 0;) {
      for (size_t k = ii + 1; k < n; ++k) x[ii] -= gsl_matrix_get(LU, ii, k) * x[k];
      x[ii] /= gsl_matrix_get(LU, ii, ii);
    }
    for (size_t i = 0; i < n; ++i) gsl_matrix_set(inv, i, c, x[i]);
  }
  return GSL_SUCCESS;
}

int gsl_blas_dgemv(CBLAS_TRANSPOSE_t t, double alpha, const gsl_matrix* A, const gsl_vector* x, double beta, gsl_vector* y) {
  const size_t n = A->size1, m = A->size2;
  if (t == CblasNoTrans) {
    for (size_t i = 0; i < n; ++i) {
      double acc = 0.0;
      for (size_t j = 0; j < m; ++j) acc += gsl_matrix_get(A, i, j) * x->data[j];
      y->data[i] = alpha * acc + (beta == 0.0 ? 0.0 : beta * y->data[i]);
    }
  } else {
    for (size_t j = 0; j < m; ++j) {
      double acc = 0.0;
      for (size_t i = 0; i < n; ++i) acc += gsl_matrix_get(A, i, j) * x->data[i];
      y->data[j] = alpha * acc + (beta == 0.0 ? 0.0 : beta * y->data[j]);
    }
  }
  return GSL_SUCCESS;
}

int gsl_blas_ddot(const gsl_vector* x, const gsl_vector* y, double* r) {
  double acc = 0.0;
  for (size_t i = 0; i < x->size; ++i) acc += x->data[i] * y->data[i];
  *r = acc;
  return GSL_SUCCESS;
}

// ----------------------------------------------------------------------------
// Powell hybrid (HYBRD) with an explicit Jacobian (n is 4: no QR bookkeeping).
namespace {
const gsl_multiroot_fsolver_type hybrids_t{"hybrids", 1}, hybrid_t{"hybrid", 0}, dnewton_t{"dnewton", 0};

struct HybridState {
  size_t n;
  int scaled;
  std::vector<double> J, diag, xt, ft, df, newton, grad, tmp;
  size_t iter = 0, ncfail = 0, ncsuc = 0, nslow1 = 0, nslow2 = 0;
  double fnorm = 0, delta = 0;
};

double enorm(const std::vector<double>& v) { double s = 0; for (double x : v) s += x * x; return std::sqrt(s); }
double enorm(const double* v, size_t n) { double s = 0; for (size_t i = 0; i < n; ++i) s += v[i] * v[i]; return std::sqrt(s); }
double scaled_enorm(const std::vector<double>& d, const double* v, size_t n) {
  double s = 0; for (size_t i = 0; i < n; ++i) { const double u = d[i] * v[i]; s += u * u; } return std::sqrt(s);
}
bool all_finite(const double* v, size_t n) { for (size_t i = 0; i < n; ++i) if (!std::isfinite(v[i])) return false; return true; }

int eval(gsl_multiroot_function* F, const double* x, double* f, size_t n) {
  gsl_vector vx{n, const_cast<double*>(x)}, vf{n, f};
  return F->f(&vx, F->params, &vf);
}

int fdjacobian(gsl_multiroot_function* F, const double* x, const double* f, HybridState& st) {
  const size_t n = st.n;
  const double epsrel = std::sqrt(DBL_EPSILON);
  std::vector<double> x1(x, x + n), f1(n);
  int status = GSL_SUCCESS;
  for (size_t j = 0; j < n; ++j) {
    const double xj = x[j];
    double dx = epsrel * std::fabs(xj);
    if (dx == 0) dx = epsrel;
    x1[j] = xj + dx;
    if (eval(F, x1.data(), f1.data(), n) != GSL_SUCCESS) return GSL_EBADFUNC;
    x1[j] = xj;
    bool nonzero = false;
    for (size_t i = 0; i < n; ++i) {
      const double g = (f1[i] - f[i]) / dx;
      if (g != 0) nonzero = true;
      st.J[i * n + j] = g;
    }
    if (!nonzero) status = GSL_ESING;
  }
  return status;
}

void compute_diag(HybridState& st) {
  const size_t n = st.n;
  for (size_t j = 0; j < n; ++j) {
    double s = 0;
    for (size_t i = 0; i < n; ++i) s += st.J[i * n + j] * st.J[i * n + j];
    s = std::sqrt(s);
    st.diag[j] = (s == 0) ? 1.0 : s;
  }
}
void update_diag(HybridState& st) {
  const size_t n = st.n;
  for (size_t j = 0; j < n; ++j) {
    double s = 0;
    for (size_t i = 0; i < n; ++i) s += st.J[i * n + j] * st.J[i * n + j];
    s = std::sqrt(s);
    if (s == 0) s = 1.0;
    if (s > st.diag[j]) st.diag[j] = s;
  }
}
double compute_delta(const HybridState& st, const double* x) {
  const double Dx = scaled_enorm(st.diag, x, st.n), factor = 100.0;
  return (Dx > 0) ? factor * Dx : factor;
}

// newton = -J^{-1} f through Gaussian elimination with partial pivoting; a zero
// pivot is replaced by a tiny one (HYBRD does the same on R's diagonal).
void newton_direction(const HybridState& st, const double* f, std::vector<double>& out) {
  const size_t n = st.n;
  std::vector<double> A(st.J), b(f, f + n);
  double amaxall = 0;
  for (double a : A) amaxall = std::max(amaxall, std::fabs(a));
  for (size_t j = 0; j < n; ++j) {
    size_t piv = j; double amax = std::fabs(A[j * n + j]);
    for (size_t i = j + 1; i < n; ++i) if (std::fabs(A[i * n + j]) > amax) { amax = std::fabs(A[i * n + j]); piv = i; }
    if (piv != j) { for (size_t k = 0; k < n; ++k) std::swap(A[j * n + k], A[piv * n + k]); std::swap(b[j], b[piv]); }
    if (A[j * n + j] == 0) A[j * n + j] = DBL_EPSILON * (amaxall > 0 ? amaxall : 1.0);
    for (size_t i = j + 1; i < n; ++i) {
      const double m = A[i * n + j] / A[j * n + j];
      for (size_t k = j; k < n; ++k) A[i * n + k] -= m * A[j * n + k];
      b[i] -= m * b[j];
    }
  }
  for (size_t ii = n; ii-- > 0;) {
    for (size_t k = ii + 1; k < n; ++k) b[ii] -= A[ii * n + k] * b[k];
    b[ii] /= A[ii * n + ii];
  }
  for (size_t i = 0; i < n; ++i) out[i] = -b[i];
}

void dogleg(HybridState& st, const double* f, double* p) {
  const size_t n = st.n;
  newton_direction(st, f, st.newton);
  const double qnorm = scaled_enorm(st.diag, st.newton.data(), n);
  if (qnorm <= st.delta) { for (size_t i = 0; i < n; ++i) p[i] = st.newton[i]; return; }
  for (size_t j = 0; j < n; ++j) {
    double s = 0;
    for (size_t i = 0; i < n; ++i) s += st.J[i * n + j] * f[i];
    st.grad[j] = -s / st.diag[j];
  }
  const double gnorm = enorm(st.grad);
  if (gnorm == 0) { const double a = st.delta / qnorm; for (size_t i = 0; i < n; ++i) p[i] = a * st.newton[i]; return; }
  for (size_t j = 0; j < n; ++j) st.grad[j] = (st.grad[j] / gnorm) / st.diag[j];
  for (size_t i = 0; i < n; ++i) { double s = 0; for (size_t j = 0; j < n; ++j) s += st.J[i * n + j] * st.grad[j]; st.tmp[i] = s; }
  const double tnorm = enorm(st.tmp);
  const double sgnorm = (gnorm / tnorm) / tnorm;
  if (sgnorm > st.delta) { for (size_t i = 0; i < n; ++i) p[i] = st.delta * st.grad[i]; return; }
  const double bnorm = enorm(f, n);
  const double bg = bnorm / gnorm, bq = bnorm / qnorm, dq = st.delta / qnorm, dq2 = dq * dq;
  const double sd = sgnorm / st.delta, sd2 = sd * sd;
  const double t1 = bg * bq * sd, u = t1 - dq;
  const double t2 = t1 - dq * sd2 + std::sqrt(u * u + (1 - dq2) * (1 - sd2));
  const double alpha = dq * (1 - sd2) / t2, beta = (1 - alpha) * sgnorm;
  for (size_t i = 0; i < n; ++i) p[i] = alpha * st.newton[i] + beta * st.grad[i];
}
}  // namespace

const gsl_multiroot_fsolver_type* gsl_multiroot_fsolver_hybrids = &hybrids_t;
const gsl_multiroot_fsolver_type* gsl_multiroot_fsolver_hybrid = &hybrid_t;
const gsl_multiroot_fsolver_type* gsl_multiroot_fsolver_dnewton = &dnewton_t;

gsl_multiroot_fsolver* gsl_multiroot_fsolver_alloc(const gsl_multiroot_fsolver_type* T, size_t n) {
  auto* s = new gsl_multiroot_fsolver;
  s->type = T; s->function = nullptr;
  s->x = gsl_vector_alloc(n); s->f = gsl_vector_alloc(n); s->dx = gsl_vector_alloc(n);
  auto* st = new HybridState;
  st->n = n; st->scaled = T->scaled;
  st->J.assign(n * n, 0); st->diag.assign(n, 1); st->xt.assign(n, 0); st->ft.assign(n, 0); st->df.assign(n, 0);
  st->newton.assign(n, 0); st->grad.assign(n, 0); st->tmp.assign(n, 0);
  s->state = st;
  return s;
}
void gsl_multiroot_fsolver_free(gsl_multiroot_fsolver* s) {
  if (!s) return;
  gsl_vector_free(s->x); gsl_vector_free(s->f); gsl_vector_free(s->dx);
  delete static_cast<HybridState*>(s->state);
  delete s;
}

int gsl_multiroot_fsolver_set(gsl_multiroot_fsolver* s, gsl_multiroot_function* F, const gsl_vector* x) {
  auto& st = *static_cast<HybridState*>(s->state);
  const size_t n = st.n;
  s->function = F;
  std::memcpy(s->x->data, x->data, n * sizeof(double));
  for (size_t i = 0; i < n; ++i) s->dx->data[i] = 0;
  eval(F, s->x->data, s->f->data, n);
  fdjacobian(F, s->x->data, s->f->data, st);
  st.iter = 1; st.fnorm = enorm(s->f->data, n);
  st.ncfail = st.ncsuc = st.nslow1 = st.nslow2 = 0;
  if (st.scaled) compute_diag(st); else std::fill(st.diag.begin(), st.diag.end(), 1.0);
  st.delta = compute_delta(st, s->x->data);
  return GSL_SUCCESS;
}

int gsl_multiroot_fsolver_iterate(gsl_multiroot_fsolver* s) {
  auto& st = *static_cast<HybridState*>(s->state);
  const size_t n = st.n;
  gsl_multiroot_function* F = s->function;
  double* x = s->x->data; double* f = s->f->data; double* dx = s->dx->data;
  const double p1 = 0.1, p5 = 0.5, p001 = 0.001, p0001 = 0.0001;

  dogleg(st, f, dx);
  for (size_t i = 0; i < n; ++i) st.xt[i] = x[i] + dx[i];
  const double pnorm = scaled_enorm(st.diag, dx, n);
  if (st.iter == 1 && pnorm < st.delta) st.delta = pnorm;
  if (eval(F, st.xt.data(), st.ft.data(), n) != GSL_SUCCESS) return GSL_EBADFUNC;
  if (!all_finite(st.ft.data(), n)) return GSL_EBADFUNC;
  for (size_t i = 0; i < n; ++i) st.df[i] = st.ft[i] - f[i];
  const double fnorm1 = enorm(st.ft);
  const double actred = (fnorm1 < st.fnorm) ? 1 - (fnorm1 / st.fnorm) * (fnorm1 / st.fnorm) : -1.0;
  for (size_t i = 0; i < n; ++i) { double sacc = f[i]; for (size_t j = 0; j < n; ++j) sacc += st.J[i * n + j] * dx[j]; st.tmp[i] = sacc; }
  const double fnorm1p = enorm(st.tmp);
  const double prered = (fnorm1p < st.fnorm) ? 1 - (fnorm1p / st.fnorm) * (fnorm1p / st.fnorm) : 0.0;
  const double ratio = (prered > 0) ? actred / prered : 0.0;

  if (ratio < p1) { st.ncsuc = 0; st.ncfail++; st.delta *= p5; }
  else {
    st.ncfail = 0; st.ncsuc++;
    if (ratio >= p5 || st.ncsuc > 1) st.delta = std::max(st.delta, pnorm / p5);
    if (std::fabs(ratio - 1) <= p1) st.delta = pnorm / p5;
  }
  if (ratio >= p0001) {
    std::memcpy(x, st.xt.data(), n * sizeof(double));
    std::memcpy(f, st.ft.data(), n * sizeof(double));
    st.fnorm = fnorm1; st.iter++;
  }
  st.nslow1++;
  if (actred >= p001) st.nslow1 = 0;
  if (actred >= p1) st.nslow2 = 0;

  if (st.ncfail == 2) {
    fdjacobian(F, x, f, st);
    st.nslow2++;
    if (st.iter == 1) { if (st.scaled) compute_diag(st); st.delta = compute_delta(st, x); }
    else if (st.scaled) update_diag(st);
    return GSL_SUCCESS;
  }
  // Broyden rank-1 update: J += (df - J dx) (D^2 dx)^T / (dx^T D^2 dx)
  if (pnorm > 0) {
    for (size_t i = 0; i < n; ++i) {
      double jdx = 0; for (size_t j = 0; j < n; ++j) jdx += st.J[i * n + j] * dx[j];
      const double wi = (st.df[i] - jdx) / pnorm;
      for (size_t j = 0; j < n; ++j) st.J[i * n + j] += wi * (st.diag[j] * st.diag[j] * dx[j] / pnorm);
    }
  }
  if (st.nslow2 == 5) return GSL_ENOPROGJ;
  if (st.nslow1 == 10) return GSL_ENOPROG;
  return GSL_SUCCESS;
}

int gsl_multiroot_test_residual(const gsl_vector* f, double epsabs) {
  double r = 0;
  for (size_t i = 0; i < f->size; ++i) r += std::fabs(f->data[i]);
  return (r < epsabs) ? GSL_SUCCESS : GSL_CONTINUE;
}
int gsl_multiroot_test_delta(const gsl_vector* dx, const gsl_vector* x, double epsabs, double epsrel) {
  for (size_t i = 0; i < x->size; ++i) {
    const double tol = epsabs + epsrel * std::fabs(x->data[i]);
    if (!(std::fabs(dx->data[i]) < tol || tol == 0)) return GSL_CONTINUE;
  }
  return GSL_SUCCESS;
}
