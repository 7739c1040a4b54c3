// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Builds the UNMODIFIED reference templates (included from /root/reference/src where they lie) with the reference's
// own CPU path — std::vector<double> containers and ArrayHandlerIterable / ArrayHandlerIterableSparse
// (reference src/molpro/linalg/array/ArrayHandlerIterable.h:34-128, ArrayHandlerIterableSparse.h:20-84,
// array/util/gemm.h:257-279) — into oracle/_ref/libitsolv_ref.so, and exposes it through a flat C ABI so that
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference arm can call it.
// Nothing under iterative_solver_b200/ links or loads this library.
//
// Two things here are not reference code:
//   * the host dense algebra the solver templates call (eigenproblem, svd_system, ...): the reference implements it
//     over Eigen/LAPACKE, neither present; iterative_solver_b200/host/helper_lapack.cpp restates it over LAPACK;
//   * the synthetic banded operator (BandedProblemHost below), the CPU twin of the harness' CUDA operator.
#include <cstdio>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../iterative_solver_b200/harness/solve_driver.h"

#include <ExampleProblem.h> // reference examples/ExampleProblem.h
#include <molpro/linalg/array/ArrayHandlerIterable.h>
#include <molpro/linalg/array/ArrayHandlerIterableSparse.h>
#include <molpro/linalg/array/ArrayHandlerSparse.h>
#include <molpro/linalg/array/util/Distribution.h>
#include <molpro/linalg/itsolv/subspace/gram_schmidt.h>
#include <molpro/linalg/itsolv/subspace/util.h>
#include <molpro/linalg/itsolv/DSpaceResetter.h>

namespace {
using Vec = std::vector<double>;
using PMap = std::map<size_t, double>;
namespace la = molpro::linalg::array;
namespace its = molpro::linalg::itsolv;
using itsolv_b200::harness::now_seconds;
using itsolv_b200::harness::trace;

thread_local std::string g_error;

//! The reference handler plus a record of every number it returns to the solver.
class TracingIterable : public la::ArrayHandlerIterable<Vec, Vec> {
public:
  using Base = la::ArrayHandlerIterable<Vec, Vec>;
  double dot(const Vec& x, const Vec& y) override {
    double d = Base::dot(x, y);
    trace().record('d', 1, 1, &d);
    return d;
  }
  Matrix<double> gemm_inner(const CVecRef<Vec>& xx, const CVecRef<Vec>& yy) override {
    // gemm_inner_default with the non-virtual dot, so that only the Gram matrix is recorded
    auto mat = Matrix<double>({xx.size(), yy.size()});
    for (size_t i = 0; i < mat.rows(); ++i)
      for (size_t j = 0; j < mat.cols(); ++j)
        mat(i, j) = Base::dot(xx.at(i).get(), yy.at(j).get());
    trace().record('g', mat.rows(), mat.cols(), mat.data().data());
    return mat;
  }
};

class TracingIterableSparse : public la::ArrayHandlerIterableSparse<Vec, PMap> {
public:
  using Base = la::ArrayHandlerIterableSparse<Vec, PMap>;
  double dot(const Vec& x, const PMap& y) override {
    double d = Base::dot(x, y);
    trace().record('d', 1, 1, &d);
    return d;
  }
  Matrix<double> gemm_inner(const CVecRef<Vec>& xx, const CVecRef<PMap>& yy) override {
    auto mat = Matrix<double>({xx.size(), yy.size()});
    for (size_t i = 0; i < mat.rows(); ++i)
      for (size_t j = 0; j < mat.cols(); ++j)
        mat(i, j) = Base::dot(xx.at(i).get(), yy.at(j).get());
    trace().record('g', mat.rows(), mat.cols(), mat.data().data());
    return mat;
  }
};

inline double band_entry(int64_t i, int64_t j, double eps) {
  return i == j ? double(i + 1) : eps * double(1 + ((i + j) % 7));
}

/*!
 * CPU twin of the harness operator: A(i,i) = i+1, A(i,j) = eps*(1 + (i+j) mod 7) for 0 < |i-j| <= b.
 * Row sums run over ascending j with a separate multiply and add, the order the CUDA kernel uses.
 */
class BandedProblemHost : public its::Problem<Vec> {
public:
  BandedProblemHost(int64_t n, int b, double eps, int rhs_kind = ITSOLV_RHS_SCALED) : n(n), b(b), eps(eps), rhs_kind(rhs_kind) {}
  const int64_t n;
  const int b;
  const double eps;
  const int rhs_kind;
  mutable double seconds_action = 0, seconds_precond = 0;

  void apply(const Vec& v, Vec& a) const {
    for (int64_t i = 0; i < n; ++i) {
      double acc = 0;
      const int64_t lo = std::max<int64_t>(0, i - b), hi = std::min<int64_t>(n - 1, i + b);
      for (int64_t j = lo; j <= hi; ++j) {
        const double t = band_entry(i, j, eps) * v[j];
        acc = acc + t;
      }
      a[i] = acc;
    }
  }
  void action(const CVecRef<Vec>& parameters, const VecRef<Vec>& actions) const override {
    const double t0 = now_seconds();
    for (size_t k = 0; k < parameters.size(); ++k)
      apply(parameters[k].get(), actions[k].get());
    seconds_action += now_seconds() - t0;
  }
  bool diagonals(Vec& d) const override {
    for (int64_t i = 0; i < n; ++i)
      d[i] = double(i + 1);
    return true;
  }
  void precondition(const VecRef<Vec>& residual, const std::vector<double>& shift, const Vec& diagonals) const override {
    const double t0 = now_seconds();
    its::precondition_default(residual, shift, diagonals); // reference IterativeSolver.h:46-55
    seconds_precond += now_seconds() - t0;
  }
  //! residual of the quadratic form used for DIIS: r = A (v - 1), value = (v-1).r / 2 (cf. examples/ExampleProblem.h:24-34)
  double residual(const Vec& v, Vec& a) const override {
    const double t0 = now_seconds();
    Vec shifted(v.size());
    for (size_t i = 0; i < v.size(); ++i) // target t(i) of the DIIS cases: 1/(i+1), or 1 for the legacy inputs
      shifted[i] = v[i] - (rhs_kind == ITSOLV_RHS_LEGACY ? 1.0 : 1.0 / double(i + 1));
    apply(shifted, a);
    double value = 0;
    for (size_t i = 0; i < v.size(); ++i)
      value += 0.5 * a[i] * shifted[i];
    seconds_action += now_seconds() - t0;
    return value;
  }
  std::vector<double> pp_action_matrix(const std::vector<PMap>& pparams) const override {
    std::vector<double> result(pparams.size() * pparams.size(), 0);
    size_t ij = 0;
    for (const auto& pi : pparams)
      for (const auto& pj : pparams) {
        for (const auto& pie : pi)
          for (const auto& pje : pj)
            if (std::llabs(int64_t(pje.first) - int64_t(pie.first)) <= b)
              result[ij] += band_entry(pje.first, pie.first, eps) * pje.second * pie.second;
        ij++;
      }
    return result;
  }
  void p_action(const std::vector<std::vector<double>>& p_coefficients, const CVecRef<PMap>& pparams,
                const VecRef<Vec>& actions) const override {
    for (size_t k = 0; k < p_coefficients.size(); k++) {
      auto& a = actions[k].get();
      for (size_t pindex = 0; pindex < pparams.size(); pindex++)
        for (const auto& pie : pparams[pindex].get()) {
          const double coeff = pie.second * p_coefficients[k][pindex];
          const int64_t c = int64_t(pie.first);
          for (int64_t i = std::max<int64_t>(0, c - b); i <= std::min<int64_t>(n - 1, c + b); ++i) {
            const double t = band_entry(i, c, eps) * coeff;
            a[i] = a[i] + t;
          }
        }
    }
  }
  //! known solutions of the right-hand sides (include/itsolv_b200_harness.h, ITSOLV_RHS_*); the CUDA harness generates the same numbers
  static double rhs_u(int k, int64_t i) { return double((i * (k + 2) + k) % (2 * k + 5)) / double(2 * k + 5) - 0.5; }
  static double rhs_solution(int rhs_kind, int k, int64_t i) {
    const double u = rhs_u(k, i);
    return rhs_kind == ITSOLV_RHS_LEGACY ? u : (double(k + 1) + u) / double(i + 1);
  }
  void make_rhs(int k, Vec& out) const {
    Vec u(n);
    for (int64_t i = 0; i < n; ++i)
      u[i] = rhs_solution(rhs_kind, k, i);
    apply(u, out);
  }
};

//! The reference's own example problem (examples/ExampleProblem.h:6-44), plus what the driver template expects.
class ExampleProblemHost : public ExampleProblem {
public:
  explicit ExampleProblemHost(size_t n) : ExampleProblem(n) {}
  mutable double seconds_action = 0, seconds_precond = 0;
  void make_rhs(int k, Vec& out) const {
    Vec u(n);
    for (size_t i = 0; i < n; ++i)
      u[i] = BandedProblemHost::rhs_u(k, int64_t(i));
    ExampleProblem::action(its::cwrap_arg(u), its::wrap_arg(out));
  }
};

template <class ProblemT>
struct HostBackend {
  using R = Vec;
  ProblemT& prob;
  size_t n;
  std::shared_ptr<its::ArrayHandlers<Vec, Vec, PMap>> h;
  HostBackend(ProblemT& p, size_t n) : prob(p), n(n) {
    auto dense = std::make_shared<TracingIterable>();
    auto sparse = std::make_shared<TracingIterableSparse>();
    h = std::make_shared<its::ArrayHandlers<Vec, Vec, PMap>>(dense, dense, std::make_shared<la::ArrayHandlerSparse<PMap, PMap>>(),
                                                             dense, sparse, dense, sparse);
  }
  auto handlers() { return h; }
  Vec make_vector() { return Vec(n, 0.0); }
  Vec make_output_vector() { return Vec(n, 0.0); }
  void export_local(const Vec& v, double* out) { std::copy(v.begin(), v.end(), out); }
  double* solutions_target(double* given, size_t) { return given; }
  size_t n_local() { return n; }
  ProblemT& problem() { return prob; }
  void synchronize() {}
  void timer_start() {}
  double timer_stop_ms() { return 0; }
  auto make_davidson(const std::shared_ptr<its::ArrayHandlers<Vec, Vec, PMap>>& handlers, const itsolv_solve_spec&) {
    return std::make_unique<its::LinearEigensystemDavidson<Vec, Vec, PMap>>(handlers);
  }
  auto make_lineq(const std::shared_ptr<its::ArrayHandlers<Vec, Vec, PMap>>& handlers, const itsolv_solve_spec&) {
    return std::make_unique<its::LinearEquationsDavidson<Vec, Vec, PMap>>(handlers);
  }
  auto make_diis(const std::shared_ptr<its::ArrayHandlers<Vec, Vec, PMap>>& handlers, const itsolv_solve_spec&) {
    return std::make_unique<its::NonLinearEquationsDIIS<Vec, Vec, PMap>>(handlers);
  }
};

Vec to_vec(const double* p, size_t n) { return Vec(p, p + n); }
PMap to_map(const int64_t* idx, const double* val, int nnz) {
  PMap m;
  for (int i = 0; i < nnz; ++i)
    m[size_t(idx[i])] = val[i];
  return m;
}
} // namespace

template <class F>
static int ref_guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "oracle: %s\n", e.what());
    return 1;
  }
}

extern "C" {

const char* ref_last_error() { return g_error.c_str(); }

//! Reference solve() on std::vector containers; same spec/result structs as the product harness.
int ref_solve(const itsolv_solve_spec* spec, itsolv_solve_result* result, double* solutions) {
  try {
    if (spec->problem == ITSOLV_PROBLEM_EXAMPLE) {
      ExampleProblemHost problem(spec->n);
      HostBackend<ExampleProblemHost> backend(problem, spec->n);
      return itsolv_b200::harness::run_solve(*spec, backend, *result, solutions);
    }
    BandedProblemHost problem(spec->n, spec->half_bandwidth, spec->eps, spec->rhs_kind);
    HostBackend<BandedProblemHost> backend(problem, spec->n);
    return itsolv_b200::harness::run_solve(*spec, backend, *result, solutions);
  } catch (const std::exception& e) {
    g_error = e.what();
    return 1;
  }
}

size_t ref_trace_entries() { return trace().entries.size(); }
size_t ref_trace_values() { return trace().values.size(); }
void ref_trace_read(itsolv_trace_entry* entries, double* values) {
  std::copy(trace().entries.begin(), trace().entries.end(), entries);
  std::copy(trace().values.begin(), trace().values.end(), values);
}

void ref_banded_apply(int64_t n, int b, double eps, const double* v, double* a) {
  BandedProblemHost p(n, b, eps);
  Vec vv = to_vec(v, n), aa(n);
  p.apply(vv, aa);
  std::copy(aa.begin(), aa.end(), a);
}

/* ---- the handler contract, called on the reference's ArrayHandlerIterable (vectors are packed row after row) ---- */

double ref_handler_dot(size_t n, const double* x, const double* y) {
  la::ArrayHandlerIterable<Vec, Vec> h;
  return h.dot(to_vec(x, n), to_vec(y, n));
}
void ref_handler_axpy(size_t n, double alpha, const double* x, double* y) {
  la::ArrayHandlerIterable<Vec, Vec> h;
  Vec yy = to_vec(y, n);
  h.axpy(alpha, to_vec(x, n), yy);
  std::copy(yy.begin(), yy.end(), y);
}
void ref_handler_scal(size_t n, double alpha, double* x) {
  la::ArrayHandlerIterable<Vec, Vec> h;
  Vec xx = to_vec(x, n);
  h.scal(alpha, xx);
  std::copy(xx.begin(), xx.end(), x);
}
void ref_handler_gemm_inner(int k, int m, size_t n, const double* X, const double* Y, double* out) {
  la::ArrayHandlerIterable<Vec, Vec> h;
  std::vector<Vec> xs, ys;
  for (int i = 0; i < k; ++i)
    xs.push_back(to_vec(X + size_t(i) * n, n));
  for (int j = 0; j < m; ++j)
    ys.push_back(to_vec(Y + size_t(j) * n, n));
  auto mat = h.gemm_inner(its::cwrap(xs), its::cwrap(ys));
  std::copy(mat.data().begin(), mat.data().end(), out);
}
void ref_handler_gemm_outer(int k, int m, size_t n, const double* alpha, const double* X, double* Y) {
  la::ArrayHandlerIterable<Vec, Vec> h;
  std::vector<Vec> xs, ys;
  for (int i = 0; i < k; ++i)
    xs.push_back(to_vec(X + size_t(i) * n, n));
  for (int j = 0; j < m; ++j)
    ys.push_back(to_vec(Y + size_t(j) * n, n));
  Matrix<double> a(Vec(alpha, alpha + size_t(k) * m), {size_t(k), size_t(m)});
  h.gemm_outer(a, its::cwrap(xs), its::wrap(ys));
  for (int j = 0; j < m; ++j)
    std::copy(ys[j].begin(), ys[j].end(), Y + size_t(j) * n);
}
int ref_handler_select(size_t nsel, size_t n, const double* x, int max, int ignore_sign, int64_t* idx, double* val) {
  la::ArrayHandlerIterable<Vec, Vec> h;
  auto sel = h.select(nsel, to_vec(x, n), max != 0, ignore_sign != 0);
  int c = 0;
  for (const auto& s : sel) {
    idx[c] = int64_t(s.first);
    val[c] = s.second;
    ++c;
  }
  return c;
}
int ref_handler_select_max_dot(size_t nsel, size_t n, const double* x, const double* y, int64_t* idx, double* val) {
  la::ArrayHandlerIterable<Vec, Vec> h;
  auto sel = h.select_max_dot(nsel, to_vec(x, n), to_vec(y, n));
  int c = 0;
  for (const auto& s : sel) {
    idx[c] = int64_t(s.first);
    val[c] = s.second;
    ++c;
  }
  return c;
}
//! precondition_default for iterable containers, reference IterativeSolver.h:46-55
void ref_precondition_default(int w, size_t n, double* r, const double* shift, const double* diag) {
  std::vector<Vec> rs;
  for (int i = 0; i < w; ++i)
    rs.push_back(to_vec(r + size_t(i) * n, n));
  its::precondition_default(its::wrap(rs), std::vector<double>(shift, shift + w), to_vec(diag, n));
  for (int i = 0; i < w; ++i)
    std::copy(rs[i].begin(), rs[i].end(), r + size_t(i) * n);
}
//! subspace::util::modified_gram_schmidt, reference subspace/gram_schmidt.h:128-145
int ref_modified_gram_schmidt(int nvec, size_t n, double* data, double thresh, int* null_idx) {
  la::ArrayHandlerIterable<Vec, Vec> h;
  std::vector<Vec> ps;
  for (int i = 0; i < nvec; ++i)
    ps.push_back(to_vec(data + size_t(i) * n, n));
  auto w = its::wrap(ps);
  auto nulls = its::subspace::util::modified_gram_schmidt(w, h, thresh);
  for (int i = 0; i < nvec; ++i)
    std::copy(ps[i].begin(), ps[i].end(), data + size_t(i) * n);
  for (size_t i = 0; i < nulls.size(); ++i)
    null_idx[i] = int(nulls[i]);
  return int(nulls.size());
}

/* ---- dense x sparse (P-space) handler: ArrayHandlerIterableSparse<vector, map>; maps are CSR-packed ---- */

void ref_sparse_copy(size_t n, double* x, int nnz, const int64_t* idx, const double* val) {
  la::ArrayHandlerIterableSparse<Vec, PMap> h;
  Vec xx = to_vec(x, n);
  h.copy(xx, to_map(idx, val, nnz));
  std::copy(xx.begin(), xx.end(), x);
}
void ref_sparse_gemm_inner(int k, int m, size_t n, const double* X, const int* map_ptr, const int64_t* idx,
                           const double* val, double* out) {
  la::ArrayHandlerIterableSparse<Vec, PMap> h;
  std::vector<Vec> xs;
  std::vector<PMap> ps;
  for (int i = 0; i < k; ++i)
    xs.push_back(to_vec(X + size_t(i) * n, n));
  for (int j = 0; j < m; ++j)
    ps.push_back(to_map(idx + map_ptr[j], val + map_ptr[j], map_ptr[j + 1] - map_ptr[j]));
  auto mat = h.gemm_inner(its::cwrap(xs), its::cwrap(ps));
  std::copy(mat.data().begin(), mat.data().end(), out);
}
//! alphas is (number of maps) x (number of dense vectors): rows <-> xx (maps), columns <-> yy (dense)
void ref_sparse_gemm_outer(int nmap, int ndense, size_t n, const double* alpha, const int* map_ptr, const int64_t* idx,
                           const double* val, double* Y) {
  la::ArrayHandlerIterableSparse<Vec, PMap> h;
  std::vector<Vec> ys;
  std::vector<PMap> ps;
  for (int j = 0; j < ndense; ++j)
    ys.push_back(to_vec(Y + size_t(j) * n, n));
  for (int i = 0; i < nmap; ++i)
    ps.push_back(to_map(idx + map_ptr[i], val + map_ptr[i], map_ptr[i + 1] - map_ptr[i]));
  Matrix<double> a(Vec(alpha, alpha + size_t(nmap) * ndense), {size_t(nmap), size_t(ndense)});
  h.gemm_outer(a, its::cwrap(ps), its::wrap(ys));
  for (int j = 0; j < ndense; ++j)
    std::copy(ys[j].begin(), ys[j].end(), Y + size_t(j) * n);
}

//! util::make_distribution_spread_remainder, reference array/util/Distribution.h:99-110; borders has nproc+1 entries
// ---- the oracle's host algebra (oracle/helper_literal.cpp), for the fixtures of tests/test_host_algebra.py
int ref_host_eigenproblem(const double* matrix, const double* metric, size_t dimension, int hermitian,
                          double svd_threshold, double* eigenvalues, double* eigenvectors, size_t* nfound) {
  return ref_guarded([&] {
    std::vector<double> evec, eval;
    std::vector<double> m(matrix, matrix + dimension * dimension), s(metric, metric + dimension * dimension);
    its::eigenproblem(evec, eval, m, s, dimension, hermitian != 0, svd_threshold, 0, false);
    *nfound = eval.size();
    std::copy(eval.begin(), eval.end(), eigenvalues);
    std::copy(evec.begin(), evec.end(), eigenvectors);
  });
}

int ref_host_svd_system(size_t nrows, size_t ncols, const double* m, double threshold, int hermitian, int reduce_to_rank,
                      double* values, double* vectors, size_t* nfound) {
  return ref_guarded([&] {
    std::vector<double> copy(m, m + nrows * ncols);
    auto svds = its::svd_system(nrows, ncols, molpro::linalg::array::Span<double>(copy.data(), copy.size()), threshold,
                                hermitian != 0, reduce_to_rank != 0);
    size_t k = 0;
    for (const auto& s : svds) {
      values[k] = s.value;
      std::copy(s.v.begin(), s.v.end(), vectors + k * ncols);
      ++k;
    }
    *nfound = k;
  });
}

int ref_host_solve_linear_equations(const double* matrix, const double* metric, const double* rhs, size_t dimension,
                                  size_t nroot, double augmented_hessian, double svd_threshold, double* solution,
                                  double* eigenvalues) {
  return ref_guarded([&] {
    std::vector<double> sol, eval;
    its::solve_LinearEquations(sol, eval, std::vector<double>(matrix, matrix + dimension * dimension),
                               std::vector<double>(metric, metric + dimension * dimension),
                               std::vector<double>(rhs, rhs + dimension * nroot), dimension, nroot, augmented_hessian,
                               svd_threshold, 0);
    std::copy(sol.begin(), sol.end(), solution);
    if (eigenvalues)
      std::copy(eval.begin(), eval.end(), eigenvalues);
  });
}

int ref_host_solve_diis(const double* matrix, size_t dimension, double svd_threshold, double* solution) {
  return ref_guarded([&] {
    std::vector<double> sol;
    its::solve_DIIS(sol, std::vector<double>(matrix, matrix + dimension * dimension), dimension, svd_threshold, 0);
    std::copy(sol.begin(), sol.end(), solution);
  });
}

void ref_distribution(size_t n, int nproc, int64_t* borders) {
  auto d = la::util::make_distribution_spread_remainder<size_t>(n, nproc);
  for (int i = 0; i <= nproc; ++i)
    borders[i] = int64_t(d.chunk_borders()[i]);
}

/*
 * Per-op timing of the reference CPU handler (bench.py cpu_baseline): op 0 dot, 1 axpy, 2 scal, 3 gemm_inner[k x m],
 * 4 gemm_outer[k x m]. Vectors are allocated and filled outside the timed region. Returns seconds per call.
 */
double ref_time_handler_op(int op, size_t n, int k, int m, int reps) {
  la::ArrayHandlerIterable<Vec, Vec> h;
  std::vector<Vec> xs(std::max(k, 1), Vec(n)), ys(std::max(m, 1), Vec(n));
  for (size_t v = 0; v < xs.size(); ++v)
    for (size_t i = 0; i < n; ++i)
      xs[v][i] = 1.0 / double(1 + (i + v) % 97);
  for (size_t v = 0; v < ys.size(); ++v)
    for (size_t i = 0; i < n; ++i)
      ys[v][i] = 1.0 / double(1 + (i + 3 * v) % 89);
  Matrix<double> alpha({size_t(std::max(k, 1)), size_t(std::max(m, 1))});
  alpha.fill(1e-3);
  volatile double sink = 0;
  const double t0 = now_seconds();
  for (int r = 0; r < reps; ++r) {
    switch (op) {
    case 0:
      sink = sink + h.dot(xs[0], ys[0]);
      break;
    case 1:
      h.axpy(1e-3, xs[0], ys[0]);
      break;
    case 2:
      h.scal(1.0000001, ys[0]);
      break;
    case 3:
      sink = sink + h.gemm_inner(its::cwrap(xs), its::cwrap(ys))(0, 0);
      break;
    case 4:
      h.gemm_outer(alpha, its::cwrap(xs), its::wrap(ys));
      break;
    default:
      return -1;
    }
  }
  return (now_seconds() - t0) / reps;
}

/*
 * The protocol of the reference's own end-to-end eigensolver test on a dense matrix (reference
 * test/itsolv/test_LinearEigensystem.cpp: test_eigen :245-344, initialize_subspace :224-243, set_options :190-212, update
 * :96-104, initial_guess :146-157, initial_pspace / apply_p :159-188), with the reference's LinearEigensystemDavidson on
 * std::vector containers: unit-vector guess on the lowest diagonal elements (or a P space of `np` such unit vectors), then
 * action / add_vector / update / end_iteration until the working set is empty. hmat: n x n row-major.
 * Out: eigenvalues[nroot], errors[nroot], solutions[nroot][n], stats = {iterations, r_creations, loop count n_iter}.
 */
int ref_dense_eigen(size_t n, const double* hmat, int nroot, int np, int hermitian, int n_working_vectors_max,
                    double* eigenvalues, double* errors, double* solutions, int64_t* stats) {
  return ref_guarded([&] {
    auto H = [&](size_t i, size_t j) { return hmat[i * n + j]; };
    auto dense = std::make_shared<la::ArrayHandlerIterable<Vec, Vec>>();
    auto sparse = std::make_shared<la::ArrayHandlerIterableSparse<Vec, PMap>>();
    auto handlers = std::make_shared<its::ArrayHandlers<Vec, Vec, PMap>>(
        dense, dense, std::make_shared<la::ArrayHandlerSparse<PMap, PMap>>(), dense, sparse, dense, sparse);
    its::LinearEigensystemDavidson<Vec, Vec, PMap> solver(handlers);
    solver.set_n_roots(size_t(nroot));
    solver.set_convergence_threshold(1.0e-8);
    solver.set_max_size_qspace(std::max(6 * nroot, std::min(int(n), std::min(1000, 6 * nroot)) - np));
    solver.set_reset_D(8);
    solver.set_hermiticity(hermitian != 0);
    solver.set_verbosity(its::Verbosity::None);
    auto action = [&](const std::vector<Vec>& x, std::vector<Vec>& g) {
      for (size_t k = 0; k < x.size(); ++k)
        for (size_t i = 0; i < n; ++i) {
          double a = 0;
          for (size_t j = 0; j < n; ++j)
            a += H(i, j) * x[k][j];
          g[k][i] = a;
        }
    };
    auto update = [&](std::vector<Vec>& g, const std::vector<double>& shift) {
      for (size_t k = 0; k < g.size() && k < shift.size(); ++k)
        for (size_t i = 0; i < n; ++i)
          g[k][i] *= -1. / (1e-12 - shift[k] + H(i, i));
    };
    const size_t nroots = size_t(nroot);
    std::vector<Vec> x(nroots, Vec(n, 0.0)), g(nroots, Vec(n, 0.0));
    std::vector<double> diagonals(n);
    for (size_t i = 0; i < n; ++i)
      diagonals[i] = H(i, i);
    auto take_lowest = [&]() {
      const size_t at = size_t(std::min_element(diagonals.begin(), diagonals.end()) - diagonals.begin());
      diagonals[at] = 1e99;
      return at;
    };
    if (np > 0) {
      std::vector<PMap> pspace;
      for (int p = 0; p < np; ++p)
        pspace.push_back(PMap{{take_lowest(), 1.0}});
      std::vector<double> PP;
      for (const auto& i : pspace)
        for (const auto& j : pspace)
          PP.push_back(H(i.begin()->first, j.begin()->first));
      auto apply_p = [&](const std::vector<std::vector<double>>& pvectors, const its::CVecRef<PMap>& ps,
                         const its::VecRef<Vec>& act) {
        for (size_t i = 0; i < pvectors.size(); ++i)
          for (size_t pi = 0; pi < ps.size(); ++pi)
            for (const auto& pel : ps[pi].get())
              for (size_t j = 0; j < n; ++j)
                act[i].get()[j] += H(j, pel.first) * pel.second * pvectors[i][pi];
      };
      solver.add_p(its::cwrap(pspace), la::Span<double>(PP.data(), PP.size()), its::wrap(x), its::wrap(g), apply_p);
    } else {
      for (int root = 0; root < nroot; ++root)
        x[size_t(root)][take_lowest()] = 1;
      action(x, g);
      solver.add_vector(x, g);
    }
    const size_t nguess = size_t(std::max(nroot, n_working_vectors_max > 0 ? n_working_vectors_max : nroot));
    x.resize(nguess, Vec(n, 0.0));
    g.resize(nguess, Vec(n, 0.0));
    update(g, solver.working_set_eigenvalues());
    solver.end_iteration(x, g);
    int64_t n_iter = 2;
    for (int iter = 1; iter < 100; ++iter, ++n_iter) {
      action(x, g);
      if (solver.add_vector(x, g) == 0)
        break;
      update(g, solver.working_set_eigenvalues());
      if (solver.end_iteration(x, g) == 0)
        break;
    }
    const auto ev = solver.eigenvalues();
    const auto err = solver.errors();
    for (int i = 0; i < nroot; ++i) {
      eigenvalues[i] = ev.at(size_t(i));
      errors[i] = err.at(size_t(i));
    }
    std::vector<int> roots;
    for (int i = 0; i < nroot; ++i)
      roots.push_back(i);
    const size_t nr = size_t(nroot);
    std::vector<Vec> par(nr, Vec(n, 0.0)), res(nr, Vec(n, 0.0));
    solver.solution(roots, par, res);
    for (int i = 0; i < nroot; ++i)
      std::copy(par[size_t(i)].begin(), par[size_t(i)].end(), solutions + size_t(i) * n);
    stats[0] = int64_t(solver.statistics().iterations);
    stats[1] = int64_t(solver.statistics().r_creations);
    stats[2] = n_iter;
  });
}

/*
 * The reference's LinearEquations test (test/itsolv/test_LinearEquations.cpp:17-98, symmetric_system): a dense problem
 * with diagonals(), the library's own precondition_default for iterable containers (IterativeSolver.h:34-55), right-hand
 * sides added one by one, solve() with the default initial guess and threshold 1e-10, then solution() for all roots.
 * matrix: n x n row-major; rhs: nroot x n; out: solutions nroot x n, stats = {iterations, converged}.
 */
int ref_dense_lineq(size_t n, const double* matrix, int nroot, const double* rhs, double threshold, double* solutions,
                    int64_t* stats) {
  return ref_guarded([&] {
    struct DenseProblem : its::Problem<Vec> {
      size_t n;
      const double* m;
      void action(const its::CVecRef<Vec>& parameters, const its::VecRef<Vec>& act) const override {
        for (size_t v = 0; v < parameters.size(); ++v)
          for (size_t i = 0; i < n; ++i) {
            double a = 0;
            for (size_t j = 0; j < n; ++j)
              a += m[i * n + j] * parameters[v].get()[j];
            act[v].get()[i] = a;
          }
      }
      bool diagonals(Vec& d) const override {
        for (size_t i = 0; i < n; ++i)
          d[i] = m[i * n + i];
        return true;
      }
    } problem;
    problem.n = n;
    problem.m = matrix;
    auto dense = std::make_shared<la::ArrayHandlerIterable<Vec, Vec>>();
    auto sparse = std::make_shared<la::ArrayHandlerIterableSparse<Vec, PMap>>();
    auto handlers = std::make_shared<its::ArrayHandlers<Vec, Vec, PMap>>(
        dense, dense, std::make_shared<la::ArrayHandlerSparse<PMap, PMap>>(), dense, sparse, dense, sparse);
    its::LinearEquationsDavidson<Vec, Vec, PMap> solver(handlers);
    for (int root = 0; root < nroot; ++root)
      solver.add_equations(Vec(rhs + size_t(root) * n, rhs + size_t(root + 1) * n));
    const size_t nr = size_t(nroot);
    std::vector<Vec> parameters(nr, Vec(n, 0.0)), actions(nr, Vec(n, 0.0));
    solver.set_convergence_threshold(threshold);
    solver.set_verbosity(its::Verbosity::None);
    const bool ok = solver.solve(parameters, actions, problem, true);
    std::vector<int> roots;
    for (int i = 0; i < nroot; ++i)
      roots.push_back(i);
    solver.solution(roots, parameters, actions);
    for (size_t r = 0; r < nr; ++r)
      std::copy(parameters[r].begin(), parameters[r].end(), solutions + r * n);
    stats[0] = int64_t(solver.statistics().iterations);
    stats[1] = ok ? 1 : 0;
  });
}

/*
 * The reference's NonLinearEquations test (test/itsolv/test_NonLinearEquations.cpp:18-121, small_quadratic_form): DIIS
 * with convergence_threshold 1e-8 and max_size_qspace 6 on f = (x-1).h.(x-1)/2, h all ones with diagonal (i+2)*param;
 * start x = e_0; per iteration action, add_vector, update (division by the diagonal) when the working set is not
 * empty, end_iteration. Out: solution[n], residual[n], stats = {iterations, r_creations, n_iter}, error.
 */
int ref_dense_diis(size_t n, double param, double* solution, double* residual, int64_t* stats, double* error) {
  return ref_guarded([&] {
    auto H = [&](size_t i, size_t j) { return i == j ? double(i + 2) * param : 1.0; };
    auto dense = std::make_shared<la::ArrayHandlerIterable<Vec, Vec>>();
    auto sparse = std::make_shared<la::ArrayHandlerIterableSparse<Vec, PMap>>();
    auto handlers = std::make_shared<its::ArrayHandlers<Vec, Vec, PMap>>(
        dense, dense, std::make_shared<la::ArrayHandlerSparse<PMap, PMap>>(), dense, sparse, dense, sparse);
    its::NonLinearEquationsDIIS<Vec, Vec, PMap> solver(handlers);
    solver.set_convergence_threshold(1e-8);
    solver.set_max_size_qspace(6);
    solver.set_verbosity(its::Verbosity::None);
    Vec x(n, 0.0), g(n, 0.0);
    x.front() = 1;
    int nwork = 1;
    int64_t n_iter = 1;
    for (int iter = 1; iter < 1000 && nwork > 0; ++iter, ++n_iter) {
      for (size_t i = 0; i < n; ++i) {
        double a = 0;
        for (size_t j = 0; j < n; ++j)
          a += H(i, j) * (x[j] - 1.0);
        g[i] = a;
      }
      if (solver.add_vector(x, g, 0.0))
        for (size_t i = 0; i < n; ++i)
          g[i] = g[i] / H(i, i);
      nwork = int(solver.end_iteration(x, g));
    }
    Vec par(n, 0.0), res(n, 0.0);
    solver.solution(par, res);
    std::copy(par.begin(), par.end(), solution);
    std::copy(res.begin(), res.end(), residual);
    stats[0] = int64_t(solver.statistics().iterations);
    stats[1] = int64_t(solver.statistics().r_creations);
    stats[2] = n_iter;
    *error = solver.errors().front();
  });
}

/*
 * Small functions of the reference that the drivers (the reference's own and the fused ones) call unmodified, exposed
 * for the known-answer tests the reference holds for them (test/itsolv/subspace/test_util.cpp): Gram-Schmidt on an
 * overlap matrix (:109-152), eye_order (:76-107), overlap (:26-74), parameter_batches (:175-188).
 */
int ref_gram_schmidt(size_t n, const double* s, double* t, double* norms) {
  return ref_guarded([&] {
    its::subspace::Matrix<double> S(std::vector<double>(s, s + n * n), {n, n});
    its::subspace::Matrix<double> T;
    const auto result = its::subspace::util::gram_schmidt(S, T);
    if (result.size() != n || T.rows() != n || T.cols() != n)
      throw std::runtime_error("gram_schmidt: unexpected shape");
    std::copy(T.data().begin(), T.data().end(), t);
    std::copy(result.begin(), result.end(), norms);
  });
}

int ref_eye_order(size_t n, const double* m, int64_t* order) {
  return ref_guarded([&] {
    its::subspace::Matrix<double> M(std::vector<double>(m, m + n * n), {n, n});
    const auto o = its::subspace::util::eye_order(M);
    for (size_t i = 0; i < o.size(); ++i)
      order[i] = int64_t(o[i]);
  });
}

//! overlap of k vectors of length n with themselves: the one-set form (k(k+1)/2 dots) and the two-set form (gemm_inner)
int ref_overlap(int k, size_t n, const double* x, double* one_set, double* two_sets) {
  return ref_guarded([&] {
    std::vector<Vec> xs;
    for (int i = 0; i < k; ++i)
      xs.push_back(to_vec(x + size_t(i) * n, n));
    la::ArrayHandlerIterable<Vec, Vec> h;
    const auto a = its::subspace::util::overlap(its::cwrap(xs), h);
    const auto b = its::subspace::util::overlap(its::cwrap(xs), its::cwrap(xs), h);
    std::copy(a.data().begin(), a.data().end(), one_set);
    std::copy(b.data().begin(), b.data().end(), two_sets);
  });
}

//! pairs (begin, end) of detail::parameter_batches(nsol, nparam); returns the number of batches
int ref_parameter_batches(size_t nsol, size_t nparam, int64_t* pairs, int capacity) {
  const auto batches = its::detail::parameter_batches(nsol, nparam);
  for (size_t i = 0; i < batches.size() && int(i) < capacity; ++i) {
    pairs[2 * i] = int64_t(batches[i].first);
    pairs[2 * i + 1] = int64_t(batches[i].second);
  }
  return int(batches.size());
}

//! detail::max_overlap_with_R of the D-space resetter (reference itsolv/DSpaceResetter.h; KATs test/itsolv/
//! testDSpaceResetter.cpp:44-76): for nq Q vectors against nr R vectors of length n, the indices it returns
int ref_max_overlap_with_R(int nr, int nq, size_t n, const double* r, const double* q, int64_t* indices) {
  return ref_guarded([&] {
    std::vector<Vec> rs, qs;
    for (int i = 0; i < nr; ++i)
      rs.push_back(to_vec(r + size_t(i) * n, n));
    for (int i = 0; i < nq; ++i)
      qs.push_back(to_vec(q + size_t(i) * n, n));
    la::ArrayHandlerIterable<Vec, Vec> h;
    its::Logger logger;
    const auto out = its::detail::max_overlap_with_R(its::cwrap(rs), its::cwrap(qs), h, logger);
    for (size_t i = 0; i < out.size(); ++i)
      indices[i] = int64_t(out[i]);
  });
}

} // extern "C"

