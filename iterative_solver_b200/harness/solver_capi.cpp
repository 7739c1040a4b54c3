// C ABI of the solve harness (include/itsolv_b200_harness.h): the reference's solver drivers
// (LinearEigensystemDavidson / LinearEquationsDavidson / NonLinearEquationsDIIS, compiled from the reference's own
// headers) run with DistrArrayCUDA containers and the CUDA handlers, on the synthetic banded operator whose action,
// diagonal, preconditioner and P-space action are CUDA kernels reached through include/itsolv_b200.h.
// There is no host arithmetic on vectors here and no fallback: a missing device or library is an error.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../harness/solve_driver.h"

#include <molpro/linalg/itsolv/helper.h>
#include <molpro/linalg/itsolv/subspace/gram_schmidt.h>

#include "ArrayHandlerCUDA.h"
#include "DistrArrayCUDA.h"
#include "FusedDavidson.h"
#include "FusedEquations.h"

namespace {
using itsolv_b200::check;
using itsolv_b200::DistrArrayCUDA;
using itsolv_b200::harness::now_seconds;
using itsolv_b200::harness::trace;
namespace its = molpro::linalg::itsolv;
using PMap = std::map<size_t, double>;
using R = DistrArrayCUDA;

thread_local std::string g_error;

inline double band_entry_host(int64_t i, int64_t j, double eps) {
  return i == j ? double(i + 1) : eps * double(1 + ((i + j) % 7));
}

/*!
 * The harness problem on the device. Operator kinds: generated banded, stored CSR (device arrays), or the reference's
 * dense ExampleProblem (single rank). Everything O(n) is a kernel launch on the context's stream.
 */
class DeviceProblem : public its::Problem<R>, public itsolv_b200::UsesDefaultDiagonalPreconditioner {
public:
  DeviceProblem(itsolv_ctx* ctx, const itsolv_solve_spec& spec)
      : ctx(ctx), n(spec.n), b(spec.half_bandwidth), eps(spec.eps), kind(spec.problem), rhs_kind(spec.rhs_kind),
        scratch(size_t(spec.n), ctx) {
    nranks = itsolv_comm_size(ctx);
    rank = itsolv_comm_rank(ctx);
    start = int64_t(scratch.local_start());
    nloc = scratch.local_size();
    if (kind == ITSOLV_PROBLEM_EXAMPLE && nranks != 1)
      throw std::invalid_argument("the dense ExampleProblem operator runs on one rank only");
    if (nranks > 1) {
      if (nloc < size_t(b))
        throw std::invalid_argument("shard shorter than the half bandwidth");
      check(itsolv_alloc(ctx, size_t(2 * std::max(b, 1)), &halo), "halo allocation");
    }
  }
  ~DeviceProblem() override {
    itsolv_free(ctx, halo);
    itsolv_free(ctx, d_val);
    itsolv_free(ctx, reinterpret_cast<double*>(d_row_ptr));
    itsolv_free(ctx, reinterpret_cast<double*>(d_col));
    itsolv_free(ctx, d_diag);
  }

  itsolv_ctx* ctx;
  const int64_t n;
  const int b;
  const double eps;
  const int kind;
  int rhs_kind; //!< ITSOLV_RHS_*: which known solutions x_k the right-hand sides b_k = A x_k are built from
  int nranks = 1, rank = 0;
  int64_t start = 0;
  size_t nloc = 0;
  mutable double seconds_action = 0, seconds_precond = 0;
  mutable R scratch;
  mutable R halos; // halo rows of a whole working set (multi-vector SpMV)
  double* halo = nullptr; // [b lower halo | b upper halo]
  // stored operator (explicit CSR)
  int64_t* d_row_ptr = nullptr;
  int32_t* d_col = nullptr;
  double* d_val = nullptr;
  double* d_diag = nullptr;

  //! generate the banded operator as CSR on the host (this rank's rows) and upload it
  void build_csr() {
    std::vector<int64_t> row_ptr(nloc + 1, 0);
    std::vector<int32_t> col;
    std::vector<double> val, diag(nloc);
    col.reserve(nloc * size_t(2 * b + 1));
    val.reserve(nloc * size_t(2 * b + 1));
    for (size_t r = 0; r < nloc; ++r) {
      const int64_t i = start + int64_t(r);
      for (int64_t j = std::max<int64_t>(0, i - b); j <= std::min<int64_t>(n - 1, i + b); ++j) {
        col.push_back(int32_t(j));
        val.push_back(band_entry_host(i, j, eps));
      }
      row_ptr[r + 1] = int64_t(col.size());
      diag[r] = double(i + 1);
    }
    upload_csr(row_ptr.data(), col.data(), val.data(), diag.data());
  }

  void upload_csr(const int64_t* row_ptr, const int32_t* col, const double* val, const double* diag) {
    static_assert(sizeof(int64_t) == sizeof(double), "row pointers travel through the double allocator");
    if (n > int64_t(INT32_MAX))
      throw std::invalid_argument("stored CSR uses 32-bit column indices");
    const size_t nnz = size_t(row_ptr[nloc]);
    double *p_rp = nullptr, *p_col = nullptr;
    check(itsolv_alloc(ctx, nloc + 1, &p_rp), "csr allocation");
    check(itsolv_alloc(ctx, (nnz + 1) / 2 + 1, &p_col), "csr allocation");
    check(itsolv_alloc(ctx, nnz, &d_val), "csr allocation");
    check(itsolv_alloc(ctx, nloc, &d_diag), "csr allocation");
    d_row_ptr = reinterpret_cast<int64_t*>(p_rp);
    d_col = reinterpret_cast<int32_t*>(p_col);
    check(itsolv_upload(ctx, p_rp, reinterpret_cast<const double*>(row_ptr), nloc + 1), "csr upload");
    check(itsolv_upload_bytes(ctx, p_col, col, nnz * sizeof(int32_t)), "csr upload");
    check(itsolv_upload(ctx, d_val, val, nnz), "csr upload");
    check(itsolv_upload(ctx, d_diag, diag, nloc), "csr upload");
  }

  void apply(const R& v, R& a) const {
    if (kind == ITSOLV_PROBLEM_EXAMPLE) {
      check(itsolv_example_apply_f64(ctx, nloc, v.data(), a.data()), "example apply");
      return;
    }
    const double *lo = nullptr, *hi = nullptr;
    if (nranks > 1 && b > 0) {
      // b boundary rows travel to the neighbouring shards (peer stores over NVLink, or ncclSend/ncclRecv)
      const double* xv = v.data();
      check(itsolv_comm_halo_exchange_multi(ctx, &xv, 1, nloc, b, halo), "halo exchange");
      lo = rank > 0 ? halo : nullptr;
      hi = rank < nranks - 1 ? halo + b : nullptr;
    }
    if (d_row_ptr)
      check(itsolv_csr_apply_f64(ctx, n, start, nloc, b, d_row_ptr, d_col, d_val, v.data(), lo, hi, a.data()), "csr apply");
    else
      check(itsolv_banded_apply_f64(ctx, n, start, nloc, b, eps, v.data(), lo, hi, a.data()), "banded apply");
  }

  void action(const CVecRef<R>& parameters, const VecRef<R>& actions) const override {
    const double t0 = now_seconds();
    const size_t w = parameters.size();
    if (d_row_ptr && w > 1 && kind == ITSOLV_PROBLEM_BANDED) {
      // stored CSR: all vectors of the working set in one pass over the matrix
      if (nranks > 1 && b > 0 && halos.size() < 2 * size_t(b) * w) {
        halos = R(2 * size_t(b) * w * size_t(nranks), ctx); // local length >= 2 b w
      }
      std::vector<const double*> x(w), lo(w, nullptr), hi(w, nullptr);
      std::vector<double*> y(w);
      for (size_t k = 0; k < w; ++k) {
        x[k] = parameters[k].get().data();
        y[k] = actions[k].get().data();
      }
      if (nranks > 1 && b > 0) {
        // boundary rows of the whole working set in one exchange
        check(itsolv_comm_halo_exchange_multi(ctx, x.data(), int(w), nloc, b, halos.data()), "halo exchange");
        for (size_t k = 0; k < w; ++k) {
          double* h = halos.data() + 2 * size_t(b) * k;
          lo[k] = rank > 0 ? h : nullptr;
          hi[k] = rank < nranks - 1 ? h + b : nullptr;
        }
      }
      check(itsolv_csr_apply_multi_f64(ctx, n, start, nloc, b, d_row_ptr, d_col, d_val, int(w), x.data(), lo.data(),
                                       hi.data(), y.data()),
            "csr apply");
    } else {
      for (size_t k = 0; k < w; ++k)
        apply(parameters[k].get(), actions[k].get());
    }
    seconds_action += now_seconds() - t0;
  }

  bool diagonals(R& d) const override {
    if (d_diag)
      check(itsolv_copy_f64(ctx, d.data(), d_diag, nloc), "diagonals");
    else
      check(itsolv_banded_fill_f64(ctx, 0, 0, start, nloc, d.data()), "diagonals");
    return true;
  }

  //! Davidson diagonal update, one kernel for all residuals (replaces precondition_default, reference IterativeSolver.h:46-55)
  void precondition(const VecRef<R>& residual, const std::vector<double>& shift, const R& diagonals) const override {
    const double t0 = now_seconds();
    std::vector<double*> r(residual.size());
    for (size_t k = 0; k < residual.size(); ++k)
      r[k] = residual[k].get().data();
    check(itsolv_precondition_f64(ctx, r.data(), int(r.size()), diagonals.data(), shift.data(), nloc), "precondition");
    seconds_precond += now_seconds() - t0;
  }

  //! r = A (v - t), value = (v-t).r / 2  (cf. reference examples/ExampleProblem.h:24-34, where t = 1); the banded operator
  //! takes t(i) = 1/(i+1) unless the legacy inputs are asked for (include/itsolv_b200.h, itsolv_banded_target_shift_f64)
  double residual(const R& v, R& a) const override {
    const double t0 = now_seconds();
    const int target = kind == ITSOLV_PROBLEM_EXAMPLE || rhs_kind == ITSOLV_RHS_LEGACY ? 1 : 0;
    check(itsolv_banded_target_shift_f64(ctx, target, start, v.data(), scratch.data(), nloc), "residual shift");
    apply(scratch, a);
    const double value = 0.5 * a.dot(scratch);
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
              result[ij] += band_entry_host(pje.first, pie.first, eps) * pje.second * pie.second;
        ij++;
      }
    return result;
  }

  void p_action(const std::vector<std::vector<double>>& p_coefficients, const CVecRef<PMap>& pparams,
                const VecRef<R>& actions) const override {
    if (pparams.empty() || p_coefficients.empty())
      return;
    std::vector<int32_t> ptr{0};
    std::vector<int64_t> idx;
    std::vector<double> val;
    for (const auto& p : pparams) {
      for (const auto& e : p.get()) {
        idx.push_back(int64_t(e.first));
        val.push_back(e.second);
      }
      ptr.push_back(int32_t(idx.size()));
    }
    const int nP = int(pparams.size()), nact = int(p_coefficients.size());
    std::vector<double> coef(size_t(nact) * nP);
    std::vector<double*> act(nact);
    for (int k = 0; k < nact; ++k) {
      for (int q = 0; q < nP; ++q)
        coef[size_t(k) * nP + q] = p_coefficients[k][q];
      act[k] = actions[k].get().data();
    }
    check(itsolv_banded_p_action_f64(ctx, n, start, nloc, b, eps, nact, act.data(), nP, ptr.data(), idx.data(),
                                     val.data(), coef.data()),
          "p_action");
  }

  void make_rhs(int k, R& out) const {
    rhs_solution(k, scratch);
    apply(scratch, out);
  }
  //! the known solution x_k of right-hand side k
  void rhs_solution(int k, R& out) const {
    check(itsolv_banded_fill_f64(ctx, rhs_kind == ITSOLV_RHS_LEGACY || kind == ITSOLV_PROBLEM_EXAMPLE ? 1 : 2, k, start, nloc, out.data()), "rhs_solution");
  }
};

struct DeviceBackend {
  using R = DistrArrayCUDA;
  itsolv_ctx* ctx;
  DeviceProblem& prob;
  size_t n;
  std::shared_ptr<itsolv_b200::HandlersCUDA> h;
  DeviceBackend(itsolv_ctx* ctx, DeviceProblem& p, size_t n) : ctx(ctx), prob(p), n(n) {
    h = itsolv_b200::make_handlers(
        [](char op, size_t rows, size_t cols, const double* values) { trace().record(op, rows, cols, values); });
  }
  auto handlers() { return h; }
  R make_vector() {
    R v(n, ctx);
    v.fill(0.0);
    return v;
  }
  R make_output_vector() { return R(n, ctx); }
  void export_local(const R& v, double* out) { v.download(out); }
  //! itsolv_harness_problem_solve_device: the solutions stay on the GPU, in memory taken from the context's pool only now,
  //! after the solver has finished (its high-water mark is over), and handed to the caller
  double** late_solutions = nullptr;
  double* solutions_target(double* given, size_t count) {
    if (given || !late_solutions)
      return given;
    check(itsolv_alloc(ctx, count, late_solutions), "solutions allocation");
    return *late_solutions;
  }
  size_t n_local() { return prob.nloc; }
  DeviceProblem& problem() { return prob; }
  void synchronize() { check(itsolv_ctx_synchronize(ctx), "synchronize"); }
  std::unique_ptr<its::LinearEigensystemDavidson<R, R, PMap>>
  make_davidson(const std::shared_ptr<itsolv_b200::HandlersCUDA>& handlers, const itsolv_solve_spec& spec) {
    if (spec.fused) { // 1: fused solve(); 2: the reference's solve() loop over the batched add_vector / end_iteration
      auto solver = std::make_unique<itsolv_b200::LinearEigensystemDavidsonFused>(handlers);
      solver->set_fuse_solve(spec.fused != 2);
      return solver;
    }
    return std::make_unique<its::LinearEigensystemDavidson<R, R, PMap>>(handlers);
  }
  std::unique_ptr<its::LinearEquationsDavidson<R, R, PMap>>
  make_lineq(const std::shared_ptr<itsolv_b200::HandlersCUDA>& handlers, const itsolv_solve_spec& spec) {
    if (spec.fused)
      return std::make_unique<itsolv_b200::LinearEquationsDavidsonFused>(handlers);
    return std::make_unique<its::LinearEquationsDavidson<R, R, PMap>>(handlers);
  }
  std::unique_ptr<its::NonLinearEquationsDIIS<R, R, PMap>>
  make_diis(const std::shared_ptr<itsolv_b200::HandlersCUDA>& handlers, const itsolv_solve_spec& spec) {
    if (spec.fused)
      return std::make_unique<itsolv_b200::NonLinearEquationsDIISFused>(handlers);
    return std::make_unique<its::NonLinearEquationsDIIS<R, R, PMap>>(handlers);
  }
  void timer_start() { check(itsolv_ctx_timer_start(ctx, 0), "timer"); }
  double timer_stop_ms() {
    double ms = 0;
    check(itsolv_ctx_timer_stop(ctx, 0, &ms), "timer");
    return ms;
  }
};

void fill_counters(itsolv_ctx* ctx, itsolv_solve_result* result) {
  itsolv_counters c;
  itsolv_ctx_counters(ctx, &c);
  result->n_dot = c.n_dot;
  result->n_axpy = c.n_axpy;
  result->n_scal = c.n_scal;
  result->n_copy = c.n_copy;
  result->n_fill = c.n_fill;
  result->n_gemm_inner = c.n_gemm_inner;
  result->n_gemm_outer = c.n_gemm_outer;
  result->handler_bytes = c.bytes;
  result->handler_device_seconds = c.device_seconds;
  result->kernel_launches = c.launches;
  result->bytes_gemm_inner = c.bytes_gemm_inner;
  result->seconds_gemm_inner = c.seconds_gemm_inner;
  result->bytes_gemm_outer = c.bytes_gemm_outer;
  result->seconds_gemm_outer = c.seconds_gemm_outer;
  result->bytes_blas1 = c.bytes_blas1;
  result->seconds_blas1 = c.seconds_blas1;
  result->bytes_residual = c.bytes_residual;
  result->seconds_residual = c.seconds_residual;
  result->calls_gemm_inner = c.calls_gemm_inner;
  result->calls_gemm_outer = c.calls_gemm_outer;
  result->calls_blas1 = c.calls_blas1;
  result->calls_residual = c.calls_residual;
}

template <class F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return 1;
  }
}

//! this rank's shard of row `v` of a packed host matrix of global-length rows
R shard_from_host(itsolv_ctx* ctx, size_t n, const double* host) {
  R x(n, ctx);
  x.upload(host + x.local_start());
  return x;
}
void shard_to_host(const R& x, double* host) { x.download(host + x.local_start()); }

} // namespace

extern "C" {

const char* itsolv_harness_last_error(void) { return g_error.c_str(); }

struct itsolv_harness_problem {
  itsolv_ctx* ctx;
  std::unique_ptr<DeviceProblem> problem;
};

int itsolv_harness_problem_create(itsolv_ctx* ctx, const itsolv_solve_spec* spec, const int64_t* row_ptr,
                                  const int32_t* col, const double* val, const double* diag,
                                  itsolv_harness_problem** out) {
  return guarded([&] {
    auto p = std::make_unique<itsolv_harness_problem>();
    p->ctx = ctx;
    p->problem = std::make_unique<DeviceProblem>(ctx, *spec);
    if (row_ptr)
      p->problem->upload_csr(row_ptr, col, val, diag);
    else if (spec->explicit_csr && spec->problem == ITSOLV_PROBLEM_BANDED)
      p->problem->build_csr();
    *out = p.release();
  });
}

int itsolv_harness_problem_solve(itsolv_harness_problem* p, const itsolv_solve_spec* spec, itsolv_solve_result* result,
                                 double* solutions) {
  return guarded([&] {
    if (spec->n != p->problem->n || spec->half_bandwidth != p->problem->b || spec->problem != p->problem->kind)
      throw std::invalid_argument("itsolv_harness_problem_solve: spec does not describe this operator");
    p->problem->rhs_kind = spec->rhs_kind;
    DeviceBackend backend(p->ctx, *p->problem, size_t(spec->n));
    itsolv_ctx_reset_counters(p->ctx);
    itsolv_b200::harness::run_solve(*spec, backend, *result, solutions);
    fill_counters(p->ctx, result);
  });
}

int itsolv_harness_problem_solve_device(itsolv_harness_problem* p, const itsolv_solve_spec* spec,
                                        itsolv_solve_result* result, double** device_solutions) {
  return guarded([&] {
    if (spec->n != p->problem->n || spec->half_bandwidth != p->problem->b || spec->problem != p->problem->kind)
      throw std::invalid_argument("itsolv_harness_problem_solve_device: spec does not describe this operator");
    if (!device_solutions)
      throw std::invalid_argument("itsolv_harness_problem_solve_device: null argument");
    *device_solutions = nullptr;
    p->problem->rhs_kind = spec->rhs_kind;
    DeviceBackend backend(p->ctx, *p->problem, size_t(spec->n));
    backend.late_solutions = device_solutions;
    itsolv_ctx_reset_counters(p->ctx);
    itsolv_b200::harness::run_solve(*spec, backend, *result, nullptr);
    fill_counters(p->ctx, result);
  });
}

void itsolv_harness_problem_destroy(itsolv_harness_problem* p) { delete p; }

int itsolv_harness_solve(itsolv_ctx* ctx, const itsolv_solve_spec* spec, itsolv_solve_result* result, double* solutions) {
  itsolv_harness_problem* p = nullptr;
  if (int rc = itsolv_harness_problem_create(ctx, spec, nullptr, nullptr, nullptr, nullptr, &p))
    return rc;
  const int rc = itsolv_harness_problem_solve(p, spec, result, solutions);
  itsolv_harness_problem_destroy(p);
  return rc;
}

int itsolv_harness_solve_host_csr(itsolv_ctx* ctx, const itsolv_solve_spec* spec, const int64_t* row_ptr,
                                  const int32_t* col, const double* val, const double* diag,
                                  itsolv_solve_result* result, double* solutions) {
  itsolv_harness_problem* p = nullptr;
  if (int rc = itsolv_harness_problem_create(ctx, spec, row_ptr, col, val, diag, &p))
    return rc;
  const int rc = itsolv_harness_problem_solve(p, spec, result, solutions);
  itsolv_harness_problem_destroy(p);
  return rc;
}

size_t itsolv_harness_trace_entries(void) { return trace().entries.size(); }
size_t itsolv_harness_trace_values(void) { return trace().values.size(); }
void itsolv_harness_trace_read(itsolv_trace_entry* entries, double* values) {
  std::copy(trace().entries.begin(), trace().entries.end(), entries);
  std::copy(trace().values.begin(), trace().values.end(), values);
}

/* ---- the handler contract through the plugin classes, host buffers in and out ---- */

int itsolv_handler_blas1(itsolv_ctx* ctx, int op, size_t n, double alpha, const double* x, double* y, double* result) {
  return guarded([&] {
    itsolv_b200::ArrayHandlerCUDA h;
    switch (op) {
    case 0: {
      R a = shard_from_host(ctx, n, x), c = shard_from_host(ctx, n, y);
      *result = h.dot(a, c);
      break;
    }
    case 1: {
      R a = shard_from_host(ctx, n, x), c = shard_from_host(ctx, n, y);
      h.axpy(alpha, a, c);
      shard_to_host(c, y);
      break;
    }
    case 2: {
      R c = shard_from_host(ctx, n, y);
      h.scal(alpha, c);
      shard_to_host(c, y);
      break;
    }
    case 3: {
      R c(n, ctx);
      h.fill(alpha, c);
      shard_to_host(c, y);
      break;
    }
    case 4: {
      R a = shard_from_host(ctx, n, x);
      R c = h.copy(a);
      R d(n, ctx);
      h.copy(d, c);
      shard_to_host(d, y);
      break;
    }
    default:
      throw std::invalid_argument("itsolv_handler_blas1: unknown op");
    }
  });
}

int itsolv_handler_gemm_inner(itsolv_ctx* ctx, int k, int m, size_t n, const double* X, const double* Y, int y_is_x,
                              double* out) {
  return guarded([&] {
    itsolv_b200::ArrayHandlerCUDA h;
    std::vector<R> xs, ys;
    for (int i = 0; i < k; ++i)
      xs.push_back(shard_from_host(ctx, n, X + size_t(i) * n));
    if (!y_is_x)
      for (int j = 0; j < m; ++j)
        ys.push_back(shard_from_host(ctx, n, Y + size_t(j) * n));
    auto mat = y_is_x ? h.gemm_inner(its::cwrap(xs), its::cwrap(xs)) : h.gemm_inner(its::cwrap(xs), its::cwrap(ys));
    std::copy(mat.data().begin(), mat.data().end(), out);
  });
}

int itsolv_handler_gemm_outer(itsolv_ctx* ctx, int k, int m, size_t n, const double* alpha, const double* X, double* Y) {
  return guarded([&] {
    itsolv_b200::ArrayHandlerCUDA h;
    std::vector<R> xs, ys;
    for (int i = 0; i < k; ++i)
      xs.push_back(shard_from_host(ctx, n, X + size_t(i) * n));
    for (int j = 0; j < m; ++j)
      ys.push_back(shard_from_host(ctx, n, Y + size_t(j) * n));
    Matrix<double> a(std::vector<double>(alpha, alpha + size_t(k) * m), {size_t(k), size_t(m)});
    h.gemm_outer(a, its::cwrap(xs), its::wrap(ys));
    for (int j = 0; j < m; ++j)
      shard_to_host(ys[j], Y + size_t(j) * n);
  });
}

int itsolv_handler_select(itsolv_ctx* ctx, size_t nsel, size_t n, const double* x, const double* y_or_null, int max,
                          int ignore_sign, int64_t* idx, double* val) {
  int count = 0;
  const int rc = guarded([&] {
    itsolv_b200::ArrayHandlerCUDA h;
    R a = shard_from_host(ctx, n, x);
    std::map<size_t, double> sel;
    if (y_or_null) {
      R c = shard_from_host(ctx, n, y_or_null);
      sel = h.select_max_dot(nsel, a, c);
    } else {
      sel = h.select(nsel, a, max != 0, ignore_sign != 0);
    }
    for (const auto& s : sel) {
      idx[count] = int64_t(s.first);
      val[count] = s.second;
      ++count;
    }
  });
  return rc ? -1 : count;
}

int itsolv_handler_precondition(itsolv_ctx* ctx, int w, size_t n, double* r, const double* shift, const double* diag) {
  return guarded([&] {
    itsolv_solve_spec spec{};
    spec.n = int64_t(n);
    spec.half_bandwidth = 0;
    spec.problem = ITSOLV_PROBLEM_BANDED;
    DeviceProblem problem(ctx, spec);
    std::vector<R> rs;
    for (int i = 0; i < w; ++i)
      rs.push_back(shard_from_host(ctx, n, r + size_t(i) * n));
    R d = shard_from_host(ctx, n, diag);
    problem.precondition(its::wrap(rs), std::vector<double>(shift, shift + w), d);
    for (int i = 0; i < w; ++i)
      shard_to_host(rs[i], r + size_t(i) * n);
  });
}

int itsolv_handler_modified_gram_schmidt(itsolv_ctx* ctx, int nvec, size_t n, double* data, double thresh, int* null_idx) {
  int count = 0;
  const int rc = guarded([&] {
    itsolv_b200::ArrayHandlerCUDA h;
    std::vector<R> ps;
    for (int i = 0; i < nvec; ++i)
      ps.push_back(shard_from_host(ctx, n, data + size_t(i) * n));
    auto w = its::wrap(ps);
    auto nulls = its::subspace::util::modified_gram_schmidt(w, h, thresh);
    for (int i = 0; i < nvec; ++i)
      shard_to_host(ps[i], data + size_t(i) * n);
    for (size_t i = 0; i < nulls.size(); ++i)
      null_idx[i] = int(nulls[i]);
    count = int(nulls.size());
  });
  return rc ? -1 : count;
}

static PMap to_map(const int64_t* idx, const double* val, int nnz);

int itsolv_handler_distr_array(itsolv_ctx* ctx, int op, size_t n, double scalar, int flags, const double* a,
                               const double* b, double* c, int nnz, const int64_t* idx, const double* val,
                               double* result, int64_t* sel_idx, double* sel_val) {
  int count = 0;
  const int rc = guarded([&] {
    R x = shard_from_host(ctx, n, c);
    switch (op) {
    case 0:
      x.add(shard_from_host(ctx, n, b));
      break;
    case 1:
      x.sub(shard_from_host(ctx, n, b));
      break;
    case 2:
      x.add(scalar);
      break;
    case 3:
      x.sub(scalar);
      break;
    case 4:
      x.recip();
      break;
    case 5:
      x.times(shard_from_host(ctx, n, a));
      break;
    case 6:
      x.times(shard_from_host(ctx, n, a), shard_from_host(ctx, n, b));
      break;
    case 7:
      x.divide(shard_from_host(ctx, n, a), shard_from_host(ctx, n, b), scalar, (flags & 1) != 0, (flags & 2) != 0);
      break;
    case 8:
      x.axpy(scalar, to_map(idx, val, nnz));
      break;
    case 9:
      *result = x.dot(to_map(idx, val, nnz));
      break;
    case 10: {
      itsolv_b200::ArrayHandlerCUDASparse h;
      const auto sel = h.select_max_dot(size_t(flags), x, to_map(idx, val, nnz));
      for (const auto& s : sel) {
        sel_idx[count] = int64_t(s.first);
        sel_val[count] = s.second;
        ++count;
      }
      break;
    }
    case 11:
      x.zero();
      break;
    default:
      throw std::invalid_argument("itsolv_handler_distr_array: unknown op");
    }
    shard_to_host(x, c);
  });
  return rc ? -1 : count;
}

static PMap to_map(const int64_t* idx, const double* val, int nnz) {
  PMap m;
  for (int i = 0; i < nnz; ++i)
    m[size_t(idx[i])] = val[i];
  return m;
}

int itsolv_handler_sparse_copy(itsolv_ctx* ctx, size_t n, double* x, int nnz, const int64_t* idx, const double* val) {
  return guarded([&] {
    itsolv_b200::ArrayHandlerCUDASparse h;
    R a = shard_from_host(ctx, n, x);
    h.copy(a, to_map(idx, val, nnz));
    shard_to_host(a, x);
  });
}

int itsolv_handler_sparse_gemm_inner(itsolv_ctx* ctx, int k, int m, size_t n, const double* X, const int32_t* map_ptr,
                                     const int64_t* idx, const double* val, double* out) {
  return guarded([&] {
    itsolv_b200::ArrayHandlerCUDASparse h;
    std::vector<R> xs;
    std::vector<PMap> ps;
    for (int i = 0; i < k; ++i)
      xs.push_back(shard_from_host(ctx, n, X + size_t(i) * n));
    for (int j = 0; j < m; ++j)
      ps.push_back(to_map(idx + map_ptr[j], val + map_ptr[j], map_ptr[j + 1] - map_ptr[j]));
    auto mat = h.gemm_inner(its::cwrap(xs), its::cwrap(ps));
    std::copy(mat.data().begin(), mat.data().end(), out);
  });
}

int itsolv_handler_sparse_gemm_outer(itsolv_ctx* ctx, int nmap, int ndense, size_t n, const double* alpha,
                                     const int32_t* map_ptr, const int64_t* idx, const double* val, double* Y) {
  return guarded([&] {
    itsolv_b200::ArrayHandlerCUDASparse h;
    std::vector<R> ys;
    std::vector<PMap> ps;
    for (int j = 0; j < ndense; ++j)
      ys.push_back(shard_from_host(ctx, n, Y + size_t(j) * n));
    for (int i = 0; i < nmap; ++i)
      ps.push_back(to_map(idx + map_ptr[i], val + map_ptr[i], map_ptr[i + 1] - map_ptr[i]));
    Matrix<double> a(std::vector<double>(alpha, alpha + size_t(nmap) * ndense), {size_t(nmap), size_t(ndense)});
    h.gemm_outer(a, its::cwrap(ps), its::wrap(ys));
    for (int j = 0; j < ndense; ++j)
      shard_to_host(ys[j], Y + size_t(j) * n);
  });
}

int itsolv_harness_banded_apply(itsolv_ctx* ctx, int64_t n, int b, double eps, int explicit_csr, const double* x, double* y) {
  return guarded([&] {
    itsolv_solve_spec spec{};
    spec.n = n;
    spec.half_bandwidth = b;
    spec.eps = eps;
    spec.problem = ITSOLV_PROBLEM_BANDED;
    DeviceProblem problem(ctx, spec);
    if (explicit_csr)
      problem.build_csr();
    R v = shard_from_host(ctx, size_t(n), x), a(size_t(n), ctx);
    problem.apply(v, a);
    shard_to_host(a, y);
  });
}

int itsolv_host_eigenproblem(const double* matrix, const double* metric, size_t dimension, int hermitian,
                             double svd_threshold, double* eigenvalues, double* eigenvectors, size_t* nfound) {
  return guarded([&] {
    std::vector<double> evec, eval;
    std::vector<double> m(matrix, matrix + dimension * dimension), s(metric, metric + dimension * dimension);
    its::eigenproblem(evec, eval, m, s, dimension, hermitian != 0, svd_threshold, 0, false);
    *nfound = eval.size();
    std::copy(eval.begin(), eval.end(), eigenvalues);
    std::copy(evec.begin(), evec.end(), eigenvectors);
  });
}

int itsolv_host_svd_system(size_t nrows, size_t ncols, const double* m, double threshold, int hermitian, int reduce_to_rank,
                      double* values, double* vectors, size_t* nfound) {
  return guarded([&] {
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

int itsolv_host_solve_linear_equations(const double* matrix, const double* metric, const double* rhs, size_t dimension,
                                  size_t nroot, double augmented_hessian, double svd_threshold, double* solution,
                                  double* eigenvalues) {
  return guarded([&] {
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

int itsolv_host_solve_diis(const double* matrix, size_t dimension, double svd_threshold, double* solution) {
  return guarded([&] {
    std::vector<double> sol;
    its::solve_DIIS(sol, std::vector<double>(matrix, matrix + dimension * dimension), dimension, svd_threshold, 0);
    std::copy(sol.begin(), sol.end(), solution);
  });
}

} // extern "C"
