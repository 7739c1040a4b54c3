// Flat C interface of the solvers over device buffers (include/itsolv_b200_solver.h): the call sequence of the
// reference's IterativeSolverC.h (implementation: reference src/molpro/linalg/IterativeSolverCMPI.cpp:56-520) with
// DistrArrayCUDA views of the caller's device memory in the place of DistrArraySpan views of host memory.
#include <cctype>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <stack>
#include <stdexcept>
#include <string>
#include <vector>

#include <itsolv_b200_solver.h>
#include <molpro/linalg/itsolv/LinearEigensystemDavidson.h>
// LinearEquationsDavidson.h is not self-contained; it needs the includes of LinearEigensystemDavidson.h first.
#include <molpro/linalg/itsolv/LinearEquationsDavidson.h>
#include <molpro/linalg/itsolv/NonLinearEquationsDIIS.h>

#include "ArrayHandlerCUDA.h"
#include "DistrArrayCUDA.h"
#include "FusedDavidson.h"
#include "FusedEquations.h"

namespace {
namespace its = molpro::linalg::itsolv;
using itsolv_b200::DistrArrayCUDA;
using R = DistrArrayCUDA;
using P = std::map<size_t, double>;
using Solver = its::IterativeSolver<R, R, P>;
typedef void (*ApplyOnP)(const double*, double*, const size_t, const size_t*);

thread_local std::string g_error;

struct Instance {
  std::unique_ptr<Solver> solver;
  std::shared_ptr<itsolv_b200::HandlersCUDA> handlers;
  itsolv_ctx* ctx = nullptr;
  size_t dimension = 0;
  size_t n_local = 0;
  std::unique_ptr<R> diagonals;
  ApplyOnP apply_on_p = nullptr;
  bool has_eigenvalues = false;
};
// Only the top solver is active at any one time, as in the reference (IterativeSolverCMPI.cpp:59).
std::stack<Instance> instances;

Instance& top() {
  if (instances.empty())
    throw std::runtime_error("IterativeSolver not initialised properly");
  return instances.top();
}

//! views of `nvec` vectors stored one after the other with a stride of the shard length (IterativeSolverCMPI.cpp:89-107)
std::vector<R> views(Instance& in, size_t nvec, double* data) {
  if (nvec > 0 && !data)
    throw std::invalid_argument("null device buffer");
  std::vector<R> v;
  v.reserve(nvec);
  for (size_t i = 0; i < nvec; ++i)
    v.emplace_back(R::view(in.dimension, in.ctx, data + i * in.n_local));
  return v;
}

std::map<std::string, double> parse_options(const char* options, std::initializer_list<const char*> known) {
  std::map<std::string, double> out;
  std::string text = options ? options : "";
  for (auto& c : text)
    if (c == ',' || c == ';')
      c = ' ';
  std::istringstream is(text);
  std::string item;
  while (is >> item) {
    const auto eq = item.find('=');
    if (eq == std::string::npos)
      throw std::invalid_argument("option without a value: " + item);
    std::string key = item.substr(0, eq);
    for (auto& c : key)
      c = char(std::tolower(static_cast<unsigned char>(c)));
    bool ok = false;
    for (auto k : known)
      ok = ok || key == k;
    if (!ok)
      throw std::invalid_argument("unknown option: " + key);
    out[key] = std::stod(item.substr(eq + 1));
  }
  return out;
}

template <class S>
void apply_dspace_options(S& solver, const std::map<std::string, double>& opt) {
  if (opt.count("max_size_qspace"))
    solver.set_max_size_qspace(int(opt.at("max_size_qspace")));
  if (opt.count("reset_d"))
    solver.set_reset_D(size_t(opt.at("reset_d")));
  if (opt.count("reset_d_max_q_size"))
    solver.set_reset_D_maxQ_size(size_t(opt.at("reset_d_max_q_size")));
}

void set_logger(its::Logger& logger, int verbosity) {
  // the levels of the reference's interface (IterativeSolverCMPI.cpp:174-179)
  logger.max_trace_level = verbosity > 3 ? its::Logger::Info : (verbosity > 2 ? its::Logger::Trace : its::Logger::None);
  logger.max_warn_level = verbosity > 1 ? its::Logger::Warn : its::Logger::Error;
  logger.data_dump = verbosity > 0;
}

Instance& push_instance(itsolv_ctx* ctx, size_t n, std::unique_ptr<Solver> solver,
                        std::shared_ptr<itsolv_b200::HandlersCUDA> handlers, size_t* range_begin, size_t* range_end) {
  if (!ctx)
    throw std::invalid_argument("null context");
  Instance in;
  in.solver = std::move(solver);
  in.handlers = std::move(handlers);
  in.ctx = ctx;
  in.dimension = n;
  const int nranks = itsolv_comm_size(ctx), rank = itsolv_comm_rank(ctx);
  std::vector<int64_t> borders(size_t(nranks) + 1);
  itsolv_distribution(n, nranks, borders.data());
  in.n_local = size_t(borders[rank + 1] - borders[rank]);
  if (range_begin)
    *range_begin = size_t(borders[rank]);
  if (range_end)
    *range_end = size_t(borders[rank + 1]);
  instances.push(std::move(in));
  return instances.top();
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
template <class F>
long guarded_count(F&& f) {
  try {
    return long(f());
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}

} // namespace

extern "C" {

const char* ItsolvB200LastError(void) { return g_error.c_str(); }

int ItsolvB200LinearEigensystemInitialize(itsolv_ctx* ctx, size_t n, size_t nroot, size_t* range_begin,
                                          size_t* range_end, double thresh, double thresh_value, int hermitian,
                                          int verbosity, const char* options) {
  return guarded([&] {
    const auto opt = parse_options(options, {"max_size_qspace", "reset_d", "reset_d_max_q_size", "max_iter", "fused"});
    auto handlers = itsolv_b200::make_handlers();
    const bool fused = !opt.count("fused") || opt.at("fused") != 0;
    std::unique_ptr<its::LinearEigensystemDavidson<R, R, P>> solver;
    if (fused)
      solver = std::make_unique<itsolv_b200::LinearEigensystemDavidsonFused>(handlers);
    else
      solver = std::make_unique<its::LinearEigensystemDavidson<R, R, P>>(handlers);
    solver->set_n_roots(nroot);
    solver->set_verbosity(verbosity);
    solver->set_hermiticity(hermitian != 0);
    solver->set_convergence_threshold(thresh);
    solver->set_convergence_threshold_value(thresh_value);
    apply_dspace_options(*solver, opt);
    if (opt.count("max_iter"))
      solver->set_max_iter(int(opt.at("max_iter")));
    set_logger(*solver->logger, verbosity);
    auto& in = push_instance(ctx, n, std::move(solver), handlers, range_begin, range_end);
    in.has_eigenvalues = true;
  });
}

int ItsolvB200LinearEquationsInitialize(itsolv_ctx* ctx, size_t n, size_t nroot, size_t* range_begin,
                                        size_t* range_end, const double* rhs, double aughes, double thresh,
                                        double thresh_value, int hermitian, int verbosity, const char* options) {
  return guarded([&] {
    const auto opt = parse_options(options, {"max_size_qspace", "reset_d", "reset_d_max_q_size", "max_iter", "fused"});
    auto handlers = itsolv_b200::make_handlers();
    std::unique_ptr<its::LinearEquationsDavidson<R, R, P>> solver;
    if (opt.count("fused") && opt.at("fused") != 0)
      solver = std::make_unique<itsolv_b200::LinearEquationsDavidsonFused>(handlers);
    else
      solver = std::make_unique<its::LinearEquationsDavidson<R, R, P>>(handlers);
    auto* s = solver.get();
    auto& in = push_instance(ctx, n, std::move(solver), handlers, range_begin, range_end);
    try {
      auto rr = views(in, nroot, const_cast<double*>(rhs));
      s->set_hermiticity(hermitian != 0);
      s->set_n_roots(nroot);
      s->add_equations(rr); // the solver keeps its own copies (reference subspace/XSpace.h:208-220)
      s->set_convergence_threshold(thresh);
      s->set_convergence_threshold_value(thresh_value);
      if (aughes != 0)
        s->set_augmented_hessian(aughes);
      apply_dspace_options(*s, opt);
      if (opt.count("max_iter"))
        s->set_max_iter(int(opt.at("max_iter")));
      set_logger(*s->logger, verbosity);
      s->set_verbosity(verbosity);
    } catch (...) {
      instances.pop();
      throw;
    }
  });
}

int ItsolvB200NonLinearEquationsInitialize(itsolv_ctx* ctx, size_t n, size_t* range_begin, size_t* range_end,
                                           double thresh, int verbosity, const char* options) {
  return guarded([&] {
    const auto opt = parse_options(options, {"max_size_qspace", "max_iter", "fused"});
    auto handlers = itsolv_b200::make_handlers();
    std::unique_ptr<its::NonLinearEquationsDIIS<R, R, P>> solver;
    if (opt.count("fused") && opt.at("fused") != 0)
      solver = std::make_unique<itsolv_b200::NonLinearEquationsDIISFused>(handlers);
    else
      solver = std::make_unique<its::NonLinearEquationsDIIS<R, R, P>>(handlers);
    solver->set_convergence_threshold(thresh);
    solver->set_verbosity(verbosity);
    if (opt.count("max_size_qspace"))
      solver->set_max_size_qspace(int(opt.at("max_size_qspace")));
    if (opt.count("max_iter"))
      solver->set_max_iter(int(opt.at("max_iter")));
    push_instance(ctx, n, std::move(solver), handlers, range_begin, range_end);
  });
}

int ItsolvB200Finalize(void) {
  return guarded([] {
    top();
    instances.pop();
  });
}

long ItsolvB200AddVector(size_t buffer_size, double* parameters, double* action) {
  return guarded_count([&] {
    auto& in = top();
    auto cc = views(in, buffer_size, parameters);
    auto gg = views(in, buffer_size, action);
    return in.solver->add_vector(cc, gg);
  });
}

int ItsolvB200Solution(int nroot, const int* roots, double* parameters, double* action) {
  return guarded([&] {
    auto& in = top();
    if (nroot < 0 || (nroot > 0 && !roots))
      throw std::invalid_argument("invalid list of roots");
    auto cc = views(in, size_t(nroot), parameters);
    auto gg = views(in, size_t(nroot), action);
    in.solver->solution(std::vector<int>(roots, roots + nroot), cc, gg);
  });
}

long ItsolvB200EndIteration(size_t buffer_size, double* solution, double* residual) {
  return guarded_count([&] {
    auto& in = top();
    auto cc = views(in, buffer_size, solution);
    auto gg = views(in, buffer_size, residual);
    return in.solver->end_iteration(cc, gg);
  });
}

int ItsolvB200EndIterationNeeded(void) {
  try {
    return top().solver->end_iteration_needed() ? 1 : 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}

long ItsolvB200AddP(size_t buffer_size, size_t nP, const size_t* offsets, const size_t* indices,
                    const double* coefficients, const double* pp, double* parameters, double* action,
                    void (*func)(const double*, double*, const size_t, const size_t*)) {
  return guarded_count([&] {
    auto& in = top();
    in.apply_on_p = func;
    auto cc = views(in, buffer_size, parameters);
    auto gg = views(in, buffer_size, action);
    std::vector<P> pvectors;
    pvectors.reserve(nP);
    for (size_t p = 0; p < nP; ++p) {
      P map;
      for (size_t k = offsets[p]; k < offsets[p + 1]; ++k)
        map.emplace(indices[k], coefficients[k]);
      pvectors.emplace_back(std::move(map));
    }
    // the caller's function receives the coefficients of all vectors, the FIRST action vector's device pointer (the
    // others follow with the shard stride) and the row range per vector (IterativeSolverCMPI.cpp:117-137)
    Solver::fapply_on_p_type apply = [](const std::vector<std::vector<double>>& pcoeff, const its::CVecRef<P>&,
                                       const its::VecRef<R>& act) {
      auto& inst = top();
      std::vector<size_t> ranges;
      std::vector<double> flat;
      for (size_t k = 0; k < pcoeff.size(); ++k) {
        ranges.push_back(act[k].get().local_start());
        ranges.push_back(act[k].get().local_start() + act[k].get().local_size());
        flat.insert(flat.end(), pcoeff[k].begin(), pcoeff[k].end());
      }
      if (!inst.apply_on_p)
        throw std::runtime_error("no function for the P-space part of the action was given");
      inst.apply_on_p(flat.data(), act.front().get().data(), pcoeff.size(), ranges.data());
    };
    return in.solver->add_p(its::cwrap(pvectors),
                            molpro::linalg::array::Span<double>(const_cast<double*>(pp),
                                                                (in.solver->dimensions().oP + nP) * nP),
                            its::wrap(cc), its::wrap(gg), apply);
  });
}

int ItsolvB200Errors(double* errors) {
  return guarded([&] {
    size_t k = 0;
    for (const auto& e : top().solver->errors())
      errors[k++] = e;
  });
}

int ItsolvB200Eigenvalues(double* eigenvalues) {
  return guarded([&] {
    auto* s = dynamic_cast<its::LinearEigensystem<R, R, P>*>(top().solver.get());
    size_t k = 0;
    if (s)
      for (const auto& e : s->eigenvalues())
        eigenvalues[k++] = e;
  });
}

int ItsolvB200WorkingSetEigenvalues(double* eigenvalues) {
  return guarded([&] {
    size_t k = 0;
    for (const auto& e : top().solver->working_set_eigenvalues())
      eigenvalues[k++] = e;
  });
}

long ItsolvB200WorkingSet(int* roots) {
  return guarded_count([&] {
    const auto& ws = top().solver->working_set();
    for (size_t k = 0; k < ws.size(); ++k)
      roots[k] = ws[k];
    return ws.size();
  });
}

int ItsolvB200NonLinear(void) {
  try {
    return top().solver->nonlinear() ? 1 : 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}

long ItsolvB200SuggestP(const double* solution, const double* residual, size_t maximumNumber, double threshold,
                        size_t* indices) {
  return guarded_count([&] {
    auto& in = top();
    const size_t nroots = in.solver->n_roots();
    auto cc = views(in, nroots, const_cast<double*>(solution));
    auto gg = views(in, nroots, const_cast<double*>(residual));
    const auto result = in.solver->suggest_p(its::cwrap(cc), its::cwrap(gg), maximumNumber, threshold);
    for (size_t i = 0; i < result.size(); ++i)
      indices[i] = result[i];
    return result.size();
  });
}

int ItsolvB200PrintStatistics(void) {
  return guarded([] { molpro::cout << top().solver->statistics() << std::endl; });
}

int ItsolvB200HasValues(void) {
  try {
    top();
    return 0; // only the reference's Optimize instances carry values (IterativeSolverCMPI.cpp:229,480)
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}

double ItsolvB200Value(void) {
  try {
    return top().solver->value();
  } catch (const std::exception& e) {
    g_error = e.what();
    return std::numeric_limits<double>::quiet_NaN();
  }
}

int ItsolvB200HasEigenvalues(void) {
  try {
    return top().has_eigenvalues ? 1 : 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}

int ItsolvB200SetDiagonals(const double* diagonals) {
  return guarded([&] {
    auto& in = top();
    auto v = views(in, 1, const_cast<double*>(diagonals));
    in.diagonals.reset(new R(v.front())); // an owning copy
  });
}

int ItsolvB200Diagonals(double* diagonals) {
  return guarded([&] {
    auto& in = top();
    if (!in.diagonals)
      throw std::runtime_error("no diagonals have been set");
    views(in, 1, diagonals).front().copy(*in.diagonals);
  });
}

int ItsolvB200PreconditionDefault(size_t nwork, double* residual) {
  return guarded([&] {
    auto& in = top();
    if (!in.diagonals)
      throw std::runtime_error("no diagonals have been set");
    const auto shift = in.solver->working_set_eigenvalues();
    if (nwork > shift.size())
      throw std::invalid_argument("more residuals than vectors in the working set");
    std::vector<double*> r(nwork);
    for (size_t k = 0; k < nwork; ++k)
      r[k] = residual + k * in.n_local;
    itsolv_b200::check(itsolv_precondition_f64(in.ctx, r.data(), int(nwork), in.diagonals->data(), shift.data(),
                                               in.n_local),
                       "ItsolvB200PreconditionDefault");
  });
}

int ItsolvB200Verbosity(void) {
  try {
    switch (top().solver->get_verbosity()) {
    case its::Verbosity::None:
      return 0;
    case its::Verbosity::Summary:
      return 1;
    case its::Verbosity::Iteration:
      return 2;
    case its::Verbosity::Detailed:
      return 3;
    }
  } catch (const std::exception& e) {
    g_error = e.what();
  }
  return -1;
}

int ItsolvB200MaxIter(void) {
  try {
    return top().solver->get_max_iter();
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}

int ItsolvB200SetMaxIter(int max_iter) {
  return guarded([&] { top().solver->set_max_iter(max_iter); });
}

long ItsolvB200Iterations(void) {
  return guarded_count([&] { return top().solver->statistics().iterations; });
}

} // extern "C"
