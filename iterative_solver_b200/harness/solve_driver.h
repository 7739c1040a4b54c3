// Container-agnostic driver around the reference's IterativeSolver::solve()
// (reference src/molpro/linalg/itsolv/IterativeSolverTemplate.h:322-408).
//
// The driver only configures and calls the reference's solver classes; it performs no vector arithmetic itself.
// It is instantiated twice: in the product harness with R = DistrArrayCUDA (harness/solver_capi.cpp) and in the
// oracle build of the reference with R = std::vector<double> (oracle/ref_driver.cpp), so that both sides of a parity
// test are configured by the same code from the same itsolv_solve_spec.
#ifndef ITSOLV_B200_HARNESS_SOLVE_DRIVER_H
#define ITSOLV_B200_HARNESS_SOLVE_DRIVER_H
#include <chrono>
#include <cstring>
#include <map>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <vector>

#include <itsolv_b200_harness.h>
#include <molpro/linalg/itsolv/LinearEigensystemDavidson.h>
// LinearEquationsDavidson.h is not self-contained; it needs the includes of LinearEigensystemDavidson.h first.
#include <molpro/linalg/itsolv/LinearEquationsDavidson.h>
#include <molpro/linalg/itsolv/NonLinearEquationsDIIS.h>

namespace itsolv_b200::harness {

//! Every inner product the handlers hand back to the solver, in call order (itsolv_solve_spec::trace).
struct Trace {
  bool enabled = false;
  std::vector<itsolv_trace_entry> entries;
  std::vector<double> values;
  void clear() {
    entries.clear();
    values.clear();
  }
  void record(char op, size_t rows, size_t cols, const double* data) {
    if (!enabled)
      return;
    entries.push_back({int32_t(op), int32_t(rows), int32_t(cols), 0, int64_t(values.size())});
    values.insert(values.end(), data, data + rows * cols);
  }
};
inline Trace& trace() {
  static Trace t;
  return t;
}

inline double now_seconds() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

/*!
 * @tparam Backend provides
 *   using R;                                          container type
 *   std::shared_ptr<ArrayHandlers<R,R,P>> handlers(); handler set
 *   R make_vector();                                  zero vector of the problem's global length
 *   R make_output_vector();                           a vector the solve writes before it reads it (may be uninitialised)
 *   void export_local(const R&, double*);             this rank's rows to the caller's memory
 *   double* solutions_target(double* given, size_t count);  where the solutions go: `given`, or memory the backend
 *                                                     provides at that moment (after the solver has finished)
 *   size_t n_local();
 *   ProblemT& problem();                              Problem<R> with make_rhs(k, R&), seconds_action, seconds_precond
 *   void synchronize();                               wait for outstanding device work (no-op on the host)
 *   void timer_start(); double timer_stop_ms();       device stopwatch around solve() (0 on the host)
 *   unique_ptr<LinearEigensystemDavidson<R,R,P>> make_davidson(handlers, spec);   the reference's class or a subclass
 *   unique_ptr<LinearEquationsDavidson<R,R,P>> make_lineq(handlers, spec);
 *   unique_ptr<NonLinearEquationsDIIS<R,R,P>> make_diis(handlers, spec);
 */
template <class Backend>
int run_solve(const itsolv_solve_spec& spec, Backend& backend, itsolv_solve_result& res, double* solutions) {
  using R = typename Backend::R;
  using P = std::map<size_t, double>;
  namespace its = molpro::linalg::itsolv;
  std::memset(&res, 0, sizeof(res));
  const int nroots = spec.kind == ITSOLV_KIND_DIIS ? 1 : spec.nroots;
  if (nroots < 1 || nroots > ITSOLV_MAX_ROOTS)
    throw std::invalid_argument("itsolv harness: nroots out of range");
  const int nbuf = spec.kind == ITSOLV_KIND_DIIS ? 1 : (spec.nbuffers > 0 ? std::min(spec.nbuffers, nroots) : nroots);
  auto handlers = backend.handlers();
  auto& problem = backend.problem();
  std::vector<R> parameters, actions;
  parameters.reserve(nbuf);
  actions.reserve(nbuf);
  // Davidson and LinearEquations start from a guess that is written in full (unit vectors by the solver, or a copy of
  // the right-hand sides); DIIS starts from the zero vector
  const bool guess_overwrites = spec.kind != ITSOLV_KIND_DIIS;
  for (int i = 0; i < nbuf; ++i) {
    parameters.emplace_back(guess_overwrites ? backend.make_output_vector() : backend.make_vector());
    actions.emplace_back(backend.make_output_vector()); // written by the operator / the solver before anything reads them
  }
  trace().clear();
  trace().enabled = spec.trace != 0;

  auto configure = [&spec](auto& solver) {
    solver.set_verbosity(int(spec.verbosity));
    if (spec.max_iter > 0)
      solver.set_max_iter(spec.max_iter);
    if (spec.convergence_threshold > 0)
      solver.set_convergence_threshold(spec.convergence_threshold);
    if (spec.max_size_qspace > 0)
      solver.set_max_size_qspace(spec.max_size_qspace);
  };
  auto finish = [&](auto& solver, bool converged, double t0) {
    res.device_ms_solve = backend.timer_stop_ms();
    backend.synchronize();
    res.seconds_solve = now_seconds() - t0;
    res.converged = converged ? 1 : 0;
    res.iterations = int32_t(solver.statistics().iterations);
    res.nroots = nroots;
    res.nwork_final = int32_t(solver.working_set().size());
    res.r_creations = solver.statistics().r_creations;
    res.q_creations = solver.statistics().q_creations;
    res.p_creations = solver.statistics().p_creations;
    res.d_creations = solver.statistics().d_creations;
    const auto& err = solver.errors();
    for (size_t i = 0; i < err.size() && i < ITSOLV_MAX_ROOTS; ++i)
      res.errors[i] = err[i];
    res.seconds_action = problem.seconds_action;
    res.seconds_precond = problem.seconds_precond;
  };
  auto export_solutions = [&](auto& solver) {
    solutions = backend.solutions_target(solutions, size_t(nroots) * backend.n_local());
    if (!solutions)
      return;
    const size_t nloc = backend.n_local();
    for (int start = 0; start < nroots; start += nbuf) {
      const int end = std::min(start + nbuf, nroots);
      std::vector<int> roots(end - start);
      std::iota(roots.begin(), roots.end(), start);
      auto wp = its::wrap(parameters.begin(), parameters.begin() + roots.size());
      auto wa = its::wrap(actions.begin(), actions.begin() + roots.size());
      solver.solution(roots, wp, wa);
      for (size_t i = 0; i < roots.size(); ++i)
        backend.export_local(parameters[i], solutions + size_t(roots[i]) * nloc);
    }
  };

  problem.seconds_action = problem.seconds_precond = 0;
  if (spec.kind == ITSOLV_KIND_DAVIDSON) {
    auto solver_ptr = backend.make_davidson(handlers, spec); // the reference's class, or a subclass of it
    auto& solver = *solver_ptr;
    configure(solver);
    solver.set_n_roots(nroots);
    solver.set_hermiticity(spec.hermitian != 0);
    if (spec.reset_D > 0)
      solver.set_reset_D(spec.reset_D);
    if (spec.max_p > 0)
      solver.set_max_p(spec.max_p);
    backend.synchronize();
    const double t0 = now_seconds();
    backend.timer_start();
    const bool ok = solver.solve(parameters, actions, problem, true);
    finish(solver, ok, t0);
    const auto ev = solver.eigenvalues();
    for (size_t i = 0; i < ev.size() && i < size_t(nroots); ++i)
      res.eigenvalues[i] = ev[i];
    export_solutions(solver);
  } else if (spec.kind == ITSOLV_KIND_LINEQ) {
    auto solver_ptr = backend.make_lineq(handlers, spec);
    auto& solver = *solver_ptr;
    configure(solver);
    solver.set_hermiticity(spec.hermitian != 0);
    if (spec.reset_D > 0)
      solver.set_reset_D(spec.reset_D);
    if (spec.max_p > 0)
      solver.set_max_p(spec.max_p);
    // Initial guess: the solver's own (unit vectors on the smallest diagonal elements, generate_initial_guess = true, as
    // the reference's test does, test/itsolv/test_LinearEquations.cpp:83); with the legacy inputs c = rhs. The latter puts a
    // vector with |A c| ~ n |c| into the subspace, whose coefficient then has to vanish to ~1e-8 / n: the error after a
    // fixed number of iterations grows with n (3e-10 at n = 1e6, 1.5e-9 at 1e7 on the reference's CPU path).
    const bool default_guess = spec.rhs_kind != ITSOLV_RHS_LEGACY;
    {
      // right-hand sides are built one at a time in a scratch R vector; the solver keeps its own Q copies
      // (XSpace::add_rhs_equations, reference subspace/XSpace.h:208-220)
      for (int k = 0; k < nroots; ++k) {
        R rhs = backend.make_vector();
        problem.make_rhs(k, rhs);
        solver.add_equations(rhs);
        if (k < nbuf && !default_guess)
          handlers->rr().copy(parameters[k], rhs); // initial guess c = rhs (reference test_simplified.cpp:131-134)
      }
    }
    backend.synchronize();
    const double t0 = now_seconds();
    backend.timer_start();
    const bool ok = solver.solve(parameters, actions, problem, default_guess);
    finish(solver, ok, t0);
    export_solutions(solver);
  } else if (spec.kind == ITSOLV_KIND_DIIS) {
    auto solver_ptr = backend.make_diis(handlers, spec);
    auto& solver = *solver_ptr;
    configure(solver);
    backend.synchronize();
    const double t0 = now_seconds();
    backend.timer_start();
    const bool ok = solver.solve(parameters[0], actions[0], problem, false);
    finish(solver, ok, t0);
    solutions = backend.solutions_target(solutions, backend.n_local());
    if (solutions) {
      solver.solution(parameters[0], actions[0]);
      backend.export_local(parameters[0], solutions);
    }
  } else {
    throw std::invalid_argument("itsolv harness: unknown solver kind");
  }
  backend.synchronize();
  trace().enabled = false;
  return 0;
}

} // namespace itsolv_b200::harness
#endif
