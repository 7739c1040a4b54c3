// The equation solvers of the reference on the fused path.
// LinearEquationsDavidsonFused: the complete fused driver of FusedDavidson.h (FusedDriver) with the residual form of
// LinearEquationsDavidson.h:173-184.
// NonLinearEquationsDIISFused: the reference's DIIS class on the fused X space (XSpaceFused: every new vector is
// contracted with the whole subspace in ONE Gram launch instead of the dots and separate contractions of
// xspace::update_qspace_data, reference itsolv/subspace/XSpace.h:31-83) with extrapolation, preconditioner and DIIS step
// in ONE pass over the subspace.
#ifndef ITSOLV_B200_HOST_FUSEDEQUATIONS_H
#define ITSOLV_B200_HOST_FUSEDEQUATIONS_H
#include <algorithm>
#include <iostream>
#include <map>
#include <memory>

#include <molpro/linalg/itsolv/LinearEigensystemDavidson.h>
// LinearEquationsDavidson.h is not self-contained; it needs the includes of LinearEigensystemDavidson.h first.
#include <molpro/linalg/itsolv/LinearEquationsDavidson.h>
#include <molpro/linalg/itsolv/NonLinearEquationsDIIS.h>

#include "FusedDavidson.h"

namespace itsolv_b200 {

//! LinearEquationsDavidson of the reference on the fused driver (FusedDavidson.h, FusedDriver): one Gram launch per set of
//! new vectors, solutions + residuals (sum_i c a_i - b) / |b| + norms + diagonal preconditioner in one pass over the
//! subspace, the fused proposal step; the same decisions, thresholds and subspace solver as the reference's class
class LinearEquationsDavidsonFused
    : public FusedDriver<its::LinearEquationsDavidson<DistrArrayCUDA, DistrArrayCUDA, std::map<size_t, double>>> {
public:
  using Base = its::LinearEquationsDavidson<DistrArrayCUDA, DistrArrayCUDA, std::map<size_t, double>>;
  explicit LinearEquationsDavidsonFused(const std::shared_ptr<HandlersCUDA>& handlers,
                                        const std::shared_ptr<its::Logger>& logger_ = std::make_shared<its::Logger>())
      : FusedDriver<Base>(handlers, logger_) {}

protected:
  ResidualForm fused_residual_form(const std::vector<int>& roots) const override {
    ResidualForm f;
    f.mode = 1;
    const auto xspace = std::dynamic_pointer_cast<its::subspace::XSpace<R, R, P>>(this->m_xspace);
    const auto& norm = xspace->rhs_norm();
    const auto all_rhs = this->rhs();
    for (auto root : roots) {
      f.rhs.emplace_back(all_rhs.at(size_t(root)));
      const double nrm = norm.at(size_t(root));
      f.rscale.push_back(nrm != 0 ? 1 / nrm : 1.0); // reference LinearEquationsDavidson.h:179-182
    }
    f.shift.assign(roots.size(), 0.0); // working_set_eigenvalues() of this solver (IterativeSolver.h:320-322)
    return f;
  }
  double fused_norm_thresh() const override { return this->m_norm_thresh; }
  double fused_svd_thresh() const override { return this->m_svd_thresh; }
  int fused_max_size_qspace() const override { return this->m_max_size_qspace; }
  its::detail::DSpaceResetter<R>& fused_resetter() override { return this->m_dspace_resetter; }
};

/*!
 * NonLinearEquationsDIIS of the reference with the O(n) work of an iteration in three passes instead of eleven.
 * The reference's iteration (IterativeSolverTemplate.h:380-398, NonLinearEquationsDIIS.h:87-119) on the handlers:
 *   dot(r, r); update_qspace (a dot and 3 contractions); solution(): 2 x (fill + gemm_outer [q x 1]); dot(r', r');
 *   copy of the diagonal into the parameters; precondition; solution_params(): fill + gemm_outer again (the parameters
 *   were overwritten by the diagonal); axpy(-1, r', x)
 * Here: dot(r, r); ONE Gram launch (XSpaceFused); ONE pass over the subspace that forms x = sum c q and r' = sum c a,
 * returns <r', r'>, divides r' by the diagonal and stores x - r' (itsolv_subspace_residual_f64, mode 3) - every
 * operation rounded as in the separate calls. The subspace solver, the deletion of the least important vector and the
 * convergence logic are the reference's own functions (add_vector is called as it is; solution() and end_iteration()
 * are the virtual functions it calls). On convergence the step is not wanted: the parameters are formed again without
 * it by the reference's end_iteration (once per solve).
 */
class NonLinearEquationsDIISFused
    : public its::NonLinearEquationsDIIS<DistrArrayCUDA, DistrArrayCUDA, std::map<size_t, double>> {
public:
  using R = DistrArrayCUDA;
  using P = std::map<size_t, double>;
  using Base = its::NonLinearEquationsDIIS<R, R, P>;
  template <class T>
  using VecRef = its::VecRef<T>;
  explicit NonLinearEquationsDIISFused(const std::shared_ptr<HandlersCUDA>& handlers,
                                       const std::shared_ptr<its::Logger>& logger_ = std::make_shared<its::Logger>())
      : Base(handlers, logger_) {
    auto xspace = std::make_shared<XSpaceFused>(handlers, logger_);
    xspace->set_hermiticity(true); // as the reference's constructor configures its X space (NonLinearEquationsDIIS.h:44-46)
    xspace->set_action_action();
    this->m_xspace = xspace;
    m_dense = dynamic_cast<ArrayHandlerCUDA*>(&handlers->rr());
  }

  bool solve(const VecRef<R>& parameters, const VecRef<R>& actions, const its::Problem<R>& problem,
             bool generate_initial_guess = false) override {
    // the fused pass applies the default diagonal preconditioner; anything else takes the reference's loop
    const bool default_preconditioner = dynamic_cast<const UsesDefaultDiagonalPreconditioner*>(&problem) != nullptr;
    if (!m_dense || !default_preconditioner || generate_initial_guess || this->m_max_p > 0 || parameters.empty() ||
        parameters.size() != actions.size())
      return Base::solve(parameters, actions, problem, generate_initial_guess);
    std::unique_ptr<R> diagonals(new R(actions.at(0).get().size(), actions.at(0).get().context()));
    if (!problem.diagonals(*diagonals))
      return Base::solve(parameters, actions, problem, generate_initial_guess);
    this->m_logger->max_trace_level = its::Logger::None;
    if (this->m_verbosity == its::Verbosity::Detailed) {
      this->m_logger->max_trace_level = its::Logger::Info;
      this->m_logger->data_dump = true;
    }
    struct Scope {
      const R*& slot;
      ~Scope() { slot = nullptr; }
    } scope{m_step_diagonals};
    m_step_diagonals = diagonals.get();
    int nwork = int(parameters.size());
    for (int iter = 0; iter < this->m_max_iter && nwork > 0; iter++) {
      const auto value = problem.residual(parameters.front().get(), actions.front().get());
      m_step_taken = false;
      nwork = this->add_vector(parameters.front().get(), actions.front().get(), value);
      while (this->end_iteration_needed()) {
        if (nwork > 0 && !m_step_taken) { // solution() took the general route: the reference's sequence
          this->m_handlers->rq().copy(parameters.at(0), *diagonals);
          problem.precondition(its::wrap(actions.begin(), actions.begin() + nwork), this->working_set_eigenvalues(),
                               parameters.at(0));
        }
        nwork = int(this->end_iteration(parameters, actions));
      }
      if (this->m_verbosity >= its::Verbosity::Iteration)
        this->report();
    }
    if (this->m_verbosity == its::Verbosity::Summary)
      this->report();
    const double worst = *std::max_element(this->m_errors.begin(), this->m_errors.end());
    if (this->m_verbosity >= its::Verbosity::Summary && worst > this->m_convergence_threshold)
      std::cerr << "Solver has not converged to threshold " << this->m_convergence_threshold << std::endl;
    return nwork == 0 && worst <= this->m_convergence_threshold;
  }

  //! extrapolated parameters and residual (reference IterativeSolverTemplate.h:191-215) in one pass; inside the fused
  //! solve() the pass also preconditions the residual and takes the DIIS step
  void solution(const std::vector<int>& roots, const VecRef<R>& parameters, const VecRef<R>& residual) override {
    auto& xs = *this->m_xspace;
    const auto dims = xs.dimensions();
    if (!m_dense || roots.size() != 1 || parameters.empty() || residual.empty() || dims.nP != 0 || dims.nD != 0 ||
        dims.nQ == 0 || this->m_normalise_solution || this->m_apply_p)
      return Base::solution(roots, parameters, residual);
    this->check_consistent_number_of_roots_and_solutions(roots, parameters.size());
    const auto& sol = this->m_subspace_solver->solutions();
    its::subspace::Matrix<double> c({dims.nQ, 1});
    for (size_t j = 0; j < dims.nQ; ++j)
      c(j, 0) = sol(roots[0], dims.oQ + j);
    const bool step = m_step_diagonals != nullptr;
    const VecRef<R> par(parameters.begin(), parameters.begin() + 1), res(residual.begin(), residual.begin() + 1);
    const std::vector<double> none, shift{0.0}; // working_set_eigenvalues() of a non-linear solver (IterativeSolver.h:320)
    const auto norms = m_dense->subspace_residual(step ? 3 : 2, false, c, xs.cparamsq(), xs.cactionsq(), none,
                                                  its::CVecRef<R>{}, none, step ? m_step_diagonals : nullptr, shift, par, res);
    // update_errors asks for <r', r'> of the residual BEFORE the preconditioner next (IterativeSolverTemplate.h:533-534)
    m_dense->prime_self_dots(its::cwrap(res), norms.residual);
    m_step_taken = step;
    its::read_handler_counts(this->m_stats, this->m_handlers);
  }

  size_t end_iteration(const VecRef<R>& parameters, const VecRef<R>& action) override {
    const bool converged = this->m_errors.front() < this->m_convergence_threshold;
    if (m_step_taken && !converged && this->m_working_set.empty()) {
      // The reference preconditions a non-empty working set only (IterativeSolverTemplate.h:388-396). Here the
      // extrapolated residual is below the threshold while the last computed one is not: the step is then taken with the
      // residual as it is. Rare; the extrapolation is formed again without the preconditioner.
      const R* diagonals = m_step_diagonals;
      m_step_diagonals = nullptr;
      solution(std::vector<int>{0}, parameters, action);
      m_step_diagonals = diagonals;
      m_step_taken = false;
      return Base::end_iteration(parameters, action);
    }
    if (!m_step_taken || converged) { // converged: the parameters without the step, formed by the reference's function
      m_step_taken = false;
      return Base::end_iteration(parameters, action);
    }
    // NonLinearEquationsDIIS.h:103-119 with solution_params() and the axpy already done by solution()
    m_step_taken = false;
    this->m_end_iteration_needed = false;
    this->m_working_set.assign(1, 0);
    this->m_stats->iterations++;
    return 1;
  }
  size_t end_iteration(std::vector<R>& parameters, std::vector<R>& action) override {
    return end_iteration(its::wrap(parameters), its::wrap(action));
  }

private:
  ArrayHandlerCUDA* m_dense = nullptr;
  const R* m_step_diagonals = nullptr; //!< inside the fused solve(): the diagonal the step is preconditioned with
  bool m_step_taken = false;           //!< solution() has preconditioned the residual and stored x - r'
};

} // namespace itsolv_b200
#endif
