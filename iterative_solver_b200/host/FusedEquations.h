// The equation solvers of the reference on the fused path.
// LinearEquationsDavidsonFused: the complete fused driver of FusedDavidson.h (FusedDriver) with the residual form of
// LinearEquationsDavidson.h:173-184.
// NonLinearEquationsDIISFused: the reference's DIIS class on the fused X space (XSpaceFused): every new vector is
// contracted with the whole subspace in ONE Gram launch instead of the dots and separate contractions of
// xspace::update_qspace_data (reference itsolv/subspace/XSpace.h:31-83); solution and update are the reference's own code
// on the CUDA handlers.
#ifndef ITSOLV_B200_HOST_FUSEDEQUATIONS_H
#define ITSOLV_B200_HOST_FUSEDEQUATIONS_H
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

class NonLinearEquationsDIISFused
    : public its::NonLinearEquationsDIIS<DistrArrayCUDA, DistrArrayCUDA, std::map<size_t, double>> {
public:
  using Base = its::NonLinearEquationsDIIS<DistrArrayCUDA, DistrArrayCUDA, std::map<size_t, double>>;
  explicit NonLinearEquationsDIISFused(const std::shared_ptr<HandlersCUDA>& handlers,
                                       const std::shared_ptr<its::Logger>& logger_ = std::make_shared<its::Logger>())
      : Base(handlers, logger_) {
    auto xspace = std::make_shared<XSpaceFused>(handlers, logger_);
    xspace->set_hermiticity(true); // as the reference's constructor configures its X space (NonLinearEquationsDIIS.h:44-46)
    xspace->set_action_action();
    this->m_xspace = xspace;
  }
};

} // namespace itsolv_b200
#endif
