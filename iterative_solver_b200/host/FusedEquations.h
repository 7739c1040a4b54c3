// LinearEquationsDavidson and NonLinearEquationsDIIS of the reference on the fused X space (FusedDavidson.h,
// XSpaceFused): every new set of vectors is contracted with the whole subspace - parameters and actions of Q and D, the
// right-hand sides - in ONE Gram launch instead of the w(w+1)/2 dots and 3-6 separate contractions of
// xspace::update_qspace_data (reference itsolv/subspace/XSpace.h:31-83), and the D-space overlaps take two launches
// (reference :85-187). Everything else - solution, residual, proposal of new vectors, all decisions - is the reference's
// own code on the CUDA handlers.
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

class LinearEquationsDavidsonFused
    : public its::LinearEquationsDavidson<DistrArrayCUDA, DistrArrayCUDA, std::map<size_t, double>> {
public:
  using Base = its::LinearEquationsDavidson<DistrArrayCUDA, DistrArrayCUDA, std::map<size_t, double>>;
  explicit LinearEquationsDavidsonFused(const std::shared_ptr<HandlersCUDA>& handlers,
                                        const std::shared_ptr<its::Logger>& logger_ = std::make_shared<its::Logger>())
      : Base(handlers, logger_) {
    this->m_xspace = std::make_shared<XSpaceFused>(handlers, logger_);
    this->set_hermiticity(this->get_hermiticity());
  }
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
