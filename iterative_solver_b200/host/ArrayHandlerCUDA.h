// ArrayHandlerCUDA / ArrayHandlerCUDASparse: the reference's vector-operation contract
// (molpro::linalg::array::ArrayHandler<AL,AR>, reference src/molpro/linalg/array/ArrayHandler.h:161-437) implemented
// on DistrArrayCUDA through the C ABI of include/itsolv_b200.h. They take the place that ArrayHandlerDistr /
// ArrayHandlerDistrSparse (reference array/ArrayHandlerDistr.h:14-73, ArrayHandlerDistrSparse.h:18-80) have for the
// MPI containers: inject them through the 7-argument ArrayHandlers constructor
// (reference itsolv/ArrayHandlers.h:30-34) or use make_handlers() below. No arithmetic is done on the host.
#ifndef ITSOLV_B200_HOST_ARRAYHANDLERCUDA_H
#define ITSOLV_B200_HOST_ARRAYHANDLERCUDA_H
#include <functional>
#include <map>
#include <memory>
#include <vector>

#include <molpro/linalg/array/ArrayHandler.h>
#include <molpro/linalg/array/ArrayHandlerSparse.h>
#include <molpro/linalg/itsolv/ArrayHandlers.h>

#include "DistrArrayCUDA.h"

namespace itsolv_b200 {

using molpro::linalg::array::ArrayHandler;
using molpro::linalg::itsolv::CVecRef;
using molpro::linalg::itsolv::VecRef;
using molpro::linalg::itsolv::subspace::Matrix;

//! Optional observer of every number a handler returns to the solver ('d' dot, 'g' gemm_inner); used by parity tests
using ResultObserver = std::function<void(char op, size_t rows, size_t cols, const double* values)>;

class ArrayHandlerCUDA : public ArrayHandler<DistrArrayCUDA, DistrArrayCUDA> {
public:
  using Base = ArrayHandler<DistrArrayCUDA, DistrArrayCUDA>;
  using AL = DistrArrayCUDA;
  using AR = DistrArrayCUDA;
  using typename Base::ProxyHandle;
  using typename Base::value_type;
  using typename Base::value_type_abs;

  ArrayHandlerCUDA() = default;
  explicit ArrayHandlerCUDA(ResultObserver observer) : m_observer(std::move(observer)) {}

  AL copy(const AR& source) override {
    this->m_counter->copy++;
    if (m_take_on_copy && !source.empty()) // see TakeOnCopy below
      return const_cast<AR&>(source).take_contents();
    return AL{source};
  }
  void copy(AL& x, const AR& y) override {
    this->m_counter->copy++;
    x.copy(y);
  }
  void scal(value_type alpha, AL& x) override {
    this->m_counter->scal++;
    x.scal(alpha);
  }
  void fill(value_type alpha, AL& x) override { x.fill(alpha); }
  void axpy(value_type alpha, const AR& x, AL& y) override {
    this->m_counter->axpy++;
    y.axpy(alpha, x);
  }
  value_type dot(const AL& x, const AR& y) override {
    this->m_counter->dot++;
    double d;
    if (!primed(x, y, d))
      d = x.dot(y);
    if (m_observer)
      m_observer('d', 1, 1, &d);
    return d;
  }

  // ---- extensions used by the fused driver path (FusedDavidson.h); not part of the reference's contract ----

  //! <x_i, x_i> of every vector in one launch
  std::vector<double> self_dots(const CVecRef<AL>& xx) {
    std::vector<double> d(xx.size());
    if (xx.empty())
      return d;
    const auto G = gemm_inner(xx, xx);
    for (size_t i = 0; i < xx.size(); ++i)
      d[i] = G(i, i);
    return d;
  }
  //! computes <x_i, x_i> for all i now and answers the caller's coming dot(x_i, x_i) calls from the stored values, as
  //! long as nothing has been written in between (the context's write epoch stands still)
  void prime_self_dots(const CVecRef<AL>& xx) {
    m_primed.clear();
    if (xx.empty())
      return;
    const auto d = self_dots(xx);
    m_primed_epoch = itsolv_ctx_write_epoch(xx[0].get().context());
    for (size_t i = 0; i < xx.size(); ++i)
      m_primed.push_back({xx[i].get().data(), xx[i].get().data(), d[i]});
  }
  //! the same with values that a kernel has already returned (<x_i, x_i> from the tail of the pass that wrote x_i)
  void prime_self_dots(const CVecRef<AL>& xx, const std::vector<double>& values) {
    m_primed.clear();
    if (xx.empty())
      return;
    m_primed_epoch = itsolv_ctx_write_epoch(xx[0].get().context());
    for (size_t i = 0; i < xx.size() && i < values.size(); ++i)
      m_primed.push_back({xx[i].get().data(), xx[i].get().data(), values[i]});
  }
  //! x_k = alpha for every vector of the set
  void fill_batch(double alpha, const VecRef<AL>& xx) {
    if (xx.empty())
      return;
    std::vector<double*> px(xx.size());
    for (size_t i = 0; i < xx.size(); ++i) {
      xx[0].get().require_compatible(xx[i].get(), "fill_batch");
      px[i] = xx[i].get().data();
    }
    const std::vector<double> a(xx.size(), alpha);
    check(itsolv_fill_batch_f64(xx[0].get().context(), a.data(), px.data(), int(px.size()), xx[0].get().local_size()),
          "ArrayHandlerCUDA::fill_batch");
  }
  //! x_k *= alpha[k], one launch
  void scal_batch(const std::vector<double>& alpha, const VecRef<AL>& xx) {
    if (xx.empty())
      return;
    this->m_counter->scal += int(xx.size());
    std::vector<double*> px(xx.size());
    const AL& first = xx[0].get();
    for (size_t i = 0; i < xx.size(); ++i) {
      first.require_compatible(xx[i].get(), "scal_batch");
      px[i] = xx[i].get().data();
    }
    check(itsolv_scal_batch_f64(first.context(), alpha.data(), px.data(), int(px.size()), first.local_size()),
          "ArrayHandlerCUDA::scal_batch");
  }
  //! y_k += alpha[k] * x_k for independent pairs, one launch
  void axpy_batch(const std::vector<double>& alpha, const CVecRef<AR>& xx, const VecRef<AL>& yy) {
    if (yy.empty())
      return;
    this->m_counter->axpy += int(yy.size());
    std::vector<const double*> px(yy.size());
    std::vector<double*> py(yy.size());
    const AL& first = yy[0].get();
    for (size_t i = 0; i < yy.size(); ++i) {
      first.require_compatible(xx[i].get(), "axpy_batch");
      first.require_compatible(yy[i].get(), "axpy_batch");
      px[i] = xx[i].get().data();
      py[i] = yy[i].get().data();
    }
    check(itsolv_axpy_batch_f64(first.context(), alpha.data(), px.data(), py.data(), int(py.size()), first.local_size()),
          "ArrayHandlerCUDA::axpy_batch");
  }
  //! ri *= inv_norm; rj[k] -= ov[k] * ri  (one step of the R-R modified Gram-Schmidt), one launch
  void mgs_step(double inv_norm, AL& ri, const std::vector<double>& ov, const VecRef<AL>& rj) {
    this->m_counter->scal++;
    this->m_counter->axpy += int(rj.size());
    std::vector<double*> pj(rj.size());
    for (size_t j = 0; j < rj.size(); ++j) {
      ri.require_compatible(rj[j].get(), "mgs_step");
      pj[j] = rj[j].get().data();
    }
    check(itsolv_mgs_step_f64(ri.context(), inv_norm, ri.data(), ov.data(), pj.data(), int(pj.size()), ri.local_size()),
          "ArrayHandlerCUDA::mgs_step");
  }
  /*!
   * While an object of this type lives, copy(const AR&) -> AL hands over the source's allocation instead of copying it
   * and leaves the source with unspecified contents. The fused driver opens it around QSpace::update
   * (reference itsolv/subspace/QSpace.h:80-84), whose sources are R vectors that the solver overwrites next
   * (IterativeSolverTemplate.h:532): 2w vector copies per iteration become pointer swaps.
   */
  struct TakeOnCopy {
    explicit TakeOnCopy(ArrayHandlerCUDA& h) : handler(h), previous(h.m_take_on_copy) { h.m_take_on_copy = true; }
    ~TakeOnCopy() { handler.m_take_on_copy = previous; }
    TakeOnCopy(const TakeOnCopy&) = delete;
    TakeOnCopy& operator=(const TakeOnCopy&) = delete;
    ArrayHandlerCUDA& handler;
    bool previous;
  };

  struct ResidualNorms {
    std::vector<double> residual; //!< <r_j, r_j> before preconditioning
    std::vector<double> written;  //!< <out_j, out_j> of the vectors that were written
  };
  /*!
   * Solutions, residuals, their norms and the diagonal preconditioner of all roots in ONE pass over the subspace
   * (itsolv_davidson_residual_f64): x_j = sum_i c(i,j) q_i, r_j = sum_i c(i,j) a_i - lambda_j x_j,
   * out_r[j] = diag ? r_j / (diag - lambda_j + 1e-15) : r_j. `solutions` may be empty (x_j is then not stored).
   */
  ResidualNorms davidson_residual(const Matrix<value_type>& c, const CVecRef<AR>& q, const CVecRef<AR>& a,
                                  const std::vector<double>& lambda, const AR* diag, const VecRef<AL>& solutions,
                                  const VecRef<AL>& residuals) {
    return subspace_residual(0, false, c, q, a, lambda, CVecRef<AR>{}, std::vector<double>{}, diag, lambda, solutions,
                             residuals);
  }
  /*!
   * The same pass for both Davidson-type solvers (itsolv_subspace_residual_f64):
   *   mode 0: r_j = sum_i c(i,j) a_i - lambda_j x_j;  mode 1: r_j = (sum_i c(i,j) a_i - rhs_j) * rscale_j
   *   (reference LinearEquationsDavidson.h:173-184). `accumulate`: x_j and r_j continue from the present contents of
   *   `solutions` / `residuals` (the P-space parts) instead of zero. `shift`: of the diagonal preconditioner.
   */
  ResidualNorms subspace_residual(int mode, bool accumulate, const Matrix<value_type>& c, const CVecRef<AR>& q,
                                  const CVecRef<AR>& a, const std::vector<double>& lambda, const CVecRef<AR>& rhs,
                                  const std::vector<double>& rscale, const AR* diag, const std::vector<double>& shift,
                                  const VecRef<AL>& solutions, const VecRef<AL>& residuals) {
    const size_t k = c.rows(), m = c.cols();
    if (k > q.size() || k > a.size() || m > residuals.size() || (mode == 0 && m > lambda.size()) ||
        (mode == 1 && (m > rhs.size() || m > rscale.size())) || (diag && m > shift.size()) ||
        (!solutions.empty() && m > solutions.size()) || (accumulate && solutions.empty()))
      throw std::out_of_range("subspace_residual: dimensions of the coefficients and the vector sets are different.");
    ResidualNorms norms{std::vector<double>(m), std::vector<double>(m)};
    if (m == 0)
      return norms;
    if (k == 0)
      throw std::out_of_range("davidson_residual: empty subspace");
    this->m_counter->gemm_outer += 2;
    this->m_counter->axpy += int(m);
    this->m_counter->dot += int(m);
    const AL& first = residuals[0].get();
    std::vector<const double*> pq(k), pa(k), prhs(mode == 1 ? m : 0);
    std::vector<double*> px(solutions.empty() ? 0 : m), pr(m);
    for (size_t j = 0; j < prhs.size(); ++j) {
      first.require_compatible(rhs[j].get(), "subspace_residual");
      prhs[j] = rhs[j].get().data();
    }
    for (size_t i = 0; i < k; ++i) {
      first.require_compatible(q[i].get(), "davidson_residual");
      first.require_compatible(a[i].get(), "davidson_residual");
      pq[i] = q[i].get().data();
      pa[i] = a[i].get().data();
    }
    for (size_t j = 0; j < m; ++j) {
      first.require_compatible(residuals[j].get(), "davidson_residual");
      pr[j] = residuals[j].get().data();
      if (!solutions.empty()) {
        first.require_compatible(solutions[j].get(), "davidson_residual");
        px[j] = solutions[j].get().data();
      }
    }
    if (diag)
      first.require_compatible(*diag, "davidson_residual");
    check(itsolv_subspace_residual_f64(first.context(), mode, accumulate ? 1 : 0, c.data().data(), int(k), int(m),
                                       pq.data(), pa.data(), mode == 0 ? lambda.data() : nullptr,
                                       mode == 1 ? prhs.data() : nullptr, mode == 1 ? rscale.data() : nullptr,
                                       diag ? diag->data() : nullptr, diag ? shift.data() : nullptr,
                                       solutions.empty() ? nullptr : px.data(), pr.data(), first.local_size(),
                                       norms.residual.data(), norms.written.data()),
          "ArrayHandlerCUDA::subspace_residual");
    if (m_observer)
      m_observer('g', 1, m, norms.residual.data());
    return norms;
  }

  //! mgs_step that also returns {<ri', ri'>, <rj[0]', rj[t]'> for every t}: what the next step needs (at most 16 rj)
  std::vector<double> mgs_step_dots(double inv_norm, AL& ri, const std::vector<double>& ov, const VecRef<AL>& rj) {
    this->m_counter->scal++;
    this->m_counter->axpy += int(rj.size());
    this->m_counter->dot += int(rj.size()) + 1;
    std::vector<double*> pj(rj.size());
    for (size_t j = 0; j < rj.size(); ++j) {
      ri.require_compatible(rj[j].get(), "mgs_step_dots");
      pj[j] = rj[j].get().data();
    }
    std::vector<double> dots(rj.size() + 1);
    check(itsolv_mgs_step_dots_f64(ri.context(), inv_norm, ri.data(), ov.data(), pj.data(), int(pj.size()),
                                   ri.local_size(), dots.data()),
          "ArrayHandlerCUDA::mgs_step_dots");
    if (m_observer)
      m_observer('g', 1, dots.size(), dots.data());
    return dots;
  }
  static constexpr size_t max_mgs_step_dots = 16;
  //! can the R-R Gram-Schmidt of these vectors run as one chain of launches (itsolv_mgs_chain_f64)?
  bool mgs_chain_supported(const VecRef<AL>& rr) const {
    // decided from rank-independent inputs only (global length, number of ranks, options): every rank must take the
    // same branch, also one whose shard is empty
    if (rr.empty())
      return false;
    const auto& first = rr[0].get();
    const size_t nranks = size_t(itsolv_comm_size(first.context()));
    return first.size() >= nranks && itsolv_mgs_chain_supported(first.context(), int(rr.size()), first.size() / nranks) != 0;
  }
  //! the chain; returns the rows of inner products in the layout of include/itsolv_b200.h
  std::vector<double> mgs_chain(const VecRef<AL>& rr, double thresh) {
    const size_t w = rr.size();
    this->m_counter->scal += int(w);
    this->m_counter->axpy += int(w * (w - 1) / 2);
    this->m_counter->dot += int(w * (w + 1) / 2 + w);
    std::vector<double*> pr(w);
    for (size_t i = 0; i < w; ++i) {
      rr[0].get().require_compatible(rr[i].get(), "mgs_chain");
      pr[i] = rr[i].get().data();
    }
    std::vector<double> rows(w + w * (w + 1) / 2);
    check(itsolv_mgs_chain_f64(rr[0].get().context(), pr.data(), int(w), rr[0].get().local_size(), thresh, rows.data()),
          "ArrayHandlerCUDA::mgs_chain");
    if (m_observer)
      m_observer('g', 1, rows.size(), rows.data());
    return rows;
  }
  //! can the projection of `m` new vectors against `k` subspace vectors and the Gram-Schmidt chain of the `kept` ones run
  //! as one chain (rank-independent decision)
  bool project_mgs_chain_supported(size_t k, size_t m, const VecRef<AL>& kept) const {
    if (kept.empty())
      return false;
    const auto& first = kept[0].get();
    const size_t nranks = size_t(itsolv_comm_size(first.context()));
    return first.size() >= nranks && itsolv_project_mgs_chain_supported(first.context(), int(k), int(m), int(kept.size()),
                                                                         first.size() / nranks) != 0;
  }
  /*!
   * gemm_outer (with `yscale`: gemm_outer_scaled) of all new vectors yy, and the Gram-Schmidt chain of the vectors
   * yy[keep[.]] that stay, as one chain of launches: the projection's tail returns the first Gram row
   * (itsolv_project_mgs_chain_f64). Returns the rows of inner products in the layout of mgs_chain().
   */
  std::vector<double> project_mgs_chain(const Matrix<value_type>& alphas, const CVecRef<AR>& xx, const VecRef<AL>& yy,
                                        const std::vector<double>* yscale, const std::vector<int>& keep, double thresh) {
    const size_t nx = alphas.rows(), ny = alphas.cols(), w = keep.size();
    if (nx > xx.size() || ny > yy.size() || (yscale && ny > yscale->size()) || w == 0 || w > ny)
      throw std::out_of_range("project_mgs_chain: dimensions of alphas do not match xx, yy, yscale, keep");
    this->m_counter->gemm_outer++;
    this->m_counter->scal += int(w) + (yscale ? int(ny) : 0);
    this->m_counter->axpy += int(w * (w - 1) / 2);
    this->m_counter->dot += int(w * (w + 1) / 2 + w);
    std::vector<const double*> px(nx);
    std::vector<double*> py(ny);
    const AL& first = yy[0].get();
    for (size_t i = 0; i < nx; ++i) {
      first.require_compatible(xx[i].get(), "project_mgs_chain");
      px[i] = xx[i].get().data();
    }
    for (size_t j = 0; j < ny; ++j) {
      first.require_compatible(yy[j].get(), "project_mgs_chain");
      py[j] = yy[j].get().data();
    }
    std::vector<double> rows(w + w * (w + 1) / 2);
    check(itsolv_project_mgs_chain_f64(first.context(), alphas.data().data(), int(nx), int(ny), px.data(), py.data(),
                                       yscale ? yscale->data() : nullptr, keep.data(), int(w), first.local_size(), thresh,
                                       rows.data()),
          "ArrayHandlerCUDA::project_mgs_chain");
    if (m_observer)
      m_observer('g', 1, rows.size(), rows.data());
    return rows;
  }
  //! yy[j] = yscale[j] * yy[j] + sum_i alphas(i,j) xx[i]: scal_batch followed by gemm_outer, in one pass
  void gemm_outer_scaled(const Matrix<value_type>& alphas, const CVecRef<AR>& xx, const VecRef<AL>& yy,
                         const std::vector<double>& yscale) {
    if (alphas.rows() > xx.size() || alphas.cols() > yy.size() || alphas.cols() > yscale.size())
      throw std::out_of_range("gemm_outer_scaled: dimensions of alphas do not match xx, yy, yscale");
    const size_t nx = alphas.rows(), ny = alphas.cols();
    if (ny == 0)
      return;
    this->m_counter->gemm_outer++;
    this->m_counter->scal += int(ny);
    std::vector<const double*> px(nx);
    std::vector<double*> py(ny);
    const AL& first = yy[0].get();
    for (size_t i = 0; i < nx; ++i) {
      first.require_compatible(xx[i].get(), "gemm_outer_scaled");
      px[i] = xx[i].get().data();
    }
    for (size_t j = 0; j < ny; ++j) {
      first.require_compatible(yy[j].get(), "gemm_outer_scaled");
      py[j] = yy[j].get().data();
    }
    check(itsolv_gemm_outer_scaled_f64(first.context(), alphas.data().data(), int(nx), int(ny), px.data(), py.data(),
                                       first.local_size(), yscale.data()),
          "ArrayHandlerCUDA::gemm_outer_scaled");
  }
  //! yy[j] = sum_i alphas(i,j) xx[i]: the targets are written, not read (fill + gemm_outer of the reference in one pass)
  void gemm_outer_assign(const Matrix<value_type>& alphas, const CVecRef<AR>& xx, const VecRef<AL>& yy) {
    expand(alphas, xx, yy, true);
  }

  void gemm_outer(const Matrix<value_type> alphas, const CVecRef<AR>& xx, const VecRef<AL>& yy) override {
    expand(alphas, xx, yy, false);
  }

protected:
  struct Primed {
    const double *x, *y;
    double value;
  };
  std::vector<Primed> m_primed;
  unsigned long long m_primed_epoch = 0;
  bool m_take_on_copy = false;

  bool primed(const AL& x, const AR& y, double& value) {
    if (m_primed.empty())
      return false;
    if (itsolv_ctx_write_epoch(x.context()) != m_primed_epoch) {
      m_primed.clear();
      return false;
    }
    const AL& cx = x;
    for (const auto& e : m_primed)
      if ((e.x == cx.data() && e.y == y.data()) || (e.x == y.data() && e.y == cx.data())) {
        value = e.value;
        return true;
      }
    return false;
  }

  void expand(const Matrix<value_type>& alphas, const CVecRef<AR>& xx, const VecRef<AL>& yy, bool assign) {
    this->m_counter->gemm_outer++;
    // Shape rules of the reference's loops (array/util/gemm.h:186-203, 258-265): one x per row of alphas, one y per
    // column; the drivers may pass MORE y vectors than columns (construct_solution hands all R buffers with
    // roots.size() columns, itsolv/IterativeSolverTemplate.h:44-64) and only the first alphas.cols() are touched.
    if (alphas.rows() > xx.size())
      throw std::out_of_range("gemm_outer: dimensions of xx and alphas are different.");
    if (alphas.cols() > yy.size())
      throw std::out_of_range("gemm_outer: dimensions of yy and alphas are different.");
    const size_t nx = alphas.rows(), ny = alphas.cols();
    if (ny == 0)
      return;
    if (nx == 0) {
      if (assign)
        for (size_t j = 0; j < ny; ++j)
          yy[j].get().fill(0.0);
      return;
    }
    std::vector<const double*> px(nx);
    std::vector<double*> py(ny);
    const AL& first = yy[0].get();
    for (size_t i = 0; i < nx; ++i) {
      first.require_compatible(xx[i].get(), "gemm_outer");
      px[i] = xx[i].get().data();
    }
    for (size_t j = 0; j < ny; ++j) {
      first.require_compatible(yy[j].get(), "gemm_outer");
      py[j] = yy[j].get().data();
    }
    check(itsolv_gemm_outer_f64(first.context(), alphas.data().data(), int(nx), int(ny), px.data(), py.data(),
                                first.local_size(), assign ? 1 : 0),
          "ArrayHandlerCUDA::gemm_outer");
  }

public:

  Matrix<value_type> gemm_inner(const CVecRef<AL>& xx, const CVecRef<AR>& yy) override {
    this->m_counter->gemm_inner++;
    auto mat = Matrix<value_type>({xx.size(), yy.size()});
    if (xx.empty() || yy.empty()) {
      if (m_observer)
        m_observer('g', mat.rows(), mat.cols(), nullptr);
      return mat;
    }
    std::vector<const double*> px(xx.size()), py(yy.size());
    const AL& first = xx[0].get();
    for (size_t i = 0; i < xx.size(); ++i) {
      first.require_compatible(xx[i].get(), "gemm_inner");
      px[i] = xx[i].get().data();
    }
    for (size_t j = 0; j < yy.size(); ++j) {
      first.require_compatible(yy[j].get(), "gemm_inner");
      py[j] = yy[j].get().data();
    }
    std::vector<double> out(xx.size() * yy.size());
    check(itsolv_gemm_inner_f64(first.context(), px.data(), int(xx.size()), py.data(), int(yy.size()),
                                first.local_size(), out.data()),
          "ArrayHandlerCUDA::gemm_inner");
    for (size_t i = 0; i < mat.rows(); ++i)
      for (size_t j = 0; j < mat.cols(); ++j)
        mat(i, j) = out[i * mat.cols() + j];
    if (m_observer)
      m_observer('g', mat.rows(), mat.cols(), out.data());
    return mat;
  }

  std::map<size_t, value_type_abs> select_max_dot(size_t n, const AL& x, const AR& y) override {
    if (n > x.size() || n > y.size())
      error("ArrayHandlerCUDA::select_max_dot() n is too large");
    return x.select_max_dot(n, y);
  }

  std::map<size_t, value_type> select(size_t n, const AL& x, bool max = false, bool ignore_sign = false) override {
    if (n > x.size())
      error("ArrayHandlerCUDA::select() n is too large");
    return x.select(n, max, ignore_sign);
  }

  ProxyHandle lazy_handle() override { return this->lazy_handle(*this); }

protected:
  using Base::error;
  using Base::lazy_handle;
  ResultObserver m_observer;
};

//! dense (device) x sparse (std::map on the host): the rp / qp handlers of the P-space
class ArrayHandlerCUDASparse : public ArrayHandler<DistrArrayCUDA, std::map<size_t, double>> {
public:
  using AL = DistrArrayCUDA;
  using AR = std::map<size_t, double>;
  using Base = ArrayHandler<AL, AR>;
  using typename Base::ProxyHandle;
  using typename Base::value_type;
  using typename Base::value_type_abs;

  ArrayHandlerCUDASparse() = default;
  explicit ArrayHandlerCUDASparse(ResultObserver observer) : m_observer(std::move(observer)) {}

  AL copy(const AR&) override { throw std::logic_error("General construction of dense from sparse is ill-defined"); }

  //! x = 0, then x[index] = value for the entries this rank owns (reference ArrayHandlerDistrSparse.h:30-38)
  void copy(AL& x, const AR& y) override {
    Packed p({std::cref(y)});
    check(itsolv_sparse_copy_f64(x.context(), x.data(), x.local_size(), x.local_start(), int(p.idx.size()), p.idx.data(),
                                 p.val.data()),
          "ArrayHandlerCUDASparse::copy");
  }
  void scal(value_type, AL&) override {} // unary operations belong to the dense handler, as in the reference
  void fill(value_type, AL&) override {}

  void axpy(value_type alpha, const AR& x, AL& y) override {
    this->m_counter->axpy++;
    Packed p({std::cref(x)});
    double* py = y.data();
    check(itsolv_sparse_gemm_outer_f64(y.context(), &alpha, 1, 1, p.ptr.data(), p.idx.data(), p.val.data(), &py,
                                       y.local_size(), y.local_start()),
          "ArrayHandlerCUDASparse::axpy");
  }

  value_type dot(const AL& x, const AR& y) override {
    this->m_counter->dot++;
    Packed p({std::cref(y)});
    const double* px = x.data();
    double d = 0;
    check(itsolv_sparse_gemm_inner_f64(x.context(), &px, 1, x.local_size(), x.local_start(), 1, p.ptr.data(),
                                       p.idx.data(), p.val.data(), &d),
          "ArrayHandlerCUDASparse::dot");
    if (m_observer)
      m_observer('d', 1, 1, &d);
    return d;
  }

  //! alphas: rows <-> sparse xx, columns <-> dense yy (reference array/util/gemm.h:207-224)
  void gemm_outer(const Matrix<value_type> alphas, const CVecRef<AR>& xx, const VecRef<AL>& yy) override {
    this->m_counter->gemm_outer++;
    if (alphas.rows() > xx.size() || alphas.cols() > yy.size())
      throw std::out_of_range("gemm_outer: dimensions of alphas do not match xx, yy");
    const size_t nx = alphas.rows(), ny = alphas.cols(); // more y vectors than columns is allowed, see ArrayHandlerCUDA
    if (nx == 0 || ny == 0)
      return;
    Packed p(CVecRef<AR>(xx.begin(), xx.begin() + nx));
    std::vector<double*> py(ny);
    const AL& first = yy[0].get();
    for (size_t j = 0; j < ny; ++j) {
      first.require_compatible(yy[j].get(), "gemm_outer");
      py[j] = yy[j].get().data();
    }
    check(itsolv_sparse_gemm_outer_f64(first.context(), alphas.data().data(), int(nx), int(ny),
                                       p.ptr.data(), p.idx.data(), p.val.data(), py.data(), first.local_size(),
                                       first.local_start()),
          "ArrayHandlerCUDASparse::gemm_outer");
  }

  Matrix<value_type> gemm_inner(const CVecRef<AL>& xx, const CVecRef<AR>& yy) override {
    this->m_counter->gemm_inner++;
    auto mat = Matrix<value_type>({xx.size(), yy.size()});
    if (xx.empty() || yy.empty()) {
      if (m_observer)
        m_observer('g', mat.rows(), mat.cols(), nullptr);
      return mat;
    }
    Packed p(yy);
    std::vector<const double*> px(xx.size());
    const AL& first = xx[0].get();
    for (size_t i = 0; i < xx.size(); ++i) {
      first.require_compatible(xx[i].get(), "gemm_inner");
      px[i] = xx[i].get().data();
    }
    std::vector<double> out(xx.size() * yy.size());
    check(itsolv_sparse_gemm_inner_f64(first.context(), px.data(), int(xx.size()), first.local_size(),
                                       first.local_start(), int(yy.size()), p.ptr.data(), p.idx.data(), p.val.data(),
                                       out.data()),
          "ArrayHandlerCUDASparse::gemm_inner");
    for (size_t i = 0; i < mat.rows(); ++i)
      for (size_t j = 0; j < mat.cols(); ++j)
        mat(i, j) = out[i * mat.cols() + j];
    if (m_observer)
      m_observer('g', mat.rows(), mat.cols(), out.data());
    return mat;
  }

  //! as ArrayHandlerDistrSparse (reference array/ArrayHandlerDistrSparse.h:65-67): the container's own member
  std::map<size_t, value_type_abs> select_max_dot(size_t n, const AL& x, const AR& y) override {
    if (n > x.size() || n > y.size())
      error("ArrayHandlerCUDASparse::select_max_dot() n is too large");
    return x.select_max_dot(n, y);
  }

  std::map<size_t, value_type> select(size_t n, const AL& x, bool max = false, bool ignore_sign = false) override {
    if (n > x.size())
      error("ArrayHandlerCUDASparse::select() n is too large");
    return x.select(n, max, ignore_sign);
  }

  ProxyHandle lazy_handle() override { return this->lazy_handle(*this); }

protected:
  using Base::error;
  using Base::lazy_handle;
  ResultObserver m_observer;

  //! std::map vectors packed CSR-like for the C ABI
  struct Packed {
    std::vector<int32_t> ptr;
    std::vector<int64_t> idx;
    std::vector<double> val;
    explicit Packed(const CVecRef<AR>& maps) {
      ptr.push_back(0);
      for (const auto& m : maps) {
        for (const auto& e : m.get()) {
          idx.push_back(int64_t(e.first));
          val.push_back(e.second);
        }
        ptr.push_back(int32_t(idx.size()));
      }
    }
  };
};

using HandlersCUDA = molpro::linalg::itsolv::ArrayHandlers<DistrArrayCUDA, DistrArrayCUDA, std::map<size_t, double>>;

//! The seven handlers of a solver whose R and Q containers are DistrArrayCUDA and whose P container is std::map
inline std::shared_ptr<HandlersCUDA> make_handlers(ResultObserver observer = nullptr) {
  using P = std::map<size_t, double>;
  auto dense = std::make_shared<ArrayHandlerCUDA>(observer);
  auto sparse = std::make_shared<ArrayHandlerCUDASparse>(observer);
  auto pp = std::make_shared<molpro::linalg::array::ArrayHandlerSparse<P, P>>();
  return std::make_shared<HandlersCUDA>(dense, dense, pp, dense, sparse, dense, sparse);
}

} // namespace itsolv_b200
#endif
