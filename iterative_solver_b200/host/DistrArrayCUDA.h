// DistrArrayCUDA: row-sharded FP64 vector resident in the HBM of the calling rank's B200.
//
// It is the R/Q container handed to the reference's solver templates
// (LinearEigensystemDavidson<DistrArrayCUDA, DistrArrayCUDA, std::map<size_t,double>>, ...). The solvers never touch
// elements; they need value_type, copy/move construction and move assignment only
// (reference src/molpro/linalg/itsolv/IterativeSolverTemplate.h:431-442, subspace/DSpace.h:41-45),
// and every O(n) operation goes through ArrayHandlerCUDA. The sharding follows the reference's DistrArray:
// contiguous chunks from util::make_distribution_spread_remainder (reference array/util/Distribution.h:99-110), one
// chunk per rank of the context's communicator. The member functions mirror the names of the reference's DistrArray
// (reference array/DistrArray.h:90-300) where the operation makes sense for device memory.
#ifndef ITSOLV_B200_HOST_DISTRARRAYCUDA_H
#define ITSOLV_B200_HOST_DISTRARRAYCUDA_H
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include <itsolv_b200.h>

namespace itsolv_b200 {

//! Error raised by the CUDA backend; message comes from itsolv_last_error()
struct CudaBackendError : public std::runtime_error {
  using std::runtime_error::runtime_error;
};

inline void check(int rc, const char* where) {
  if (rc != 0)
    throw CudaBackendError(std::string(where) + ": " + itsolv_last_error());
}

class DistrArrayCUDA {
public:
  using value_type = double;
  using index_type = size_t;

  DistrArrayCUDA() = default;

  //! Allocates this rank's shard of a vector of global length `dimension`; contents are unspecified
  DistrArrayCUDA(size_t dimension, itsolv_ctx* ctx) : m_ctx(ctx), m_dimension(dimension) {
    const int nranks = itsolv_comm_size(ctx), rank = itsolv_comm_rank(ctx);
    std::vector<int64_t> borders(size_t(nranks) + 1);
    itsolv_distribution(dimension, nranks, borders.data());
    m_start = size_t(borders[rank]);
    m_local = size_t(borders[rank + 1] - borders[rank]);
    check(itsolv_alloc(m_ctx, m_local, &m_data), "DistrArrayCUDA: allocation");
  }

  /*!
   * Non-owning view of caller-owned DEVICE memory as this rank's shard of a vector of global length `dimension`
   * (`local_size()` doubles at `data`): what DistrArraySpan is for host buffers in the reference's C interface
   * (reference array/DistrArraySpan.h:8-48, IterativeSolverCMPI.cpp:89-107). Copies of a view own their memory.
   */
  static DistrArrayCUDA view(size_t dimension, itsolv_ctx* ctx, double* data) {
    DistrArrayCUDA v;
    v.m_ctx = ctx;
    v.m_dimension = dimension;
    const int nranks = itsolv_comm_size(ctx), rank = itsolv_comm_rank(ctx);
    std::vector<int64_t> borders(size_t(nranks) + 1);
    itsolv_distribution(dimension, nranks, borders.data());
    v.m_start = size_t(borders[rank]);
    v.m_local = size_t(borders[rank + 1] - borders[rank]);
    v.m_data = data;
    v.m_owner = false;
    return v;
  }
  bool owns() const { return m_owner; }

  DistrArrayCUDA(const DistrArrayCUDA& source)
      : m_ctx(source.m_ctx), m_dimension(source.m_dimension), m_start(source.m_start), m_local(source.m_local) {
    if (source.m_data) {
      check(itsolv_alloc(m_ctx, m_local, &m_data), "DistrArrayCUDA: allocation");
      check(itsolv_copy_f64(m_ctx, m_data, source.m_data, m_local), "DistrArrayCUDA: copy");
    }
  }

  DistrArrayCUDA(DistrArrayCUDA&& source) noexcept { swap(source); }

  DistrArrayCUDA& operator=(DistrArrayCUDA&& source) noexcept {
    if (this != &source) {
      release();
      swap(source);
    }
    return *this;
  }

  DistrArrayCUDA& operator=(const DistrArrayCUDA& source) {
    if (this == &source)
      return *this;
    if (!m_data || !compatible(source)) {
      DistrArrayCUDA t(source);
      release();
      swap(t);
    } else {
      check(itsolv_copy_f64(m_ctx, m_data, source.m_data, m_local), "DistrArrayCUDA: copy");
    }
    return *this;
  }

  ~DistrArrayCUDA() { release(); }

  //! global length, as DistrArray::size() (reference array/DistrArray.h:112)
  size_t size() const { return m_dimension; }
  bool empty() const { return m_data == nullptr; }
  //! this rank's rows [start, start + local_size)
  size_t local_size() const { return m_local; }
  size_t local_start() const { return m_start; }
  std::pair<size_t, size_t> local_range() const { return {m_start, m_start + m_local}; }
  //! mutable access: assume the caller writes (invalidates results cached by the handlers)
  double* data() {
    if (m_ctx)
      itsolv_ctx_note_write(m_ctx);
    return m_data;
  }
  const double* data() const { return m_data; }
  itsolv_ctx* context() const { return m_ctx; }

  //! same global length and same sharding (reference DistrArray::compatible, array/DistrArray.h:117-120)
  bool compatible(const DistrArrayCUDA& other) const {
    return m_ctx == other.m_ctx && m_dimension == other.m_dimension && m_local == other.m_local && m_start == other.m_start;
  }

  // ---- device operations named after the reference's DistrArray members (array/DistrArray.cpp:43-138) ----
  void fill(double a) { check(itsolv_fill_f64(m_ctx, a, m_data, m_local), "DistrArrayCUDA::fill"); }
  void scal(double a) { check(itsolv_scal_f64(m_ctx, a, m_data, m_local), "DistrArrayCUDA::scal"); }
  void copy(const DistrArrayCUDA& y) {
    require_compatible(y, "copy");
    check(itsolv_copy_f64(m_ctx, m_data, y.m_data, m_local), "DistrArrayCUDA::copy");
  }
  void axpy(double a, const DistrArrayCUDA& x) {
    require_compatible(x, "axpy");
    check(itsolv_axpy_f64(m_ctx, a, x.m_data, m_data, m_local), "DistrArrayCUDA::axpy");
  }
  double dot(const DistrArrayCUDA& y) const {
    require_compatible(y, "dot");
    double r = 0;
    check(itsolv_dot_f64(m_ctx, m_data, y.m_data, m_local, &r), "DistrArrayCUDA::dot");
    return r;
  }
  // the remaining element-wise members of the reference's DistrArray (array/DistrArray.cpp:79-167); not used by the
  // solvers, present so that the reference's conformance tests (test/array/testDistrArray.h:496-678) apply
  void add(const DistrArrayCUDA& y) { axpy(1, y); }
  void sub(const DistrArrayCUDA& y) { axpy(-1, y); }
  void add(double a) { elementwise(ITSOLV_EW_ADD_SCALAR, nullptr, nullptr, a, "add"); }
  void sub(double a) { add(-a); }
  void recip() { elementwise(ITSOLV_EW_RECIP, nullptr, nullptr, 0.0, "recip"); }
  void times(const DistrArrayCUDA& y) { elementwise(ITSOLV_EW_TIMES_INPLACE, &y, nullptr, 0.0, "times"); }
  void times(const DistrArrayCUDA& y, const DistrArrayCUDA& z) { elementwise(ITSOLV_EW_TIMES, &y, &z, 0.0, "times"); }
  //! this[i] (=|+=|-=) (-)y[i] / (z[i] + shift)   (reference DistrArray::_divide, array/DistrArray.cpp:140-167)
  void divide(const DistrArrayCUDA& y, const DistrArrayCUDA& z, double shift = 0, bool append = false,
              bool negative = false) {
    const int op = append ? (negative ? ITSOLV_EW_DIVIDE_APPEND_NEGATIVE : ITSOLV_EW_DIVIDE_APPEND)
                          : (negative ? ITSOLV_EW_DIVIDE_NEGATIVE : ITSOLV_EW_DIVIDE);
    elementwise(op, &y, &z, shift, "divide");
  }
  void zero() { fill(0.0); }
  //! this[index] += a * value for the entries of the sparse array that this rank owns (reference DistrArray.cpp:419-437)
  void axpy(double a, const std::map<size_t, double>& y) {
    if (y.empty())
      return;
    std::vector<int32_t> ptr{0, int32_t(y.size())};
    std::vector<int64_t> idx;
    std::vector<double> val;
    for (const auto& e : y) {
      idx.push_back(int64_t(e.first));
      val.push_back(e.second);
    }
    double* py = data();
    check(itsolv_sparse_gemm_outer_f64(m_ctx, &a, 1, 1, ptr.data(), idx.data(), val.data(), &py, m_local, m_start),
          "DistrArrayCUDA::axpy(sparse)");
  }
  //! sum over the entries of the sparse array of this[index] * value (reference DistrArray.cpp:439-465)
  double dot(const std::map<size_t, double>& y) const {
    if (y.empty())
      return 0.0;
    std::vector<int32_t> ptr{0, int32_t(y.size())};
    std::vector<int64_t> idx;
    std::vector<double> val;
    for (const auto& e : y) {
      idx.push_back(int64_t(e.first));
      val.push_back(e.second);
    }
    const double* px = m_data;
    double d = 0;
    check(itsolv_sparse_gemm_inner_f64(m_ctx, &px, 1, m_local, m_start, 1, ptr.data(), idx.data(), val.data(), &d),
          "DistrArrayCUDA::dot(sparse)");
    return d;
  }
  //! the elements at the given global indices, on every rank (each entry is owned by one rank; the others add zero)
  std::vector<double> gather(const std::vector<size_t>& indices) const {
    std::vector<double> out(indices.size(), 0.0);
    if (indices.empty())
      return out;
    std::vector<int32_t> ptr(indices.size() + 1);
    std::vector<int64_t> idx(indices.size());
    std::vector<double> one(indices.size(), 1.0);
    for (size_t e = 0; e < indices.size(); ++e) {
      ptr[e] = int32_t(e);
      idx[e] = int64_t(indices[e]);
    }
    ptr[indices.size()] = int32_t(indices.size());
    const double* px = m_data;
    check(itsolv_sparse_gemm_inner_f64(m_ctx, &px, 1, m_local, m_start, int(indices.size()), ptr.data(), idx.data(),
                                       one.data(), out.data()),
          "DistrArrayCUDA::gather");
    return out;
  }
  /*!
   * The n entries of the sparse array with the largest |this[index] * value| (reference DistrArray.cpp:248-262 with
   * util::select_max_dot_iter_sparse, array/util/select_max_dot.h:60-83: a min-heap of (value, index) pairs keeps the n
   * largest pairs in lexicographic order).
   */
  std::map<size_t, double> select_max_dot(size_t n, const std::map<size_t, double>& y) const {
    if (!y.empty() && size() < y.rbegin()->first + 1)
      throw std::runtime_error("DistrArrayCUDA::select_max_dot: sparse array x is too large");
    if (n > size() || n > y.size())
      throw std::runtime_error("DistrArrayCUDA::select_max_dot: n is too large");
    std::vector<size_t> indices;
    for (const auto& e : y)
      indices.push_back(e.first);
    const auto x = gather(indices);
    std::vector<std::pair<double, size_t>> pairs;
    size_t e = 0;
    for (const auto& item : y) {
      pairs.emplace_back(std::abs(x[e] * item.second), item.first);
      ++e;
    }
    std::sort(pairs.begin(), pairs.end(), std::greater<std::pair<double, size_t>>());
    std::map<size_t, double> result;
    for (size_t i = 0; i < n && i < pairs.size(); ++i)
      result.emplace(pairs[i].second, pairs[i].first);
    return result;
  }
  std::map<size_t, double> select(size_t n, bool max = false, bool ignore_sign = false) const {
    return select_impl(n, nullptr, max, ignore_sign);
  }
  std::map<size_t, double> select_max_dot(size_t n, const DistrArrayCUDA& y) const {
    require_compatible(y, "select_max_dot");
    return select_impl(n, y.m_data, true, false);
  }

  // ---- host transfers of the local shard (tests, harness) ----
  void upload(const double* host) { check(itsolv_upload(m_ctx, m_data, host, m_local), "DistrArrayCUDA::upload"); }
  void download(double* host) const { check(itsolv_download(m_ctx, host, m_data, m_local), "DistrArrayCUDA::download"); }

  void swap(DistrArrayCUDA& o) noexcept {
    std::swap(m_ctx, o.m_ctx);
    std::swap(m_dimension, o.m_dimension);
    std::swap(m_start, o.m_start);
    std::swap(m_local, o.m_local);
    std::swap(m_data, o.m_data);
    std::swap(m_owner, o.m_owner);
  }

  //! Hands this vector's contents to a new array without moving a byte: the new array owns the old allocation, this one
  //! continues with a fresh allocation of the same shape whose contents are unspecified. For callers that are about to
  //! overwrite this vector anyway (the fused driver path: R vectors entering the Q space).
  DistrArrayCUDA take_contents() {
    if (!m_owner) // a view cannot give its memory away
      return DistrArrayCUDA(*this);
    DistrArrayCUDA fresh(m_dimension, m_ctx);
    std::swap(m_data, fresh.m_data);
    itsolv_ctx_note_write(m_ctx);
    return fresh;
  }

  //! Frees the device memory now and leaves an empty() vector of the same shape: for owners that know the contents are
  //! dead before the object itself is destroyed (the fused D-space construction, memory-capped runs)
  void release_storage() noexcept {
    if (m_data && m_ctx && m_owner) {
      itsolv_free(m_ctx, m_data);
      m_data = nullptr;
    }
  }

  void require_compatible(const DistrArrayCUDA& o, const char* op) const {
    if (!m_data || !o.m_data || !compatible(o))
      throw std::runtime_error(std::string("DistrArrayCUDA::") + op + ": incompatible arrays");
  }

private:
  void elementwise(int op, const DistrArrayCUDA* a, const DistrArrayCUDA* b, double scalar, const char* name) {
    if (a)
      require_compatible(*a, name);
    if (b)
      require_compatible(*b, name);
    check(itsolv_elementwise_f64(m_ctx, op, data(), a ? a->m_data : nullptr, b ? b->m_data : nullptr, scalar, m_local),
          "DistrArrayCUDA: element-wise operation");
  }
  std::map<size_t, double> select_impl(size_t n, const double* y, bool max, bool ignore_sign) const {
    if (n > m_dimension)
      throw std::runtime_error("DistrArrayCUDA::select: n is too large");
    std::vector<int64_t> idx(n);
    std::vector<double> val(n);
    int found = 0;
    check(itsolv_select_f64(m_ctx, m_data, y, m_local, m_start, n, max ? 1 : 0, ignore_sign ? 1 : 0, idx.data(),
                            val.data(), &found),
          "DistrArrayCUDA::select");
    std::map<size_t, double> result;
    for (int i = 0; i < found; ++i)
      result.emplace(size_t(idx[i]), val[i]);
    return result;
  }

  void release() noexcept {
    if (m_data && m_ctx && m_owner)
      itsolv_free(m_ctx, m_data);
    m_data = nullptr;
    m_owner = true;
  }

  itsolv_ctx* m_ctx = nullptr;
  size_t m_dimension = 0;
  size_t m_start = 0;
  size_t m_local = 0;
  double* m_data = nullptr;
  bool m_owner = true;
};

} // namespace itsolv_b200
#endif
