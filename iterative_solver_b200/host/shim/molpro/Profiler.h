// Stand-in for the un-vendored molpro profiler 0.5.4 (dependencies/profiler_SHA1 of the reference).
// Only the surface the iterative-solver headers use; every call is a no-op.
#ifndef ITSOLV_B200_SHIM_MOLPRO_PROFILER_H
#define ITSOLV_B200_SHIM_MOLPRO_PROFILER_H
#include <cstddef>
#include <memory>
#include <ostream>
#include <string>

namespace molpro {
namespace profiler {
class Profiler {
public:
  struct Proxy {
    Proxy& operator+=(size_t) { return *this; }
    Proxy& operator++() { return *this; }
  };
  explicit Profiler(std::string name = "") : m_name(std::move(name)) {}
  static std::shared_ptr<Profiler> single(const std::string& name = "") {
    static std::shared_ptr<Profiler> instance = std::make_shared<Profiler>("ITSOLV");
    (void)name;
    return instance;
  }
  Profiler& start(const std::string&) { return *this; }
  Profiler& stop(const std::string& = "") { return *this; }
  Proxy push(const std::string&) { return {}; }
  Profiler& reset(const std::string&) { return *this; }
  int get_max_depth() const { return m_max_depth; }
  void set_max_depth(int d) { m_max_depth = d; }
  void dotgraph(const std::string&, double = 0.01) {}
  Profiler& operator+=(size_t) { return *this; }
  std::string str() const { return "Profiler(" + m_name + ") [shim]"; }

private:
  std::string m_name;
  int m_max_depth = 0;
};
inline std::ostream& operator<<(std::ostream& os, const Profiler& p) { return os << p.str(); }
} // namespace profiler
using Profiler = profiler::Profiler;
} // namespace molpro
#endif
