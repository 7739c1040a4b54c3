// Stand-in for molpro::Options from the un-vendored molpro utilities 0.5.5. Always returns the default.
#ifndef ITSOLV_B200_SHIM_MOLPRO_OPTIONS_H
#define ITSOLV_B200_SHIM_MOLPRO_OPTIONS_H
#include <string>
namespace molpro {
class Options {
public:
  Options(std::string program = "", std::string options = "") : m_program(std::move(program)) { (void)options; }
  int parameter(const std::string&, int def) const { return def; }
  double parameter(const std::string&, double def) const { return def; }
  std::string parameter(const std::string&, const std::string& def) const { return def; }
  std::string parameter(const std::string&, const char* def) const { return def; }

private:
  std::string m_program;
};
} // namespace molpro
#endif
