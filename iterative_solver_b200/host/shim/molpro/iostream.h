#ifndef ITSOLV_B200_SHIM_MOLPRO_IOSTREAM_H
#define ITSOLV_B200_SHIM_MOLPRO_IOSTREAM_H
#include <iostream>
namespace molpro {
using std::cout;
using std::cerr;
} // namespace molpro
#endif
