// Stand-in for molpro/mpi.h (molpro utilities 0.5.5) in a build without MPI: one rank, no communicator.
// HAVE_MPI_H is deliberately NOT defined, so every real MPI call in the reference is compiled out.
#ifndef ITSOLV_B200_SHIM_MOLPRO_MPI_H
#define ITSOLV_B200_SHIM_MOLPRO_MPI_H
#include <cstdint>
using MPI_Comm = int;
#ifndef MPI_COMM_NULL
#define MPI_COMM_NULL 0
#endif
inline int MPI_Barrier(MPI_Comm) { return 0; }
namespace molpro {
namespace mpi {
inline MPI_Comm comm_global() { return 1; }
inline MPI_Comm comm_self() { return 2; }
inline int rank_global() { return 0; }
inline int size_global() { return 1; }
inline int init() { return 0; }
inline int finalize() { return 0; }
} // namespace mpi
} // namespace molpro
#endif
