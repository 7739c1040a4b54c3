"""iterative_solver_b200: a B200 (sm_100a) vector backend for molpro::linalg::itsolv.

Native code lives in lib/libitsolv_b200.so (CUDA kernels, C ABI include/itsolv_b200.h) and lib/libitsolv_b200_host.so
(DistrArrayCUDA / ArrayHandlerCUDA behind the reference's ArrayHandler contract, and the solve harness). This Python
package only binds them for tests and benchmarks."""
from . import _native
from .api import BackendError, Context, distribution, select_merge

__all__ = ["BackendError", "Context", "distribution", "select_merge", "_native"]
