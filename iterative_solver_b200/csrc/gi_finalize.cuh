// Shared tail of the gemm_inner kernels: the last CTA to publish its partial k x m sums adds all of them in CTA order
// (deterministic: the order does not depend on which CTA happens to be last) and hands the result to the host,
// either in device memory or straight into mapped pinned memory followed by a sequence word the host spins on.
#pragma once
#include <cuda_runtime.h>

namespace itsolv {

struct GiFinalize {
  double* partials;         // [gridDim.x][km]
  double* out;              // km final sums (device or mapped host memory)
  unsigned int* counter;    // CTAs that have published (reset by the last one)
  unsigned long long* flag; // mapped host word that receives `seq` once `out` is complete (or null)
  unsigned long long seq;
  int fused;                // 0: a separate kernel reduces the partials
};

//! call by ALL threads of the CTA after the CTA's partial sums were written to f.partials + blockIdx.x * km
__device__ __forceinline__ void gi_finalize(const GiFinalize& f, int km, int* s_is_last) {
  if (!f.fused)
    return;
  const int tid = threadIdx.x;
  __threadfence();
  __syncthreads();
  if (tid == 0)
    *s_is_last = (atomicAdd(f.counter, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!*s_is_last)
    return;
  __threadfence();
  const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  for (int e = warp; e < km; e += nwarps) {
    double sum = 0.0;
    for (int c = lane; c < int(gridDim.x); c += 32)
      sum += __ldcg(f.partials + size_t(c) * km + e);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
      sum += __shfl_down_sync(0xffffffffu, sum, off);
    if (lane == 0)
      f.out[e] = sum;
  }
  __syncthreads();
  if (tid == 0) {
    *f.counter = 0u;
    if (f.flag) {
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long*>(f.flag) = f.seq;
    }
  }
}

} // namespace itsolv
