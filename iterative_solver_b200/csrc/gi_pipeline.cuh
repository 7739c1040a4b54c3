// Shared pieces of the shared-memory-pipeline gemm_inner kernels (gemm_inner.cu: FP64 FMA register tiles;
// gemm_inner_mma.cu: FP64 tensor-core tiles): kernel parameters, mbarrier / TMA / cp.async wrappers.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "common.cuh"
#include "gi_finalize.cuh"

namespace itsolv {

constexpr int kMaxStages = 8;
constexpr int kMaxProducerWarps = 8;

struct GiParams {
  const double* vec[2 * ITSOLV_MAX_PANEL]; // distinct vectors of the call
  unsigned char xslot[ITSOLV_MAX_PANEL];   // xx[i] -> index into vec
  unsigned char yslot[ITSOLV_MAX_PANEL];   // yy[j] -> index into vec
  GiFinalize fin;                          // per-CTA partial sums and how they become the final result
  size_t n;
  long long nfull; // number of full tiles of `rows` rows
  int nvec, k, m;
  int rows;   // T: rows per tile (multiple of 2)
  int stride; // doubles between consecutive vectors inside a stage (T + 2: shifts each vector by one 16-byte bank group)
  int stages;
  int KB, MB, G; // thread-tile grid (KB x MB tiles) and number of row groups
  int chunk_rows; // rows per TMA copy: a vector's tile travels as ceil(rows / chunk_rows) copies (many small copies in
                  // flight stream faster than a few large ones)
  int nprod;     // producer warps (each issues the TMA copies of the vectors v == warp (mod nprod))
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n"
               ".reg .pred P1;\n"
               "LAB_WAIT:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
               "@P1 bra DONE;\n"
               "bra LAB_WAIT;\n"
               "DONE:\n"
               "}" ::"r"(smem_u32(bar)),
               "r"(parity)
               : "memory");
}
//! 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src_gmem), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n"
               ".reg .pred P;\n"
               "elect.sync _|P, 0xffffffff;\n"
               "selp.u32 %0, 1, 0, P;\n"
               "}"
               : "=r"(pred));
  return pred != 0;
}

// How a tile reaches shared memory
enum Loader {
  LOAD_TMA = 0,    // one elected lane per producer warp issues 1-D TMA bulk copies (UBLKCP), one per vector
  LOAD_CPASYNC16 = 1, // all producer lanes issue 16-byte cp.async (LDGSTS) pieces; completion counted on the mbarrier
  LOAD_CPASYNC8 = 2   // the same with 8-byte pieces: vectors that are only 8-byte aligned
};

template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(uint32_t dst_smem, const void* src_gmem, uint32_t src_bytes) {
  if constexpr (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src_gmem), "r"(src_bytes));
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst_smem), "l"(src_gmem), "r"(src_bytes));
}
//! the calling thread arrives on the mbarrier once all of its earlier cp.async copies have landed
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

/*!
 * TMA producer warps of a CTA (threads nconsumers .. blockDim.x): stream the CTA's tiles (tile s of the CTA is global
 * tile blockIdx.x + s*gridDim.x) into the stage ring. Control flow and addresses are warp-uniform (everything derives
 * from kernel parameters and loop counters), so the copies are issued from the uniform datapath; one elected lane per
 * warp executes the arrive.expect_tx and the UBLKCPs of the vectors v == warp (mod nprod).
 */
__device__ __forceinline__ void gi_tma_producer(const GiParams& p, double* tiles, uint64_t* full_bar, uint64_t* empty_bar,
                                                int nconsumers, long long my_tiles, size_t stage_doubles) {
  const int tid = threadIdx.x;
  const int pw = (tid - nconsumers) >> 5;
  const bool leader = elect_one();
  const uint32_t vec_bytes = uint32_t(p.rows) * 8u;
  const int my_nvec = (p.nvec - pw + p.nprod - 1) / p.nprod;
  const uint32_t tiles_u32 = smem_u32(tiles);
  for (long long s = 0; s < my_tiles; ++s) {
    const int stage = int(s % p.stages);
    const long long use = s / p.stages;
    if (use > 0)
      mbar_wait(&empty_bar[stage], uint32_t((use - 1) & 1));
    const uint32_t bar = smem_u32(&full_bar[stage]);
    if (leader)
      mbar_expect_tx(&full_bar[stage], vec_bytes * uint32_t(my_nvec));
    const size_t row0 = size_t(blockIdx.x + s * gridDim.x) * size_t(p.rows);
    const uint32_t st = tiles_u32 + uint32_t(size_t(stage) * stage_doubles * 8);
    for (int v = pw; v < p.nvec; v += p.nprod) {
      const double* src = p.vec[v] + row0;
      const uint32_t dst = st + uint32_t(v) * uint32_t(p.stride) * 8u;
#pragma unroll 4
      for (int r0 = 0; r0 < p.rows; r0 += p.chunk_rows) {
        const int nr = p.rows - r0 < p.chunk_rows ? p.rows - r0 : p.chunk_rows;
        if (leader)
          bulk_load(dst + uint32_t(r0) * 8u, src + r0, uint32_t(nr) * 8u, bar);
      }
    }
  }
}

//! all threads of the CTA copy rows [row0, row0+nrows) of every vector into stage 0, zero-filling up to `rows`
__device__ __forceinline__ void cooperative_fill(const GiParams& p, double* st, size_t row0, int nrows) {
  for (int v = 0; v < p.nvec; ++v) {
    const double* __restrict__ src = p.vec[v] + row0;
    double* dst = st + size_t(v) * p.stride;
    for (int r = threadIdx.x; r < p.rows; r += blockDim.x)
      dst[r] = r < nrows ? src[r] : 0.0;
  }
}

} // namespace itsolv
