// Fused "solution -> residual -> error norm -> preconditioner" step of one Davidson iteration.
//
// The reference performs, per iteration and for all m roots (IterativeSolverTemplate.h:518-563):
//   construct_solution(parameters)   x_j  = sum_i C(i,j) q_i        fill + gemm_outer      (IterativeSolverTemplate.h:34-65)
//   construct_solution(residual)     r_j  = sum_i C(i,j) a_i        fill + gemm_outer
//   construct_residual               r_j -= lambda_j x_j            m axpy                 (LinearEigensystemDavidson.h:186-192)
//   update_errors                    e_j  = sqrt(<r_j, r_j>)        m dot                  (IterativeSolverTemplate.h:96-102)
//   precondition_default             r_j /= (d - shift_j + 1e-15)   m transforms           (IterativeSolver.h:46-55)
// which moves 8n(2k + 8m + 1) bytes even when each of those steps runs once at the roofline. None of the intermediate
// vectors is needed outside this chain while the iteration goes on (solve() overwrites parameters[0] with the diagonal
// and the proposal step overwrites all of them, IterativeSolverTemplate.h:383-390), so one pass that keeps x_j and r_j in
// registers does the same arithmetic with 8n(2k + m + 1) bytes: every q_i and a_i is read once, the preconditioned
// residuals are written once.
//
// The arithmetic is that of the separate kernels, operation by operation: the two expansions are FMA chains in ascending
// i starting from zero (gemm_outer.cu, beta_zero), the residual is r + round(-lambda*x) with product and sum rounded
// separately (axpy, blas1.cu), the preconditioner is r / ((d - shift) + 1e-15) with every operation rounded
// (precondition_kernel, blas1.cu). The vectors it writes are therefore bit-identical to the unfused sequence; the norms
// are summed in this kernel's own (deterministic) order.
#include <algorithm>

#include "common.cuh"
#include "gi_finalize.cuh"
#include "gi_pipeline.cuh"

namespace itsolv {

void fill_finalize(itsolv_ctx* ctx, int grid_bound, int km, GiFinalize* f, bool* host_direct); // gemm_inner.cu
int launch_reduce_partials(itsolv_ctx* ctx, int grid, int km);                             // gemm_inner.cu
int finish_with_peers(itsolv_ctx* ctx, int km, bool* host_direct);                         // gemm_inner.cu
int finish_result(itsolv_ctx* ctx, int count, double* out, bool host_direct);              // gemm_inner.cu

constexpr int kDsThreads = 256;
constexpr int kDsMaxRoots = 16; // roots per launch (running sums live in registers)

struct DsParams {
  const double* q[ITSOLV_MAX_PANEL]; // parameter vectors of the subspace (Q then D)
  const double* a[ITSOLV_MAX_PANEL]; // their actions
  double* out_x[kDsMaxRoots];        // solutions, or null entries when they are not wanted
  double* out_r[kDsMaxRoots];        // residuals (preconditioned when diag != null)
  double lambda[kDsMaxRoots];
  double shift[kDsMaxRoots];
  const double* coef; // device, k x ld row-major, zero padded columns
  const double* diag; // null: residuals are written as they are
  // linear equations (mode 1): r_j = (sum_i c_ij a_i - rhs_j) * rscale_j instead of r_j = sum_i c_ij a_i - lambda_j x_j
  const double* rhs[kDsMaxRoots];
  double rscale[kDsMaxRoots];
  GiFinalize fin;
  size_t n;
  int k, m, ld;
  int write_x;
  int mode;       // 0: eigenproblem residual, 1: linear equations residual, 2: r_j = sum_i c_ij a_i as it is (DIIS)
  int update_x;   // the stored x_j is x_j - out_r_j (the DIIS step with the preconditioned residual, mode 3 of the entry)
  int accumulate; // the expansions start from the present contents of out_x / out_r (P-space parts) instead of zero
};

template <class RV>
struct DsOps;
template <>
struct DsOps<double2> {
  static __device__ __forceinline__ double2 zero() { return make_double2(0.0, 0.0); }
  static __device__ __forceinline__ void fma_to(double2& acc, double c, const double2& v) {
    acc.x = fma(c, v.x, acc.x);
    acc.y = fma(c, v.y, acc.y);
  }
  static __device__ __forceinline__ double2 residual(const double2& r, double neg_lambda, const double2& x) {
    return make_double2(__dadd_rn(r.x, __dmul_rn(neg_lambda, x.x)), __dadd_rn(r.y, __dmul_rn(neg_lambda, x.y)));
  }
  //! (r - b) * s, every operation rounded: axpy(-1, rhs, r) then scal(s, r) of the reference
  //! (LinearEquationsDavidson.h:173-184)
  static __device__ __forceinline__ double2 residual_lineq(const double2& r, const double2& b, double s) {
    return make_double2(__dmul_rn(__dadd_rn(r.x, -b.x), s), __dmul_rn(__dadd_rn(r.y, -b.y), s));
  }
  static __device__ __forceinline__ void square_to(double& acc, const double2& r) {
    acc = fma(r.x, r.x, acc);
    acc = fma(r.y, r.y, acc);
  }
  static __device__ __forceinline__ double2 sub(const double2& x, const double2& r) {
    return make_double2(__dsub_rn(x.x, r.x), __dsub_rn(x.y, r.y));
  }
  static __device__ __forceinline__ double2 precondition(const double2& r, const double2& d, double shift) {
    return make_double2(div_rn(r.x, __dadd_rn(__dsub_rn(d.x, shift), 1e-15)),
                        div_rn(r.y, __dadd_rn(__dsub_rn(d.y, shift), 1e-15)));
  }
};
template <>
struct DsOps<double> {
  static __device__ __forceinline__ double zero() { return 0.0; }
  static __device__ __forceinline__ void fma_to(double& acc, double c, const double& v) { acc = fma(c, v, acc); }
  static __device__ __forceinline__ double residual(const double& r, double neg_lambda, const double& x) {
    return __dadd_rn(r, __dmul_rn(neg_lambda, x));
  }
  static __device__ __forceinline__ double residual_lineq(const double& r, const double& b, double s) {
    return __dmul_rn(__dadd_rn(r, -b), s);
  }
  static __device__ __forceinline__ void square_to(double& acc, const double& r) { acc = fma(r, r, acc); }
  static __device__ __forceinline__ double sub(const double& x, const double& r) { return __dsub_rn(x, r); }
  static __device__ __forceinline__ double precondition(const double& r, const double& d, double shift) {
    return div_rn(r, __dadd_rn(__dsub_rn(d, shift), 1e-15));
  }
};

//! running sums of squares of one thread: <r_j, r_j> in [0, MJ), <out_r_j, out_r_j> in [MJ, 2 MJ). Registers for up to
//! 8 roots; for 16 roots the 64 running sums of the expansions fill the register file and these live in shared memory
//! (one column per thread, conflict-free).
template <int MJ, bool SMEM>
struct DsNorms;
template <int MJ>
struct DsNorms<MJ, false> {
  double v[2 * MJ];
  __device__ __forceinline__ void init(double*) {
#pragma unroll
    for (int b = 0; b < 2 * MJ; ++b)
      v[b] = 0.0;
  }
  template <class RV>
  __device__ __forceinline__ void add(int b, const RV& r) {
    DsOps<RV>::square_to(v[b], r);
  }
  __device__ __forceinline__ double get(int b) const { return v[b]; }
};
template <int MJ>
struct DsNorms<MJ, true> {
  double* col; // this thread's column of the [2 MJ][blockDim.x] array
  __device__ __forceinline__ void init(double* base) {
    col = base + threadIdx.x;
#pragma unroll
    for (int b = 0; b < 2 * MJ; ++b)
      col[b * kDsThreads] = 0.0;
  }
  template <class RV>
  __device__ __forceinline__ void add(int b, const RV& r) {
    double t = col[b * kDsThreads];
    DsOps<RV>::square_to(t, r);
    col[b * kDsThreads] = t;
  }
  __device__ __forceinline__ double get(int b) const { return col[b * kDsThreads]; }
};

//! residual of root j at row r (units of RV) from the two expansions
template <class RV>
__device__ __forceinline__ RV ds_residual(const DsParams& p, int j, size_t r, const RV& ar, const RV& ax) {
  if (p.mode == 1)
    return DsOps<RV>::residual_lineq(ar, reinterpret_cast<const RV*>(p.rhs[j])[r], p.rscale[j]);
  if (p.mode == 2)
    return ar;
  return DsOps<RV>::residual(ar, -p.lambda[j], ax);
}

//! what is stored as x_j: the expansion, or (update_x) the expansion minus the residual that was just written - the
//! rounding of axpy(-1, out_r, x)
template <class RV>
__device__ __forceinline__ RV ds_solution(const DsParams& p, const RV& ax, const RV& out) {
  return p.update_x ? DsOps<RV>::sub(ax, out) : ax;
}

//! one thread, rows `r` (in units of RV): both expansions, residual, square, preconditioner
template <int MJ, class RV, class Norms>
__device__ __forceinline__ void ds_rows(const DsParams& p, const double* __restrict__ sc, size_t r, Norms& nrm) {
  using Ops = DsOps<RV>;
  RV ax[MJ], ar[MJ];
#pragma unroll
  for (int b = 0; b < MJ; ++b) {
    ax[b] = Ops::zero();
    ar[b] = Ops::zero();
    if (p.accumulate && b < p.m) {
      ax[b] = reinterpret_cast<const RV*>(p.out_x[b])[r];
      ar[b] = reinterpret_cast<const RV*>(p.out_r[b])[r];
    }
  }
  // subspace vectors per trip: 2U independent loads in flight per thread
  constexpr int U = (sizeof(RV) == sizeof(double) || MJ <= 4) ? 4 : 2;
  int i = 0;
  for (; i + U <= p.k; i += U) {
    RV qv[U], av[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      qv[u] = reinterpret_cast<const RV*>(p.q[i + u])[r];
      av[u] = reinterpret_cast<const RV*>(p.a[i + u])[r];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double* cu = sc + size_t(i + u) * p.ld;
#pragma unroll
      for (int b = 0; b < MJ; ++b) {
        const double c = cu[b];
        Ops::fma_to(ax[b], c, qv[u]);
        Ops::fma_to(ar[b], c, av[u]);
      }
    }
  }
  for (; i < p.k; ++i) {
    const RV q0 = reinterpret_cast<const RV*>(p.q[i])[r];
    const RV a0 = reinterpret_cast<const RV*>(p.a[i])[r];
    const double* c0 = sc + size_t(i) * p.ld;
#pragma unroll
    for (int b = 0; b < MJ; ++b) {
      const double c = c0[b];
      Ops::fma_to(ax[b], c, q0);
      Ops::fma_to(ar[b], c, a0);
    }
  }
  RV d = Ops::zero();
  if (p.diag)
    d = reinterpret_cast<const RV*>(p.diag)[r];
#pragma unroll
  for (int b = 0; b < MJ; ++b) {
    if (b < p.m) {
      RV res = ds_residual<RV>(p, b, r, ar[b], ax[b]);
      nrm.add(b, res);
      if (p.diag) {
        res = Ops::precondition(res, d, p.shift[b]);
        nrm.add(MJ + b, res);
      }
      reinterpret_cast<RV*>(p.out_r[b])[r] = res;
      if (p.write_x)
        reinterpret_cast<RV*>(p.out_x[b])[r] = ds_solution<RV>(p, ax[b], res);
    }
  }
}

//! CTA fold of the running sums, per-CTA partials laid out [<r,r> of the m roots | <out_r,out_r> of the m roots], finish
template <int MJ, class Norms>
__device__ __forceinline__ void ds_finish(const DsParams& p, const Norms& nrm) {
  __shared__ double s_part[kDsThreads / 32][2 * MJ];
  __shared__ int s_is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int b = 0; b < 2 * MJ; ++b) {
    double v = nrm.get(b);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
      v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0)
      s_part[warp][b] = v;
  }
  __syncthreads();
  if (threadIdx.x < 2 * p.m) {
    const int half = threadIdx.x / p.m, j = threadIdx.x % p.m;
    double sum = 0.0;
#pragma unroll
    for (int w = 0; w < kDsThreads / 32; ++w)
      sum += s_part[w][half * MJ + j];
    p.fin.partials[size_t(blockIdx.x) * (2 * p.m) + threadIdx.x] = sum;
  }
  gi_finalize(p.fin, 2 * p.m, &s_is_last);
}

template <int MJ, bool VEC>
__global__ void __launch_bounds__(kDsThreads, MJ <= 2 ? 3 : 2)
    davidson_residual_kernel(const __grid_constant__ DsParams p) {
  extern __shared__ __align__(16) double sc[]; // k x ld coefficients [, 2 MJ x blockDim.x running sums]
  constexpr bool kSmemNorms = MJ > 8;
  for (int e = threadIdx.x; e < p.k * p.ld; e += blockDim.x)
    sc[e] = p.coef[e];
  DsNorms<MJ, kSmemNorms> nrm;
  nrm.init(sc + p.k * p.ld);
  __syncthreads();
  const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = size_t(gridDim.x) * blockDim.x;
  if (VEC) {
    const size_t npairs = p.n / 2;
    for (size_t r = tid; r < npairs; r += nthreads)
      ds_rows<MJ, double2>(p, sc, r, nrm);
    if ((p.n & 1) && tid == 0)
      ds_rows<MJ, double>(p, sc, p.n - 1, nrm);
  } else {
    for (size_t r = tid; r < p.n; r += nthreads)
      ds_rows<MJ, double>(p, sc, r, nrm);
  }
  ds_finish<MJ>(p, nrm);
}

// ---- the same pass with a thread-private cp.async ring --------------------------------------------------------------
// The kernel above keeps 2U loads per thread in flight only while it waits for them; during the arithmetic, the divisions
// and the stores of a row nothing is outstanding, and with the register budget of 16 warps per SM the average falls short
// of the ~45 KB per SM that HBM3e needs (measured 4.6 TB/s at k = 4..16). Here every thread copies the operands of its
// NEXT S-1 trips (U subspace vectors and their actions, plus the diagonal with the first trip of a row) into its own
// slots of a shared-memory ring with cp.async (LDGSTS, 16 bytes each, no registers, no barriers: slots are private) and
// computes the current trip out of shared memory, so (S-1) 2U 16-byte loads per thread stay outstanding throughout.
constexpr int kRingU = 2; // subspace vectors (and as many actions) per trip
constexpr int kRingSlots = 2 * kRingU + 1;
//! ring depth: S-1 trips are in flight while one is consumed. With 8 roots the running sums of squares move to shared
//! memory (the 32 sums of the expansions need the registers) and the ring gives up one stage for them.
__host__ __device__ constexpr int ring_depth(int mj) { return mj >= 8 ? 3 : 4; }
static size_t ds_ring_bytes(int mj) {
  return size_t(ring_depth(mj)) * kRingSlots * kDsThreads * sizeof(double2) +
         (mj >= 8 ? size_t(2 * mj) * kDsThreads * sizeof(double) : 0);
}

__device__ __forceinline__ void ds_cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void ds_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ds_cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int MJ>
__global__ void __launch_bounds__(kDsThreads, 2) davidson_residual_ring_kernel(const __grid_constant__ DsParams p) {
  // k x ld coefficients, then the ring [S][2U+1][blockDim.x] of double2 [, then 2 MJ x blockDim.x running sums]
  extern __shared__ __align__(16) double sc[];
  constexpr int U = kRingU, S = ring_depth(MJ);
  for (int e = threadIdx.x; e < p.k * p.ld; e += blockDim.x)
    sc[e] = p.coef[e];
  double2* ring_base = reinterpret_cast<double2*>(sc + ((p.k * p.ld + 1) & ~1));
  double2* ring = ring_base + threadIdx.x;
  DsNorms<MJ, (MJ >= 8)> nrm;
  nrm.init(reinterpret_cast<double*>(ring_base + size_t(S) * kRingSlots * kDsThreads));
  __syncthreads();
  const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = size_t(gridDim.x) * blockDim.x;
  const size_t npairs = p.n / 2;
  const int nchunks = (p.k + U - 1) / U;
  const size_t my_rows = tid < npairs ? (npairs - tid + nthreads - 1) / nthreads : 0;
  const size_t trips = my_rows * size_t(nchunks);
  // cursor of the copies
  size_t it = 0, ir = tid;
  int ic = 0, is = 0;
  auto issue = [&]() {
    if (it < trips) {
      double2* st = ring + size_t(is) * (kRingSlots * kDsThreads);
      const int i0 = ic * U;
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (i0 + u < p.k) {
          ds_cp_async16(st + u * kDsThreads, reinterpret_cast<const double2*>(p.q[i0 + u]) + ir);
          ds_cp_async16(st + (U + u) * kDsThreads, reinterpret_cast<const double2*>(p.a[i0 + u]) + ir);
        }
      if (ic == 0 && p.diag)
        ds_cp_async16(st + 2 * U * kDsThreads, reinterpret_cast<const double2*>(p.diag) + ir);
      if (++ic == nchunks) {
        ic = 0;
        ir += nthreads;
      }
      is = is + 1 == S ? 0 : is + 1;
      ++it;
    }
    ds_cp_async_commit();
  };
#pragma unroll
  for (int d = 0; d < S - 1; ++d)
    issue();
  using Ops = DsOps<double2>;
  double2 ax[MJ], ar[MJ];
  double2 dg = Ops::zero();
  size_t cr = tid;
  int cc = 0, cs = 0;
  for (size_t t = 0; t < trips; ++t) {
    issue();
    ds_cp_async_wait<S - 1>();
    const double2* st = ring + size_t(cs) * (kRingSlots * kDsThreads);
    if (cc == 0) {
#pragma unroll
      for (int b = 0; b < MJ; ++b) {
        ax[b] = Ops::zero();
        ar[b] = Ops::zero();
        if (p.accumulate && b < p.m) {
          ax[b] = reinterpret_cast<const double2*>(p.out_x[b])[cr];
          ar[b] = reinterpret_cast<const double2*>(p.out_r[b])[cr];
        }
      }
      if (p.diag)
        dg = st[2 * U * kDsThreads];
    }
    const int i0 = cc * U;
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + u < p.k) {
        const double2 qv = st[u * kDsThreads], av = st[(U + u) * kDsThreads];
        const double* cu = sc + size_t(i0 + u) * p.ld;
#pragma unroll
        for (int b = 0; b < MJ; ++b) {
          const double c = cu[b];
          Ops::fma_to(ax[b], c, qv);
          Ops::fma_to(ar[b], c, av);
        }
      }
    if (cc == nchunks - 1) {
#pragma unroll
      for (int b = 0; b < MJ; ++b) {
        if (b < p.m) {
          double2 res = ds_residual<double2>(p, b, cr, ar[b], ax[b]);
          nrm.add(b, res);
          if (p.diag) {
            res = Ops::precondition(res, dg, p.shift[b]);
            nrm.add(MJ + b, res);
          }
          reinterpret_cast<double2*>(p.out_r[b])[cr] = res;
          if (p.write_x)
            reinterpret_cast<double2*>(p.out_x[b])[cr] = ds_solution<double2>(p, ax[b], res);
        }
      }
      cr += nthreads;
      cc = 0;
    } else {
      ++cc;
    }
    cs = cs + 1 == S ? 0 : cs + 1;
  }
  ds_cp_async_wait<0>();
  if ((p.n & 1) && tid == 0)
    ds_rows<MJ, double>(p, sc, p.n - 1, nrm);
  ds_finish<MJ>(p, nrm);
}


// ---- 9..16 roots: shared-memory tiles moved by TMA, consumed by two root groups -----------------------------------
// With 16 roots the 2 x 16 running sums of one row pair (64 registers for a double2 lane) leave no room to keep enough
// loads in flight from registers: the kernel above falls back to one row per thread and reaches half the copy rate. Here
// the CTA's producer warps stream tiles of kTileRows rows of KC vector pairs at a time into a ring of stages with 1-D TMA
// bulk copies (cp.async.bulk + mbarrier, as the Gram kernels, gi_pipeline.cuh); 512 consumer threads form two groups of
// 256: thread (rp, g) owns row pair rp of the tile and roots 8g .. 8g+7, i.e. 2 x 8 double2 running sums, and both groups
// read the same tile, so every HBM byte is still moved once: 8n(2k + m + 1) bytes. The k-loop runs across the stages of
// a row tile in ascending i (one FMA per term, from zero), the epilogue is the one of the kernels above: bit-identical
// vectors.
constexpr int kTileRows = 512;               // rows per tile = 2 x 256 row pairs
constexpr int kTileKC = 4;                   // vector pairs per stage
constexpr int kTileSlots = 2 * kTileKC + 1;  // q and a of the chunk + the diagonal (with the first chunk of a tile)
constexpr int kTileStages = 4;               // 4 x 9 x 4 KB = 144 KB (+ 64 KB of running squares, + the coefficients)
constexpr int kTileConsumers = 512;
constexpr int kTileProducers = 4;            // warps
constexpr int kTileThreads = kTileConsumers + 32 * kTileProducers;
constexpr int kTileGroupRoots = 8;
static size_t ds_tile_bytes(int k) {
  return size_t(kTileStages) * kTileSlots * kTileRows * sizeof(double) +
         size_t(2 * kTileGroupRoots) * kTileConsumers * sizeof(double) + size_t(k) * 16 * sizeof(double) + 128;
}

__global__ void __launch_bounds__(kTileThreads, 1) davidson_residual_tile_kernel(const __grid_constant__ DsParams p) {
  extern __shared__ __align__(128) unsigned char ds_smem_raw[];
  double* stages = reinterpret_cast<double*>(ds_smem_raw);
  // running <r,r> [0,8) and <out,out> [8,16) of each consumer thread's roots: one column per thread (conflict-free);
  // in shared memory because the 2 x 8 double2 running sums of the expansions take the registers
  double* sq = stages + size_t(kTileStages) * kTileSlots * kTileRows;
  double* sc = sq + size_t(2 * kTileGroupRoots) * kTileConsumers; // k x 16 coefficients
  __shared__ uint64_t full_bar[kTileStages];
  __shared__ uint64_t empty_bar[kTileStages];
  __shared__ double s_part[kTileConsumers / 32][2 * kTileGroupRoots];
  __shared__ int s_is_last;
  using Ops = DsOps<double2>;
  const int tid = threadIdx.x;
  const bool is_producer = tid >= kTileConsumers;
  for (int e = tid; e < p.k * 16; e += blockDim.x)
    sc[e] = p.coef[e];
  if (tid == 0) {
    for (int s = 0; s < kTileStages; ++s) {
      mbar_init(&full_bar[s], kTileProducers);
      mbar_init(&empty_bar[s], kTileConsumers / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long nfull = (long long)(p.n / kTileRows);
  const long long my_tiles = nfull > (long long)blockIdx.x ? (nfull - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int nchunks = (p.k + kTileKC - 1) / kTileKC;
  constexpr uint32_t kVecBytes = kTileRows * 8u;
  if (!is_producer)
    for (int b = 0; b < 2 * kTileGroupRoots; ++b)
      sq[b * kTileConsumers + tid] = 0.0;

  if (is_producer) {
    const int pw = (tid - kTileConsumers) >> 5;
    const bool leader = elect_one();
    const uint32_t base = smem_u32(stages);
    long long seq = 0;
    for (long long s = 0; s < my_tiles; ++s) {
      const size_t row0 = size_t(blockIdx.x + s * gridDim.x) * kTileRows;
      for (int c = 0; c < nchunks; ++c, ++seq) {
        const int stage = int(seq % kTileStages);
        const long long use = seq / kTileStages;
        if (use > 0)
          mbar_wait(&empty_bar[stage], uint32_t((use - 1) & 1));
        const int i0 = c * kTileKC;
        const int kc = p.k - i0 < kTileKC ? p.k - i0 : kTileKC;
        const int ncopies = 2 * kc + ((c == 0 && p.diag) ? 1 : 0);
        const int mine = ncopies > pw ? (ncopies - pw + kTileProducers - 1) / kTileProducers : 0;
        const uint32_t bar = smem_u32(&full_bar[stage]);
        const uint32_t st = base + uint32_t(stage) * uint32_t(kTileSlots) * kVecBytes;
        if (leader) {
          mbar_expect_tx(&full_bar[stage], kVecBytes * uint32_t(mine));
          for (int e = pw; e < ncopies; e += kTileProducers) {
            // copy e of the stage: q of the chunk, then a of the chunk, then the diagonal
            const double* src;
            int slot;
            if (e < kc) {
              src = p.q[i0 + e];
              slot = e;
            } else if (e < 2 * kc) {
              src = p.a[i0 + e - kc];
              slot = kTileKC + (e - kc);
            } else {
              src = p.diag;
              slot = 2 * kTileKC;
            }
            bulk_load(st + uint32_t(slot) * kVecBytes, src + row0, kVecBytes, bar);
          }
        }
        __syncwarp();
      }
    }
  } else {
    const int rp = tid & 255, g = tid >> 8;
    double* mysq = sq + tid;
    const int lane = tid & 31;
    const double* cg = sc + g * kTileGroupRoots;
    double2 ax[kTileGroupRoots], ar[kTileGroupRoots];
    double2 dg = Ops::zero();
    long long seq = 0;
    for (long long s = 0; s < my_tiles; ++s) {
      const size_t prow = (size_t(blockIdx.x + s * gridDim.x) * kTileRows) / 2 + rp; // this thread's row pair
#pragma unroll
      for (int b = 0; b < kTileGroupRoots; ++b) {
        ax[b] = Ops::zero();
        ar[b] = Ops::zero();
        if (p.accumulate && g * kTileGroupRoots + b < p.m) {
          ax[b] = reinterpret_cast<const double2*>(p.out_x[g * kTileGroupRoots + b])[prow];
          ar[b] = reinterpret_cast<const double2*>(p.out_r[g * kTileGroupRoots + b])[prow];
        }
      }
      for (int c = 0; c < nchunks; ++c, ++seq) {
        const int stage = int(seq % kTileStages);
        mbar_wait(&full_bar[stage], uint32_t((seq / kTileStages) & 1));
        const double2* st = reinterpret_cast<const double2*>(stages + size_t(stage) * kTileSlots * kTileRows) + rp;
        if (c == 0 && p.diag)
          dg = st[2 * kTileKC * (kTileRows / 2)];
        const int i0 = c * kTileKC;
#pragma unroll
        for (int u = 0; u < kTileKC; ++u) {
          if (i0 + u < p.k) {
            const double2 qv = st[u * (kTileRows / 2)], av = st[(kTileKC + u) * (kTileRows / 2)];
            const double2* cu = reinterpret_cast<const double2*>(cg + size_t(i0 + u) * 16);
#pragma unroll
            for (int b2 = 0; b2 < kTileGroupRoots / 2; ++b2) {
              const double2 cc = cu[b2];
              Ops::fma_to(ax[2 * b2], cc.x, qv);
              Ops::fma_to(ar[2 * b2], cc.x, av);
              Ops::fma_to(ax[2 * b2 + 1], cc.y, qv);
              Ops::fma_to(ar[2 * b2 + 1], cc.y, av);
            }
          }
        }
        __syncwarp();
        if (lane == 0)
          mbar_arrive(&empty_bar[stage]);
      }
#pragma unroll
      for (int b = 0; b < kTileGroupRoots; ++b) {
        const int j = g * kTileGroupRoots + b;
        if (j < p.m) {
          double2 res = ds_residual<double2>(p, j, prow, ar[b], ax[b]);
          double t = mysq[b * kTileConsumers];
          Ops::square_to(t, res);
          mysq[b * kTileConsumers] = t;
          if (p.diag) {
            res = Ops::precondition(res, dg, p.shift[j]);
            t = mysq[(kTileGroupRoots + b) * kTileConsumers];
            Ops::square_to(t, res);
            mysq[(kTileGroupRoots + b) * kTileConsumers] = t;
          }
          reinterpret_cast<double2*>(p.out_r[j])[prow] = res;
          if (p.write_x)
            reinterpret_cast<double2*>(p.out_x[j])[prow] = ds_solution<double2>(p, ax[b], res);
        }
      }
    }
    // rows past the last full tile: the CTA that would own tile number nfull takes them straight from global memory,
    // one row per thread of group g (same chain of operations)
    const size_t tail0 = size_t(nfull) * kTileRows;
    if (tail0 < p.n && int(nfull % gridDim.x) == int(blockIdx.x)) {
      for (size_t r = tail0 + size_t(rp); r < p.n; r += 256) {
        double x1[kTileGroupRoots], r1[kTileGroupRoots];
#pragma unroll
        for (int b = 0; b < kTileGroupRoots; ++b) {
          x1[b] = r1[b] = 0.0;
          if (p.accumulate && g * kTileGroupRoots + b < p.m) {
            x1[b] = p.out_x[g * kTileGroupRoots + b][r];
            r1[b] = p.out_r[g * kTileGroupRoots + b][r];
          }
        }
        for (int i = 0; i < p.k; ++i) {
          const double q0 = p.q[i][r], a0 = p.a[i][r];
          const double* cu = cg + size_t(i) * 16;
#pragma unroll
          for (int b = 0; b < kTileGroupRoots; ++b) {
            x1[b] = fma(cu[b], q0, x1[b]);
            r1[b] = fma(cu[b], a0, r1[b]);
          }
        }
        const double d0 = p.diag ? p.diag[r] : 0.0;
#pragma unroll
        for (int b = 0; b < kTileGroupRoots; ++b) {
          const int j = g * kTileGroupRoots + b;
          if (j < p.m) {
            double res = ds_residual<double>(p, j, r, r1[b], x1[b]);
            double t = mysq[b * kTileConsumers];
            DsOps<double>::square_to(t, res);
            mysq[b * kTileConsumers] = t;
            if (p.diag) {
              res = DsOps<double>::precondition(res, d0, p.shift[j]);
              t = mysq[(kTileGroupRoots + b) * kTileConsumers];
              DsOps<double>::square_to(t, res);
              mysq[(kTileGroupRoots + b) * kTileConsumers] = t;
            }
            p.out_r[j][r] = res;
            if (p.write_x)
              p.out_x[j][r] = ds_solution<double>(p, x1[b], res);
          }
        }
      }
    }
    // warp fold of the 16 running sums of this thread's group
    const int warp = tid >> 5;
#pragma unroll
    for (int b = 0; b < kTileGroupRoots; ++b) {
      double v = mysq[b * kTileConsumers], w = mysq[(kTileGroupRoots + b) * kTileConsumers];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        v += __shfl_down_sync(0xffffffffu, v, off);
        w += __shfl_down_sync(0xffffffffu, w, off);
      }
      if (lane == 0) {
        s_part[warp][b] = v;
        s_part[warp][kTileGroupRoots + b] = w;
      }
    }
  }
  __syncthreads();
  if (tid < 2 * p.m) { // partials laid out [<r,r> of the m roots | <out,out> of the m roots]
    const int half = tid / p.m, j = tid % p.m;
    const int g = j / kTileGroupRoots, b = j % kTileGroupRoots;
    double sum = 0.0;
    for (int w = 0; w < 8; ++w) // the 8 warps of group g, in warp order
      sum += s_part[g * 8 + w][half * kTileGroupRoots + b];
    p.fin.partials[size_t(blockIdx.x) * (2 * p.m) + tid] = sum;
  }
  gi_finalize(p.fin, 2 * p.m, &s_is_last);
}

using DsKernel = void (*)(const DsParams);
template <int MJ>
static DsKernel ds_pick_vec(bool vec) {
  return vec ? davidson_residual_kernel<MJ, true> : davidson_residual_kernel<MJ, false>;
}
static DsKernel ds_pick_ring(int mj) {
  switch (mj) {
  case 1:
    return davidson_residual_ring_kernel<1>;
  case 2:
    return davidson_residual_ring_kernel<2>;
  case 4:
    return davidson_residual_ring_kernel<4>;
  case 8:
    return davidson_residual_ring_kernel<8>;
  }
  return nullptr;
}
static DsKernel ds_pick(int mj, bool vec) {
  switch (mj) {
  case 1:
    return ds_pick_vec<1>(vec);
  case 2:
    return ds_pick_vec<2>(vec);
  case 4:
    return ds_pick_vec<4>(vec);
  case 8:
    return ds_pick_vec<8>(vec);
  case 16:
    return davidson_residual_kernel<16, false>;
  }
  return nullptr;
}

} // namespace itsolv

using namespace itsolv;

extern "C" {

int itsolv_davidson_residual_f64(itsolv_ctx* ctx, const double* coef, int k, int m, const double* const* q,
                                 const double* const* a, const double* lambda, const double* diag, const double* shift,
                                 double* const* out_x, double* const* out_r, size_t n, double* norm2,
                                 double* norm2_out) {
  return itsolv_subspace_residual_f64(ctx, 0, 0, coef, k, m, q, a, lambda, nullptr, nullptr, diag, shift, out_x, out_r, n,
                                      norm2, norm2_out);
}

int itsolv_subspace_residual_f64(itsolv_ctx* ctx, int mode, int accumulate, const double* coef, int k, int m,
                                 const double* const* q, const double* const* a, const double* lambda,
                                 const double* const* rhs, const double* rscale, const double* diag, const double* shift,
                                 double* const* out_x, double* const* out_r, size_t n, double* norm2,
                                 double* norm2_out) {
  ITSOLV_REQUIRE(k >= 1 && k <= ITSOLV_MAX_PANEL, "davidson_residual: 1 <= k <= ITSOLV_MAX_PANEL subspace vectors");
  ITSOLV_REQUIRE(m >= 1, "davidson_residual: m >= 1 roots");
  ITSOLV_REQUIRE(mode >= 0 && mode <= 3,
                 "subspace_residual: mode 0 (eigenproblem), 1 (linear equations), 2 (plain expansion), 3 (2 + DIIS step)");
  const int update_x = mode == 3 ? 1 : 0;
  if (mode == 3)
    mode = 2;
  ITSOLV_REQUIRE(out_r != nullptr && q != nullptr && a != nullptr && coef != nullptr && (mode != 0 || lambda != nullptr),
                 "davidson_residual: null argument");
  ITSOLV_REQUIRE(mode != 1 || (rhs != nullptr && rscale != nullptr), "subspace_residual: right-hand sides and scales");
  ITSOLV_REQUIRE(!update_x || out_x != nullptr, "subspace_residual: the DIIS step needs the solution vectors");
  ITSOLV_REQUIRE(!accumulate || out_x != nullptr, "subspace_residual: accumulation needs the solution vectors");
  ITSOLV_REQUIRE(diag == nullptr || shift != nullptr, "davidson_residual: shifts are needed with a diagonal");
  if (mode == 1)
    for (int j = 0; j < m; ++j)
      for (int j2 = 0; j2 < m; ++j2)
        ITSOLV_REQUIRE(out_r[j] != rhs[j2] && (!out_x || out_x[j] != rhs[j2]),
                       "subspace_residual: an output vector is also a right-hand side");
  for (int j = 0; j < m; ++j) { // the outputs must not be inputs of the same pass
    for (int i = 0; i < k; ++i)
      ITSOLV_REQUIRE(out_r[j] != q[i] && out_r[j] != a[i] && (!out_x || (out_x[j] != q[i] && out_x[j] != a[i])),
                     "davidson_residual: an output vector is also a subspace vector");
    ITSOLV_REQUIRE(out_r[j] != diag && (!out_x || out_x[j] != diag), "davidson_residual: an output vector is the diagonal");
  }
  ctx->counters.n_gemm_outer += 2;
  ctx->counters.n_axpy += m;
  ctx->counters.n_dot += m;
  if (diag)
    ctx->counters.n_precondition++;
  for (int j0 = 0; j0 < m; j0 += kDsMaxRoots) {
    const int mb = std::min(kDsMaxRoots, m - j0);
    const double bytes = 8.0 * double(n) * (2.0 * k + mb * (out_x ? 2.0 : 1.0) + (diag ? 1.0 : 0.0) +
                                            (mode == 1 ? mb : 0) + (accumulate ? 2.0 * mb : 0.0));
    CallScope scope(ctx, OP_RESIDUAL, bytes);
    bool direct = false;
    if (n == 0) { // an empty shard still takes part in the all-reduce
      ITSOLV_CUDA(cudaMemsetAsync(ctx->d_result, 0, size_t(2 * mb) * sizeof(double), ctx->stream));
      if (finish_with_peers(ctx, 2 * mb, &direct))
        return 1;
    } else {
      DsParams p;
      bool vec = diag == nullptr || aligned16(diag);
      for (int i = 0; i < k; ++i) {
        p.q[i] = q[i];
        p.a[i] = a[i];
        vec = vec && aligned16(q[i]) && aligned16(a[i]);
      }
      for (int j = 0; j < kDsMaxRoots; ++j) {
        const int jj = j0 + std::min(j, mb - 1);
        p.out_x[j] = out_x ? out_x[jj] : nullptr;
        p.out_r[j] = out_r[jj];
        p.lambda[j] = mode == 0 ? lambda[jj] : 0.0;
        p.shift[j] = diag ? shift[jj] : 0.0;
        p.rhs[j] = mode == 1 ? rhs[jj] : nullptr;
        p.rscale[j] = mode == 1 ? rscale[jj] : 1.0;
        vec = vec && aligned16(p.out_r[j]) && (!out_x || aligned16(p.out_x[j])) && (mode == 0 || aligned16(p.rhs[j]));
      }
      int mj = 1;
      while (mj < mb)
        mj *= 2;
      p.ld = mj;
      p.k = k;
      p.m = mb;
      p.n = n;
      p.diag = diag;
      p.write_x = out_x ? 1 : 0;
      p.mode = mode;
      p.update_x = update_x;
      p.accumulate = accumulate ? 1 : 0;
      char *h = nullptr, *d = nullptr;
      int slot = 0;
      const size_t cbytes = size_t(k) * p.ld * sizeof(double);
      if (stage_acquire(ctx, cbytes, &h, &d, &slot))
        return 1;
      double* hc = reinterpret_cast<double*>(h);
      for (int i = 0; i < k; ++i)
        for (int j = 0; j < p.ld; ++j)
          hc[size_t(i) * p.ld + j] = j < mb ? coef[size_t(i) * m + (j0 + j)] : 0.0;
      if (stage_commit(ctx, slot, cbytes))
        return 1;
      p.coef = reinterpret_cast<const double*>(d);
      // 9..16 roots: the TMA tile kernel (needs 16-byte aligned vectors and at least one full tile per CTA to pay off)
      const bool tile = mj == 16 && vec && ctx->opt_ds_ring >= 0 && n >= size_t(kTileRows) * 8 &&
                        ds_tile_bytes(k) <= size_t(ctx->max_smem_optin);
      if (mj == 16 && !tile)
        vec = false; // 16 roots: one row per thread keeps the 32 running sums in registers at two CTAs per SM
      // measured (profiles/opbench): the ring wins while the epilogue (divisions, stores) is a large part of a row's work,
      // the register kernel once the subspace has more than ~10 vector pairs
      const bool ring = vec && mj <= 8 && ctx->opt_ds_ring >= 0 && (k <= 10 || mj == 8 || ctx->opt_ds_ring > 0) &&
                        ((cbytes + 15) & ~size_t(15)) + ds_ring_bytes(mj) <= size_t(ctx->max_smem_optin) / 2 - 1024;
      DsKernel kernel = tile ? davidson_residual_tile_kernel : (ring ? ds_pick_ring(mj) : ds_pick(mj, vec));
      ITSOLV_REQUIRE(kernel != nullptr, "davidson_residual: root tile not instantiated");
      const size_t smem = tile ? ds_tile_bytes(k)
                               : ring ? ((cbytes + 15) & ~size_t(15)) + ds_ring_bytes(mj)
                                      : cbytes + (mj > 8 ? size_t(2 * mj) * kDsThreads * sizeof(double) : 0);
      if (ensure_dynamic_smem(ctx, reinterpret_cast<const void*>(kernel), smem))
        return 1;
      const int per_sm = tile ? 1 : (ring ? 2 : (mj <= 2 ? 3 : 2));
      const size_t units = vec ? n / 2 : n;
      const int threads = tile ? kTileThreads : kDsThreads;
      const int grid = tile ? int(std::min<size_t>((n + kTileRows - 1) / kTileRows, size_t(ctx->num_sms)))
                            : int(std::max<size_t>(1, std::min<size_t>((units + kDsThreads - 1) / kDsThreads,
                                                                       size_t(ctx->num_sms) * per_sm)));
      if (ensure_partials(ctx, size_t(grid) * 2 * mb))
        return 1;
      fill_finalize(ctx, ctx->num_sms * per_sm, 2 * mb, &p.fin, &direct);
      mark_launch(ctx);
      kernel<<<grid, threads, smem, ctx->stream>>>(p);
      ITSOLV_CUDA(cudaGetLastError());
      ctx->counters.launches += 1;
      if (stage_done(ctx, slot))
        return 1;
      if (!p.fin.fused) {
        if (launch_reduce_partials(ctx, grid, 2 * mb))
          return 1;
        if (finish_with_peers(ctx, 2 * mb, &direct))
          return 1;
      }
    }
    scope.stop();
    double sums[2 * kDsMaxRoots];
    if (finish_result(ctx, 2 * mb, sums, direct))
      return 1;
    for (int j = 0; j < mb; ++j) {
      if (norm2)
        norm2[j0 + j] = sums[j];
      if (norm2_out)
        norm2_out[j0 + j] = diag ? sums[mb + j] : sums[j];
    }
  }
  return 0;
}

} // extern "C"
