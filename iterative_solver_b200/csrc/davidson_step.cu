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
  GiFinalize fin;
  size_t n;
  int k, m, ld;
  int write_x;
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
  static __device__ __forceinline__ void square_to(double& acc, const double2& r) {
    acc = fma(r.x, r.x, acc);
    acc = fma(r.y, r.y, acc);
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
  static __device__ __forceinline__ void square_to(double& acc, const double& r) { acc = fma(r, r, acc); }
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

//! one thread, rows `r` (in units of RV): both expansions, residual, square, preconditioner
template <int MJ, class RV, class Norms>
__device__ __forceinline__ void ds_rows(const DsParams& p, const double* __restrict__ sc, size_t r, Norms& nrm) {
  using Ops = DsOps<RV>;
  RV ax[MJ], ar[MJ];
#pragma unroll
  for (int b = 0; b < MJ; ++b) {
    ax[b] = Ops::zero();
    ar[b] = Ops::zero();
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
      if (p.write_x)
        reinterpret_cast<RV*>(p.out_x[b])[r] = ax[b];
      RV res = Ops::residual(ar[b], -p.lambda[b], ax[b]);
      nrm.add(b, res);
      if (p.diag) {
        res = Ops::precondition(res, d, p.shift[b]);
        nrm.add(MJ + b, res);
      }
      reinterpret_cast<RV*>(p.out_r[b])[r] = res;
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
          if (p.write_x)
            reinterpret_cast<double2*>(p.out_x[b])[cr] = ax[b];
          double2 res = Ops::residual(ar[b], -p.lambda[b], ax[b]);
          nrm.add(b, res);
          if (p.diag) {
            res = Ops::precondition(res, dg, p.shift[b]);
            nrm.add(MJ + b, res);
          }
          reinterpret_cast<double2*>(p.out_r[b])[cr] = res;
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
  ITSOLV_REQUIRE(k >= 1 && k <= ITSOLV_MAX_PANEL, "davidson_residual: 1 <= k <= ITSOLV_MAX_PANEL subspace vectors");
  ITSOLV_REQUIRE(m >= 1, "davidson_residual: m >= 1 roots");
  ITSOLV_REQUIRE(out_r != nullptr && q != nullptr && a != nullptr && coef != nullptr && lambda != nullptr,
                 "davidson_residual: null argument");
  ITSOLV_REQUIRE(diag == nullptr || shift != nullptr, "davidson_residual: shifts are needed with a diagonal");
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
    const double bytes = 8.0 * double(n) * (2.0 * k + mb * (out_x ? 2.0 : 1.0) + (diag ? 1.0 : 0.0));
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
        p.lambda[j] = lambda[jj];
        p.shift[j] = diag ? shift[jj] : 0.0;
        vec = vec && aligned16(p.out_r[j]) && (!out_x || aligned16(p.out_x[j]));
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
      if (mj == 16)
        vec = false; // 16 roots: one row per thread keeps the 32 running sums in registers at two CTAs per SM
      // measured (profiles/opbench): the ring wins while the epilogue (divisions, stores) is a large part of a row's work,
      // the register kernel once the subspace has more than ~10 vector pairs
      const bool ring = vec && mj <= 8 && ctx->opt_ds_ring >= 0 && (k <= 10 || mj == 8 || ctx->opt_ds_ring > 0) &&
                        ((cbytes + 15) & ~size_t(15)) + ds_ring_bytes(mj) <= size_t(ctx->max_smem_optin) / 2 - 1024;
      DsKernel kernel = ring ? ds_pick_ring(mj) : ds_pick(mj, vec);
      ITSOLV_REQUIRE(kernel != nullptr, "davidson_residual: root tile not instantiated");
      const size_t smem = ring ? ((cbytes + 15) & ~size_t(15)) + ds_ring_bytes(mj)
                               : cbytes + (mj > 8 ? size_t(2 * mj) * kDsThreads * sizeof(double) : 0);
      if (ensure_dynamic_smem(ctx, reinterpret_cast<const void*>(kernel), smem))
        return 1;
      const int per_sm = ring ? 2 : (mj <= 2 ? 3 : 2);
      const size_t units = vec ? n / 2 : n;
      const int grid = int(std::max<size_t>(
          1, std::min<size_t>((units + kDsThreads - 1) / kDsThreads, size_t(ctx->num_sms) * per_sm)));
      if (ensure_partials(ctx, size_t(grid) * 2 * mb))
        return 1;
      fill_finalize(ctx, ctx->num_sms * per_sm, 2 * mb, &p.fin, &direct);
      mark_launch(ctx);
      kernel<<<grid, kDsThreads, smem, ctx->stream>>>(p);
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
