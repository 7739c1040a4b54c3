// Harness operator kernels: the user's side of the solver (Problem::action / diagonals / p_action, reference
// src/molpro/linalg/itsolv/IterativeSolver.h:90-172) for the synthetic benchmark operator of SURVEY.md section 8(d):
//   A(i,i) = i+1,  A(i,j) = eps*(1 + ((i+j) mod 7))  for 0 < |i-j| <= b      (symmetric, diagonally dominant, banded)
// either generated on the fly or read from CSR arrays. Row sums run over ascending column with the product rounded before
// the sum, the same order as the CPU twin used by the oracle (oracle/ref_driver.cpp BandedProblemHost::apply), so both
// sides of a parity test see bit-identical operator actions. Not part of the measured subspace path.
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace itsolv {

__device__ __forceinline__ double band_entry(long long i, long long j, double eps) {
  return i == j ? double(i + 1) : __dmul_rn(eps, double(1 + ((i + j) % 7)));
}

//! x value at global column c given the local shard [off, off+n) and the two halos of b rows
__device__ __forceinline__ double x_at(long long c, long long off, long long n, int b, const double* __restrict__ x,
                                       const double* __restrict__ x_lo, const double* __restrict__ x_hi) {
  const long long l = c - off;
  if (l < 0)
    return x_lo[b + l];
  if (l >= n)
    return x_hi[l - n];
  return x[l];
}

__global__ void __launch_bounds__(256)
    banded_apply_kernel(long long n_global, long long off, long long n, int b, double eps, const double* __restrict__ x,
                        const double* __restrict__ x_lo, const double* __restrict__ x_hi, double* __restrict__ y) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const long long i = off + r;
    const long long lo = i - b > 0 ? i - b : 0;
    const long long hi = i + b < n_global - 1 ? i + b : n_global - 1;
    // (i + j) mod 7 with one 64-bit remainder per row: j = i + d, so (i + j) mod 7 = (2 (i mod 7) + d) mod 7
    const int i7 = int(i % 7);
    int e7 = (2 * i7 + int(lo - i) + 7 * ((b + 6) / 7 + 1)) % 7;
    double acc = 0.0;
    for (long long j = lo; j <= hi; ++j) {
      const double aij = i == j ? double(i + 1) : __dmul_rn(eps, double(1 + e7));
      acc = __dadd_rn(acc, __dmul_rn(aij, x_at(j, off, n, b, x, x_lo, x_hi)));
      e7 = e7 == 6 ? 0 : e7 + 1;
    }
    y[r] = acc;
  }
}

__global__ void __launch_bounds__(256)
    csr_apply_kernel(long long off, long long n, int b, const long long* __restrict__ row_ptr, const int* __restrict__ col,
                     const double* __restrict__ val, const double* __restrict__ x, const double* __restrict__ x_lo,
                     const double* __restrict__ x_hi, double* __restrict__ y) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (long long e = row_ptr[r]; e < row_ptr[r + 1]; ++e)
      acc = __dadd_rn(acc, __dmul_rn(val[e], x_at(col[e], off, n, b, x, x_lo, x_hi)));
    y[r] = acc;
  }
}

struct CsrMultiParams {
  const double* x[16];
  const double* x_lo[16];
  const double* x_hi[16];
  double* y[16];
  const long long* row_ptr;
  const int* col;
  const double* val;
  long long off, n;
  int b, w;
};

//! y_k = A x_k for w vectors at once: the matrix (12 bytes per entry) is read once instead of w times.
//! A CTA takes kCsrRows consecutive rows per trip. Their entries are contiguous in val/col and the x values they touch
//! lie in a window of kCsrRows + 2b columns, so everything is staged into shared memory with asynchronous copies
//! (cp.async / LDGSTS: all loads of a trip are in flight at once, no register staging) and each thread then walks its own
//! row entirely out of shared memory. Columns outside the window (a matrix that is not banded) are read from global
//! memory. Row sums keep the ascending-column order with the product rounded before the sum (bit-identical to the CPU twin).
constexpr int kCsrRows = 256;
constexpr int kCsrChunk = kCsrRows * 10; // entries staged at a time; rows with more entries take several rounds
constexpr int kCsrMaxHalfBand = 128;     // widest x window staged in shared memory

__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async_elem(double* dst, const double* src) { cp_async8(dst, src); }
__device__ __forceinline__ void cp_async_elem(long long* dst, const long long* src) { cp_async8(dst, src); }
__device__ __forceinline__ void cp_async_elem(int* dst, const int* src) { cp_async4(dst, src); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

/*!
 * Asynchronous copy of g0[0, count) into shared memory by the whole CTA, in 16-byte pieces where the global address
 * allows and single elements at the two ends. Element j lands at s[j + mis], where mis (returned) is the misalignment of
 * g0 in elements, so that global and shared addresses are aligned alike; `s` must be 16-byte aligned with
 * 16/sizeof(T) - 1 elements of slack. 32-bit index arithmetic throughout.
 */
template <class T>
__device__ __forceinline__ int stage_range(T* s, const T* __restrict__ g0, int count, int t) {
  constexpr int PER = 16 / int(sizeof(T));
  const int mis = int((reinterpret_cast<uintptr_t>(g0) / sizeof(T)) & (PER - 1));
  const int head = min((PER - mis) & (PER - 1), count);
  const int body = (count - head) & ~(PER - 1);
  T* sd = s + mis;
  if (t < head)
    cp_async_elem(sd + t, g0 + t);
  for (int j = head + PER * t; j < head + body; j += PER * kCsrRows)
    cp_async16(sd + j, g0 + j);
  const int tail = head + body;
  if (t < count - tail)
    cp_async_elem(sd + tail + t, g0 + tail + t);
  return mis;
}

static size_t csr_smem_bytes(int w, int b) {
  const int win = kCsrRows + 2 * b + 2;
  return size_t(kCsrChunk + 4) * 12 + size_t(w) * win * 8 + size_t(kCsrRows + 4) * 8;
}

template <int W>
__global__ void __launch_bounds__(kCsrRows) csr_apply_multi_kernel(const __grid_constant__ CsrMultiParams p) {
  extern __shared__ __align__(16) unsigned char csr_smem[];
  const int b = p.b;
  const int win = kCsrRows + 2 * b + 2;                               // per-vector stride of the x windows (even)
  double* s_val = reinterpret_cast<double*>(csr_smem);                // [kCsrChunk + 4]
  double* s_x = s_val + kCsrChunk + 4;                                // [W][win]
  long long* s_rp = reinterpret_cast<long long*>(s_x + W * win);      // [kCsrRows + 4]
  int* s_col = reinterpret_cast<int*>(s_rp + kCsrRows + 4);           // [kCsrChunk + 4]
  const int t = threadIdx.x;
  const long long nblocks = (p.n + kCsrRows - 1) / kCsrRows;
  long long blk = blockIdx.x;
  long long e0 = 0, e1 = 0;
  if (blk < nblocks) {
    const long long r0 = blk * kCsrRows, r1 = r0 + kCsrRows < p.n ? r0 + kCsrRows : p.n;
    e0 = p.row_ptr[r0];
    e1 = p.row_ptr[r1];
  }
  for (; blk < nblocks; blk += gridDim.x) {
    const long long r0 = blk * kCsrRows;
    const int nrows = int(r0 + kCsrRows < p.n ? kCsrRows : p.n - r0);
    const int rp_mis = stage_range(s_rp, p.row_ptr + r0, nrows + 1, t);
    // x window: local rows [r0 - b, r0 + nrows + b); `below`/`above` of the b rows on either side lie inside the shard
    const int below = int(r0 < b ? r0 : b);
    const int above = int(p.n - (r0 + nrows) < b ? p.n - (r0 + nrows) : b);
    const int front = (b - below + 1) & ~1; // room for the rows below the shard (first block only), kept even
    int xpos[W];                            // window position of local row l in vector k: (l - (r0 - b)) + xpos[k]
#pragma unroll
    for (int k = 0; k < W; ++k) {
      double* sx = s_x + k * win;
      const int mis = stage_range(sx + front, p.x[k] + (r0 - below), below + nrows + above, t);
      xpos[k] = front + mis - (b - below);
      // rows outside the shard: halo vectors, or zero beyond the global ends (never referenced)
      if (t < b - below)
        sx[xpos[k] + t] = p.x_lo[k] ? p.x_lo[k][(r0 - b + t) + b] : 0.0;
      if (t < b - above)
        sx[xpos[k] + b + nrows + above + t] = p.x_hi[k] ? p.x_hi[k][(r0 + nrows + above + t) - p.n] : 0.0;
    }
    // the next block's entry range, requested now and used after this block's arithmetic
    long long ne0 = 0, ne1 = 0;
    const long long nblk = blk + gridDim.x;
    if (nblk < nblocks) {
      const long long q0 = nblk * kCsrRows, q1 = q0 + kCsrRows < p.n ? q0 + kCsrRows : p.n;
      ne0 = p.row_ptr[q0];
      ne1 = p.row_ptr[q1];
    }
    double acc[W];
#pragma unroll
    for (int k = 0; k < W; ++k)
      acc[k] = 0.0;
    const int cshift = int(p.off + r0 - b); // global column of window position 0 (columns are 32-bit)
    const unsigned wn = unsigned(nrows + 2 * b);
    for (long long c0 = e0;; c0 += kCsrChunk) {
      const int cnt = int(e1 - c0 < kCsrChunk ? e1 - c0 : kCsrChunk);
      if (c0 > e0)
        __syncthreads(); // the previous round's entries have been consumed
      const double* __restrict__ sv = s_val + stage_range(s_val, p.val + c0, cnt, t);
      const int* __restrict__ sc = s_col + stage_range(s_col, p.col + c0, cnt, t);
      cp_async_wait_all();
      __syncthreads();
      if (t < nrows) {
        const long long my0 = s_rp[t + rp_mis] - c0, my1 = s_rp[t + 1 + rp_mis] - c0;
        const int lo = int(my0 > 0 ? my0 : 0), hi = int(my1 < cnt ? my1 : cnt);
#pragma unroll 3
        for (int e = lo; e < hi; ++e) {
          const double a = sv[e];
          const int wi = sc[e] - cshift;
          if (unsigned(wi) < wn) {
#pragma unroll
            for (int k = 0; k < W; ++k)
              acc[k] = __dadd_rn(acc[k], __dmul_rn(a, s_x[k * win + xpos[k] + wi]));
          } else {
            const long long c = (long long)sc[e];
#pragma unroll
            for (int k = 0; k < W; ++k)
              acc[k] = __dadd_rn(acc[k], __dmul_rn(a, x_at(c, p.off, p.n, b, p.x[k], p.x_lo[k], p.x_hi[k])));
          }
        }
      }
      if (c0 + cnt >= e1)
        break;
    }
    if (t < nrows) {
#pragma unroll
      for (int k = 0; k < W; ++k)
        p.y[k][r0 + t] = acc[k];
    }
    __syncthreads(); // shared memory is refilled by the next trip
    e0 = ne0;
    e1 = ne1;
  }
}

//! the same without shared-memory staging, for half bandwidths whose x window does not fit
template <int W>
__global__ void __launch_bounds__(256) csr_apply_multi_wide_kernel(const __grid_constant__ CsrMultiParams p) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < p.n; r += (long long)gridDim.x * blockDim.x) {
    double acc[W];
#pragma unroll
    for (int k = 0; k < W; ++k)
      acc[k] = 0.0;
    for (long long e = p.row_ptr[r]; e < p.row_ptr[r + 1]; ++e) {
      const double a = p.val[e];
      const long long c = p.col[e];
#pragma unroll
      for (int k = 0; k < W; ++k)
        if (k < p.w)
          acc[k] = __dadd_rn(acc[k], __dmul_rn(a, x_at(c, p.off, p.n, p.b, p.x[k], p.x_lo[k], p.x_hi[k])));
    }
#pragma unroll
    for (int k = 0; k < W; ++k)
      if (k < p.w)
        p.y[k][r] = acc[k];
  }
}

__global__ void __launch_bounds__(256) banded_fill_kernel(int kind, int k, long long off, long long n, double* __restrict__ out) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const long long i = off + r;
    if (kind == 0) {
      out[r] = double(i + 1);
    } else {
      const double u = __dsub_rn(__ddiv_rn(double((i * (k + 2) + k) % (2 * k + 5)), double(2 * k + 5)), 0.5);
      out[r] = kind == 1 ? u : __ddiv_rn(__dadd_rn(double(k + 1), u), double(i + 1));
    }
  }
}

__global__ void __launch_bounds__(256) target_shift_kernel(int kind, long long off, long long n, const double* __restrict__ x,
                                                           double* __restrict__ out) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const double t = kind == 1 ? 1.0 : __ddiv_rn(1.0, double(off + r + 1));
    out[r] = __dsub_rn(x[r], t);
  }
}

__global__ void __launch_bounds__(128) example_apply_kernel(long long n, const double* __restrict__ x, double* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  double acc = 0.0;
  for (long long j = 0; j < n; ++j) {
    const double mij = i == j ? double(i + 1) : __dmul_rn(0.001, double((i + j) % n));
    acc = __dadd_rn(acc, __dmul_rn(mij, x[j]));
  }
  y[i] = acc;
}

struct PActionParams {
  double* actions[ITSOLV_MAX_PANEL];
  const long long* rows; // candidate global rows (sorted, unique)
  const int* map_ptr;
  const long long* idx;
  const double* val;
  const double* pcoef; // nact x nP
  long long off, n;
  int nrows, nact, nP, b;
  double eps;
};

//! actions[k][i] += sum over P vectors p (ascending) and their entries (c,v): A(i,c) * (v * pcoef[k][p])
__global__ void p_action_kernel(const __grid_constant__ PActionParams p) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.nrows * p.nact)
    return;
  const int k = t / p.nrows;
  const long long i = p.rows[t % p.nrows];
  if (i < p.off || i >= p.off + p.n)
    return;
  double a = p.actions[k][i - p.off];
  for (int q = 0; q < p.nP; ++q) {
    const double ck = p.pcoef[size_t(k) * p.nP + q];
    for (int e = p.map_ptr[q]; e < p.map_ptr[q + 1]; ++e) {
      const long long c = p.idx[e];
      const long long d = i > c ? i - c : c - i;
      if (d <= p.b)
        a = __dadd_rn(a, __dmul_rn(band_entry(i, c, p.eps), __dmul_rn(p.val[e], ck)));
    }
  }
  p.actions[k][i - p.off] = a;
}

static int rows_grid(itsolv_ctx* ctx, size_t n) {
  return int(std::max<size_t>(1, std::min<size_t>((n + 255) / 256, size_t(ctx->num_sms) * 8)));
}

} // namespace itsolv

using namespace itsolv;

extern "C" {

int itsolv_banded_apply_f64(itsolv_ctx* ctx, int64_t n_global, int64_t row_offset, size_t n, int b, double eps,
                            const double* x, const double* x_lo, const double* x_hi, double* y) {
  ++ctx->write_epoch;
  if (n == 0)
    return 0;
  ITSOLV_REQUIRE(x != y, "itsolv_banded_apply_f64: in-place application is not supported");
  ITSOLV_REQUIRE((row_offset == 0 || x_lo) && (row_offset + int64_t(n) == n_global || x_hi),
                 "itsolv_banded_apply_f64: interior shard needs halo rows");
  banded_apply_kernel<<<rows_grid(ctx, n), 256, 0, ctx->stream>>>(n_global, row_offset, (long long)n, b, eps, x, x_lo,
                                                                  x_hi, y);
  ITSOLV_CUDA(cudaGetLastError());
  ctx->counters.launches += 1;
  return 0;
}

int itsolv_csr_apply_multi_f64(itsolv_ctx* ctx, int64_t n_global, int64_t row_offset, size_t n, int b,
                               const int64_t* row_ptr, const int32_t* col, const double* val, int w,
                               const double* const* x, const double* const* x_lo, const double* const* x_hi,
                               double* const* y);

int itsolv_csr_apply_f64(itsolv_ctx* ctx, int64_t n_global, int64_t row_offset, size_t n, int b, const int64_t* row_ptr,
                         const int32_t* col, const double* val, const double* x, const double* x_lo, const double* x_hi,
                         double* y) {
  ++ctx->write_epoch;
  if (n == 0)
    return 0;
  ITSOLV_REQUIRE(x != y, "itsolv_csr_apply_f64: in-place application is not supported");
  ITSOLV_REQUIRE((row_offset == 0 || x_lo) && (row_offset + int64_t(n) == n_global || x_hi),
                 "itsolv_csr_apply_f64: interior shard needs halo rows");
  const double* xs[1] = {x};
  const double* los[1] = {x_lo};
  const double* his[1] = {x_hi};
  double* ys[1] = {y};
  return itsolv_csr_apply_multi_f64(ctx, n_global, row_offset, n, b, row_ptr, col, val, 1, xs, los, his, ys);
}

int itsolv_csr_apply_multi_f64(itsolv_ctx* ctx, int64_t n_global, int64_t row_offset, size_t n, int b,
                               const int64_t* row_ptr, const int32_t* col, const double* val, int w,
                               const double* const* x, const double* const* x_lo, const double* const* x_hi,
                               double* const* y) {
  ++ctx->write_epoch;
  if (n == 0 || w <= 0)
    return 0;
  for (int start = 0; start < w; start += 8) {
    const int cnt = w - start < 8 ? w - start : 8;
    CsrMultiParams p;
    for (int k = 0; k < cnt; ++k) {
      p.x[k] = x[start + k];
      p.x_lo[k] = x_lo ? x_lo[start + k] : nullptr;
      p.x_hi[k] = x_hi ? x_hi[start + k] : nullptr;
      p.y[k] = y[start + k];
      ITSOLV_REQUIRE(p.x[k] != p.y[k], "itsolv_csr_apply_multi_f64: in-place application is not supported");
      ITSOLV_REQUIRE((row_offset == 0 || p.x_lo[k]) && (row_offset + int64_t(n) == n_global || p.x_hi[k]),
                     "itsolv_csr_apply_multi_f64: interior shard needs halo rows");
    }
    p.row_ptr = reinterpret_cast<const long long*>(row_ptr);
    p.col = col;
    p.val = val;
    p.off = row_offset;
    p.n = (long long)n;
    p.b = b;
    p.w = cnt;
    const int wt = cnt <= 1 ? 1 : (cnt <= 2 ? 2 : (cnt <= 4 ? 4 : 8)); // the unstaged kernel predicates on p.w
    if (b <= kCsrMaxHalfBand) {
      using Kernel = void (*)(const CsrMultiParams);
      static const Kernel kernels[8] = {csr_apply_multi_kernel<1>, csr_apply_multi_kernel<2>, csr_apply_multi_kernel<3>,
                                        csr_apply_multi_kernel<4>, csr_apply_multi_kernel<5>, csr_apply_multi_kernel<6>,
                                        csr_apply_multi_kernel<7>, csr_apply_multi_kernel<8>};
      const Kernel kernel = kernels[cnt - 1]; // the vector count is a template argument: no predication in the row loop
      const size_t smem = csr_smem_bytes(cnt, b);
      if (ensure_dynamic_smem(ctx, reinterpret_cast<const void*>(kernel), smem))
        return 1;
      const int per_sm = std::max(1, std::min(8, int(size_t(ctx->max_smem_optin) / (smem + 1024))));
      const int grid = int(std::max<size_t>(1, std::min<size_t>((n + kCsrRows - 1) / kCsrRows, size_t(ctx->num_sms) * per_sm)));
      kernel<<<grid, kCsrRows, smem, ctx->stream>>>(p);
    } else if (wt == 1) {
      csr_apply_multi_wide_kernel<1><<<rows_grid(ctx, n), 256, 0, ctx->stream>>>(p);
    } else if (wt == 2) {
      csr_apply_multi_wide_kernel<2><<<rows_grid(ctx, n), 256, 0, ctx->stream>>>(p);
    } else if (wt == 4) {
      csr_apply_multi_wide_kernel<4><<<rows_grid(ctx, n), 256, 0, ctx->stream>>>(p);
    } else {
      csr_apply_multi_wide_kernel<8><<<rows_grid(ctx, n), 256, 0, ctx->stream>>>(p);
    }
    ITSOLV_CUDA(cudaGetLastError());
    ctx->counters.launches += 1;
  }
  return 0;
}

int itsolv_banded_fill_f64(itsolv_ctx* ctx, int kind, int k, int64_t row_offset, size_t n, double* out) {
  ++ctx->write_epoch;
  if (n == 0)
    return 0;
  banded_fill_kernel<<<rows_grid(ctx, n), 256, 0, ctx->stream>>>(kind, k, row_offset, (long long)n, out);
  ITSOLV_CUDA(cudaGetLastError());
  ctx->counters.launches += 1;
  return 0;
}

int itsolv_banded_target_shift_f64(itsolv_ctx* ctx, int target_kind, int64_t row_offset, const double* x, double* out,
                                   size_t n) {
  ++ctx->write_epoch;
  if (n == 0)
    return 0;
  target_shift_kernel<<<rows_grid(ctx, n), 256, 0, ctx->stream>>>(target_kind, row_offset, (long long)n, x, out);
  ITSOLV_CUDA(cudaGetLastError());
  ctx->counters.launches += 1;
  return 0;
}

int itsolv_example_apply_f64(itsolv_ctx* ctx, size_t n, const double* x, double* y) {
  ++ctx->write_epoch;
  if (n == 0)
    return 0;
  ITSOLV_REQUIRE(x != y, "itsolv_example_apply_f64: in-place application is not supported");
  example_apply_kernel<<<int((n + 127) / 128), 128, 0, ctx->stream>>>((long long)n, x, y);
  ITSOLV_CUDA(cudaGetLastError());
  ctx->counters.launches += 1;
  return 0;
}

int itsolv_banded_p_action_f64(itsolv_ctx* ctx, int64_t n_global, int64_t row_offset, size_t n, int b, double eps,
                               int nact, double* const* actions, int nP, const int32_t* map_ptr, const int64_t* idx,
                               const double* val, const double* pcoef) {
  ++ctx->write_epoch;
  if (nact <= 0 || nP <= 0 || n == 0)
    return 0;
  ITSOLV_REQUIRE(nact <= ITSOLV_MAX_PANEL, "itsolv_banded_p_action_f64: too many action vectors");
  const int nnz = map_ptr[nP];
  std::vector<long long> rows;
  rows.reserve(size_t(nnz) * (2 * b + 1));
  for (int e = 0; e < nnz; ++e)
    for (long long i = std::max<long long>(0, idx[e] - b); i <= std::min<long long>(n_global - 1, idx[e] + b); ++i)
      if (i >= row_offset && i < row_offset + (long long)n)
        rows.push_back(i);
  std::sort(rows.begin(), rows.end());
  rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
  if (rows.empty())
    return 0;
  auto a16 = [](size_t v) { return (v + 15) & ~size_t(15); };
  const size_t off_rows = 0, off_ptr = a16(rows.size() * 8), off_idx = a16(off_ptr + size_t(nP + 1) * 4),
               off_val = a16(off_idx + size_t(nnz) * 8), off_coef = a16(off_val + size_t(nnz) * 8),
               bytes = a16(off_coef + size_t(nact) * nP * 8);
  char *h = nullptr, *d = nullptr;
  int slot = 0;
  if (stage_acquire(ctx, bytes, &h, &d, &slot))
    return 1;
  std::memcpy(h + off_rows, rows.data(), rows.size() * 8);
  std::memcpy(h + off_ptr, map_ptr, size_t(nP + 1) * 4);
  std::memcpy(h + off_idx, idx, size_t(nnz) * 8);
  std::memcpy(h + off_val, val, size_t(nnz) * 8);
  std::memcpy(h + off_coef, pcoef, size_t(nact) * nP * 8);
  if (stage_commit(ctx, slot, bytes))
    return 1;
  PActionParams p;
  for (int k = 0; k < nact; ++k)
    p.actions[k] = actions[k];
  p.rows = reinterpret_cast<const long long*>(d + off_rows);
  p.map_ptr = reinterpret_cast<const int*>(d + off_ptr);
  p.idx = reinterpret_cast<const long long*>(d + off_idx);
  p.val = reinterpret_cast<const double*>(d + off_val);
  p.pcoef = reinterpret_cast<const double*>(d + off_coef);
  p.off = row_offset;
  p.n = (long long)n;
  p.nrows = int(rows.size());
  p.nact = nact;
  p.nP = nP;
  p.b = b;
  p.eps = eps;
  const int total = p.nrows * nact;
  p_action_kernel<<<(total + 127) / 128, 128, 0, ctx->stream>>>(p);
  ITSOLV_CUDA(cudaGetLastError());
  ctx->counters.launches += 1;
  return stage_done(ctx, slot);
}

} // extern "C"
