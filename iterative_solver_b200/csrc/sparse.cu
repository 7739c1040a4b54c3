// Dense x sparse (P-space) operations: the P vectors are std::map<size_t,double> on the host
// (reference src/molpro/linalg/array/ArrayHandlerIterableSparse.h:35-63, ArrayHandlerDistrSparse.h:30-65,
// array/util/gemm.h:207-253). They are packed CSR-like, staged to the device through the pinned ring and applied with
// gather / scatter kernels. Work is O(nnz) (a few hundred entries); the point is that the dense vectors never leave HBM.
// Arithmetic follows the reference term by term (product rounded, then added, entries in map order), so on one GPU the
// results are bit-identical to the CPU handler.
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace itsolv {

int finish_result(itsolv_ctx* ctx, int count, double* out, bool host_direct); // gemm_inner.cu

struct SparseInnerParams {
  const double* x[ITSOLV_MAX_PANEL];
  const int* map_ptr;
  const long long* idx;
  const double* val;
  double* out; // k x nmap
  long long lo, hi; // local index range [lo, hi)
  int k, nmap;
};

__global__ void sparse_scatter_kernel(double* __restrict__ x, long long lo, long long hi, int nnz,
                                      const long long* __restrict__ idx, const double* __restrict__ val) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nnz && idx[e] >= lo && idx[e] < hi)
    x[idx[e] - lo] = val[e];
}

__global__ void sparse_inner_kernel(const __grid_constant__ SparseInnerParams p) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.k * p.nmap)
    return;
  const int i = t / p.nmap, j = t % p.nmap;
  const double* __restrict__ x = p.x[i];
  double tot = 0.0;
  for (int e = p.map_ptr[j]; e < p.map_ptr[j + 1]; ++e) {
    const long long g = p.idx[e];
    if (g >= p.lo && g < p.hi)
      tot = __dadd_rn(tot, __dmul_rn(x[g - p.lo], p.val[e]));
  }
  p.out[t] = tot;
}

struct SparseOuterParams {
  double* y[ITSOLV_MAX_PANEL];
  const double* alpha; // nmap x ndense
  const int* map_ptr;
  const long long* idx;
  const double* val;
  long long lo, hi;
  int nmap, ndense, nnz;
};

//! every index occurs once over all maps: one thread per (entry, dense vector)
__global__ void sparse_outer_unique_kernel(const __grid_constant__ SparseOuterParams p, const int* __restrict__ entry_map) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.nnz * p.ndense)
    return;
  const int e = t / p.ndense, j = t % p.ndense;
  const long long g = p.idx[e];
  if (g < p.lo || g >= p.hi)
    return;
  double* y = p.y[j];
  y[g - p.lo] = __dadd_rn(y[g - p.lo], __dmul_rn(p.alpha[size_t(entry_map[e]) * p.ndense + j], p.val[e]));
}

//! general case: one thread per dense vector walks maps and entries in the reference's order
__global__ void sparse_outer_serial_kernel(const __grid_constant__ SparseOuterParams p) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p.ndense)
    return;
  double* y = p.y[j];
  for (int i = 0; i < p.nmap; ++i) {
    const double a = p.alpha[size_t(i) * p.ndense + j];
    for (int e = p.map_ptr[i]; e < p.map_ptr[i + 1]; ++e) {
      const long long g = p.idx[e];
      if (g >= p.lo && g < p.hi)
        y[g - p.lo] = __dadd_rn(y[g - p.lo], __dmul_rn(a, p.val[e]));
    }
  }
}

//! host-side packer into one staging slot; returns device addresses of the pieces
struct Packed {
  int slot = 0;
  char* d = nullptr;
  size_t off_alpha = 0, off_ptr = 0, off_idx = 0, off_val = 0, off_emap = 0, bytes = 0;
};

static size_t align16(size_t v) { return (v + 15) & ~size_t(15); }

static int pack_maps(itsolv_ctx* ctx, const double* alpha, size_t nalpha, int nmap, const int32_t* map_ptr,
                     const int64_t* idx, const double* val, bool with_entry_map, Packed& pk) {
  const int base = map_ptr[0];
  const int nnz = map_ptr[nmap] - base;
  pk.off_alpha = 0;
  pk.off_ptr = align16(nalpha * 8);
  pk.off_idx = align16(pk.off_ptr + size_t(nmap + 1) * 4);
  pk.off_val = align16(pk.off_idx + size_t(nnz) * 8);
  pk.off_emap = align16(pk.off_val + size_t(nnz) * 8);
  pk.bytes = align16(pk.off_emap + (with_entry_map ? size_t(nnz) * 4 : 0));
  char* h = nullptr;
  if (stage_acquire(ctx, pk.bytes, &h, &pk.d, &pk.slot))
    return 1;
  if (nalpha)
    std::memcpy(h + pk.off_alpha, alpha, nalpha * 8);
  int* hp = reinterpret_cast<int*>(h + pk.off_ptr);
  for (int j = 0; j <= nmap; ++j)
    hp[j] = map_ptr[j] - base;
  std::memcpy(h + pk.off_idx, idx + base, size_t(nnz) * 8);
  std::memcpy(h + pk.off_val, val + base, size_t(nnz) * 8);
  if (with_entry_map) {
    int* em = reinterpret_cast<int*>(h + pk.off_emap);
    for (int j = 0; j < nmap; ++j)
      for (int e = hp[j]; e < hp[j + 1]; ++e)
        em[e] = j;
  }
  return stage_commit(ctx, pk.slot, pk.bytes);
}

} // namespace itsolv

using namespace itsolv;

extern "C" {

int itsolv_sparse_copy_f64(itsolv_ctx* ctx, double* x, size_t n, size_t global_offset, int nnz, const int64_t* idx,
                           const double* val) {
  ++ctx->write_epoch;
  ctx->counters.n_sparse++;
  if (itsolv_fill_f64(ctx, 0.0, x, n))
    return 1;
  ctx->counters.n_fill--;
  ++ctx->write_epoch; // before any early return: every rank advances alike
  if (nnz <= 0 || n == 0)
    return 0;
  const int32_t ptr[2] = {0, nnz};
  Packed pk;
  if (pack_maps(ctx, nullptr, 0, 1, ptr, idx, val, false, pk))
    return 1;
  sparse_scatter_kernel<<<(nnz + 127) / 128, 128, 0, ctx->stream>>>(
      x, (long long)global_offset, (long long)(global_offset + n), nnz,
      reinterpret_cast<const long long*>(pk.d + pk.off_idx), reinterpret_cast<const double*>(pk.d + pk.off_val));
  ITSOLV_CUDA(cudaGetLastError());
  ctx->counters.launches += 1;
  return stage_done(ctx, pk.slot);
}

int itsolv_sparse_gemm_inner_f64(itsolv_ctx* ctx, const double* const* xx, int k, size_t n, size_t global_offset,
                                 int nmap, const int32_t* map_ptr, const int64_t* idx, const double* val, double* out) {
  ctx->counters.n_sparse++;
  if (k <= 0 || nmap <= 0)
    return 0;
  const int max_maps = std::max(1, (ITSOLV_MAX_PANEL * ITSOLV_MAX_PANEL) / std::max(k, 1));
  for (int i0 = 0; i0 < k; i0 += ITSOLV_MAX_PANEL) {
    const int kb = std::min(ITSOLV_MAX_PANEL, k - i0);
    for (int j0 = 0; j0 < nmap;) {
      int mb = std::min(max_maps, nmap - j0);
      // keep the packed chunk inside one staging slot
      while (mb > 1 && size_t(map_ptr[j0 + mb] - map_ptr[j0]) * 16 + size_t(mb) * 4 + 4096 > ctx->stage_slot_bytes)
        mb /= 2;
      Packed pk;
      if (pack_maps(ctx, nullptr, 0, mb, map_ptr + j0, idx, val, false, pk))
        return 1;
      SparseInnerParams p;
      for (int i = 0; i < kb; ++i)
        p.x[i] = xx[i0 + i];
      p.map_ptr = reinterpret_cast<const int*>(pk.d + pk.off_ptr);
      p.idx = reinterpret_cast<const long long*>(pk.d + pk.off_idx);
      p.val = reinterpret_cast<const double*>(pk.d + pk.off_val);
      p.out = ctx->d_result;
      p.lo = (long long)global_offset;
      p.hi = (long long)(global_offset + n);
      p.k = kb;
      p.nmap = mb;
      const int total = kb * mb;
      sparse_inner_kernel<<<(total + 127) / 128, 128, 0, ctx->stream>>>(p);
      ITSOLV_CUDA(cudaGetLastError());
      ctx->counters.launches += 1;
      if (stage_done(ctx, pk.slot))
        return 1;
      std::vector<double> block(static_cast<size_t>(total), 0.0);
      if (finish_result(ctx, total, block.data(), false))
        return 1;
      for (int i = 0; i < kb; ++i)
        for (int j = 0; j < mb; ++j)
          out[size_t(i0 + i) * nmap + (j0 + j)] = block[size_t(i) * mb + j];
      j0 += mb;
    }
  }
  return 0;
}

int itsolv_sparse_gemm_outer_f64(itsolv_ctx* ctx, const double* alpha, int nmap, int ndense, const int32_t* map_ptr,
                                 const int64_t* idx, const double* val, double* const* yy, size_t n,
                                 size_t global_offset) {
  ++ctx->write_epoch;
  ctx->counters.n_sparse++;
  ++ctx->write_epoch; // before any early return: every rank advances alike
  if (nmap <= 0 || ndense <= 0 || n == 0)
    return 0;
  // the parallel kernel needs every (index) to be touched by one entry only
  bool unique = true;
  {
    std::vector<int64_t> all(idx + map_ptr[0], idx + map_ptr[nmap]);
    std::sort(all.begin(), all.end());
    unique = std::adjacent_find(all.begin(), all.end()) == all.end();
  }
  for (int j0 = 0; j0 < ndense; j0 += ITSOLV_MAX_PANEL) {
    const int mb = std::min(ITSOLV_MAX_PANEL, ndense - j0);
    for (int i0 = 0; i0 < nmap;) {
      int kb = std::min(nmap - i0, std::max(1, (ITSOLV_MAX_PANEL * ITSOLV_MAX_PANEL) / mb));
      while (kb > 1 &&
             size_t(map_ptr[i0 + kb] - map_ptr[i0]) * 20 + size_t(kb) * (4 + 8 * size_t(mb)) + 4096 > ctx->stage_slot_bytes)
        kb /= 2;
      std::vector<double> a(size_t(kb) * mb);
      for (int i = 0; i < kb; ++i)
        for (int j = 0; j < mb; ++j)
          a[size_t(i) * mb + j] = alpha[size_t(i0 + i) * ndense + (j0 + j)];
      Packed pk;
      if (pack_maps(ctx, a.data(), a.size(), kb, map_ptr + i0, idx, val, true, pk))
        return 1;
      SparseOuterParams p;
      for (int j = 0; j < mb; ++j)
        p.y[j] = yy[j0 + j];
      p.alpha = reinterpret_cast<const double*>(pk.d + pk.off_alpha);
      p.map_ptr = reinterpret_cast<const int*>(pk.d + pk.off_ptr);
      p.idx = reinterpret_cast<const long long*>(pk.d + pk.off_idx);
      p.val = reinterpret_cast<const double*>(pk.d + pk.off_val);
      p.lo = (long long)global_offset;
      p.hi = (long long)(global_offset + n);
      p.nmap = kb;
      p.ndense = mb;
      p.nnz = map_ptr[i0 + kb] - map_ptr[i0];
      if (p.nnz > 0) {
        if (unique) {
          const int total = p.nnz * mb;
          sparse_outer_unique_kernel<<<(total + 127) / 128, 128, 0, ctx->stream>>>(
              p, reinterpret_cast<const int*>(pk.d + pk.off_emap));
        } else {
          sparse_outer_serial_kernel<<<(mb + 31) / 32, 32, 0, ctx->stream>>>(p);
        }
        ITSOLV_CUDA(cudaGetLastError());
        ctx->counters.launches += 1;
      }
      if (stage_done(ctx, pk.slot))
        return 1;
      i0 += kb;
    }
  }
  return 0;
}

} // extern "C"
