// gemm_inner for wide panels (k*m >= 256, e.g. the 16 x q blocks of a 16-root Davidson run): FP64 tensor cores.
//
// Why tensor cores here and nowhere else (BASELINE.json north_star: "tensor cores are used only if ncu shows a panel
// shape is compute-bound"): with FMA register tiles (gemm_inner.cu, 4 x 4 accumulators per thread) the 16 x 24 ... 16 x 64
// panels reach 0.62-0.70 of the HBM copy rate with the FP64 pipe at ~35 % — the kernel is bound by issue slots and by the
// shared-memory operand loads (8 LDS.128 per 32 DFMA), not by HBM (profiles/notes_r01.md). mma.sync.m8n8k4.f64 (SASS DMMA)
// computes an 8 x 8 x 4 block per warp instruction from ONE A and ONE B register per lane: 10 LDS.64 feed 16 MMAs
// (= 4096 FMAs) for a 16 x 64 panel, 6x fewer shared-memory loads per FMA and 8x fewer issue slots.
// tcgen05 has no FP64 kind, so the warp-level DMMA is the FP64 tensor path on sm_100a.
//
// Structure: the same TMA + mbarrier tile pipeline as gemm_inner.cu (gi_pipeline.cuh), one CTA per SM, up to 8 producer
// warps, 8 consumer warps. A consumer warp owns NI x NJ MMA tiles (8 x 8 outputs each) = a block of the k x m output,
// and a share of the rows of every tile (row group); warps that share an output block are added in warp order at the
// end, then the CTA partial sums are finished as everywhere else (CTA order, deterministic).
// The reduction dimension of the MMA is the row index: A(i, r) = x_i[r], B(r, j) = y_j[r], four rows per instruction.
#include <algorithm>

#include "common.cuh"
#include "gi_finalize.cuh"
#include "gi_pipeline.cuh"

namespace itsolv {

void fill_finalize(itsolv_ctx* ctx, int grid_bound, int km, GiFinalize* f, bool* host_direct); // gemm_inner.cu
int launch_reduce_partials(itsolv_ctx* ctx, int grid, int km);                             // gemm_inner.cu
int finish_with_peers(itsolv_ctx* ctx, int km, bool* host_direct);                         // gemm_inner.cu

constexpr int kMmaConsumerWarps = 8;

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

/*!
 * GiParams reuse: KB = WO (number of output blocks side by side along j), MB = WR (row groups), G unused.
 * Warp w: output block wo = w % WO (tile columns [wo*NJ, wo*NJ+NJ)), row group wr = w / WO.
 */
template <int NI, int NJ>
__global__ void __launch_bounds__(32 * kMmaConsumerWarps + 32 * kMaxProducerWarps, 1)
    gemm_inner_mma_kernel(const __grid_constant__ GiParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* tiles = reinterpret_cast<double*>(smem_raw);
  __shared__ uint64_t full_bar[kMaxStages];
  __shared__ uint64_t empty_bar[kMaxStages];
  __shared__ int s_is_last;

  const int tid = threadIdx.x;
  const int nconsumers = 32 * kMmaConsumerWarps;
  const bool is_producer = tid >= nconsumers;
  const int warp = tid >> 5, lane = tid & 31;
  const int WO = p.KB, WR = p.MB;
  const int wo = warp % WO, wr = warp / WO;
  const int grp = lane >> 2, tig = lane & 3; // MMA fragment coordinates of this lane

  // shared-memory offsets (in doubles) of the vectors this lane feeds into its A and B fragments
  int aoff[NI], boff[NJ];
#pragma unroll
  for (int a = 0; a < NI; ++a) {
    const int i = a * 8 + grp;
    aoff[a] = int(p.xslot[i < p.k ? i : 0]) * p.stride + tig;
  }
#pragma unroll
  for (int b = 0; b < NJ; ++b) {
    const int j = (wo * NJ + b) * 8 + grp;
    boff[b] = int(p.yslot[j < p.m ? j : 0]) * p.stride + tig;
  }
  double acc[NI][NJ][2];
#pragma unroll
  for (int a = 0; a < NI; ++a)
#pragma unroll
    for (int b = 0; b < NJ; ++b)
      acc[a][b][0] = acc[a][b][1] = 0.0;

  const long long my_tiles =
      p.nfull > (long long)blockIdx.x ? (p.nfull - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const size_t stage_doubles = size_t(p.nvec) * p.stride;
  const int ksteps = p.rows / 4;

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], uint32_t(p.nprod));
      mbar_init(&empty_bar[s], kMmaConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto consume = [&](const double* __restrict__ st) {
    if (wr >= WR)
      return;
#pragma unroll 2
    for (int ks = wr; ks < ksteps; ks += WR) {
      double af[NI], bf[NJ];
#pragma unroll
      for (int a = 0; a < NI; ++a)
        af[a] = st[aoff[a] + 4 * ks];
#pragma unroll
      for (int b = 0; b < NJ; ++b)
        bf[b] = st[boff[b] + 4 * ks];
#pragma unroll
      for (int a = 0; a < NI; ++a)
#pragma unroll
        for (int b = 0; b < NJ; ++b)
          dmma_m8n8k4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
  };

  if (is_producer) {
    gi_tma_producer(p, tiles, full_bar, empty_bar, nconsumers, my_tiles, stage_doubles);
  } else {
    for (long long s = 0; s < my_tiles; ++s) {
      const int stage = int(s % p.stages);
      mbar_wait(&full_bar[stage], uint32_t((s / p.stages) & 1));
      consume(tiles + size_t(stage) * stage_doubles);
      __syncwarp();
      if (lane == 0)
        mbar_arrive(&empty_bar[stage]);
    }
  }
  __syncthreads();

  // the last, partial tile (zero-filled beyond the end of the vectors) belongs to the CTA that would own tile nfull
  const size_t tail0 = size_t(p.nfull) * size_t(p.rows);
  if (tail0 < p.n && int(p.nfull % gridDim.x) == int(blockIdx.x)) {
    cooperative_fill(p, tiles, tail0, int(p.n - tail0));
    __syncthreads();
    if (!is_producer)
      consume(tiles);
    __syncthreads();
  }

  // fold the row groups in warp order: red[wr][wo][a][b][lane][2]
  double* red = tiles;
  if (!is_producer && wr < WR) {
    double* mine = red + (size_t(wr) * WO + wo) * (NI * NJ * 64);
#pragma unroll
    for (int a = 0; a < NI; ++a)
#pragma unroll
      for (int b = 0; b < NJ; ++b) {
        mine[(a * NJ + b) * 64 + lane * 2 + 0] = acc[a][b][0];
        mine[(a * NJ + b) * 64 + lane * 2 + 1] = acc[a][b][1];
      }
  }
  __syncthreads();
  const int km = p.k * p.m;
  double* out = p.fin.partials + size_t(blockIdx.x) * km;
  for (int e = tid; e < km; e += blockDim.x) {
    const int i = e / p.m, j = e % p.m;
    const int a = i >> 3, g = i & 7;         // A tile row, row inside the tile (= fragment group)
    const int tj = j >> 3, c = j & 7;        // tile column, column inside the tile
    const int o = tj / NJ, b = tj % NJ;      // output block, tile inside the block
    const int ln = g * 4 + (c >> 1), reg = c & 1;
    const double* src = red + size_t(o) * (NI * NJ * 64) + (a * NJ + b) * 64 + ln * 2 + reg;
    double sum = 0.0;
    for (int r = 0; r < WR; ++r)
      sum += src[size_t(r) * WO * (NI * NJ * 64)];
    out[e] = sum;
  }
  gi_finalize(p.fin, km, &s_is_last);
}

using MmaKernel = void (*)(const GiParams);

struct MmaShape {
  int ni, nj;
  MmaKernel kernel;
};

static const MmaShape* mma_shapes(int* count) {
  static const MmaShape shapes[] = {
      // at most 16 tiles (32 accumulator doubles per lane) per warp: 512 threads per CTA leave 128 registers per thread
      {1, 8, gemm_inner_mma_kernel<1, 8>}, {1, 16, gemm_inner_mma_kernel<1, 16>}, {2, 4, gemm_inner_mma_kernel<2, 4>},
      {2, 8, gemm_inner_mma_kernel<2, 8>}, {4, 4, gemm_inner_mma_kernel<4, 4>},   {8, 2, gemm_inner_mma_kernel<8, 2>},
      {16, 1, gemm_inner_mma_kernel<16, 1>},
  };
  *count = int(sizeof(shapes) / sizeof(shapes[0]));
  return shapes;
}

/*!
 * Tensor-core path of gemm_inner. *handled = false leaves the call to the FMA kernels (small panels, vectors that are
 * not 16-byte aligned, more than 256 distinct vectors cannot happen).
 */
int gemm_inner_mma_device(itsolv_ctx* ctx, const double* const* xx, int k, const double* const* yy, int m, size_t n,
                          bool* host_direct, bool* handled) {
  *handled = false;
  const int min_outputs = ctx->opt_gi_mma > 0 ? ctx->opt_gi_mma : 256; // 16 x 16 and wider
  if (ctx->opt_gi_mma < 0 || k * m < min_outputs || n < 4096)
    return 0;
  GiParams p;
  p.nvec = 0;
  auto slot_of = [&](const double* ptr) {
    for (int v = 0; v < p.nvec; ++v)
      if (p.vec[v] == ptr)
        return v;
    p.vec[p.nvec] = ptr;
    return p.nvec++;
  };
  for (int i = 0; i < ITSOLV_MAX_PANEL; ++i)
    p.xslot[i] = p.yslot[i] = 0;
  for (int i = 0; i < k; ++i) {
    if (!xx[i] || !aligned16(xx[i]))
      return 0;
    p.xslot[i] = (unsigned char)slot_of(xx[i]);
  }
  for (int j = 0; j < m; ++j) {
    if (!yy[j] || !aligned16(yy[j]))
      return 0;
    p.yslot[j] = (unsigned char)slot_of(yy[j]);
  }
  p.k = k;
  p.m = m;
  p.n = n;
  // smallest warp block (NI x NJ tiles of 8 x 8) that covers the panel with at most 8 output blocks side by side
  const int ti = (k + 7) / 8, tj = (m + 7) / 8;
  int count = 0;
  const MmaShape* shapes = mma_shapes(&count);
  const MmaShape* best = nullptr;
  int best_wo = 0;
  for (int c = 0; c < count; ++c) {
    const MmaShape& s = shapes[c];
    if (s.ni < ti)
      continue;
    int wo = (tj + s.nj - 1) / s.nj;
    int wo2 = 1;
    while (wo2 < wo)
      wo2 *= 2;
    if (wo2 > kMmaConsumerWarps)
      continue;
    // fewer padded tiles first, then fewer output blocks (more row groups per block)
    const int cost = s.ni * s.nj * wo2;
    if (!best || cost < best->ni * best->nj * best_wo || (cost == best->ni * best->nj * best_wo && wo2 < best_wo)) {
      best = &s;
      best_wo = wo2;
    }
  }
  if (!best)
    return 0;
  p.KB = best_wo;                     // WO
  p.MB = kMmaConsumerWarps / best_wo; // WR
  p.G = 1;
  p.nprod = std::min(kMaxProducerWarps, p.nvec);
  if (ctx->opt_gi_nprod > 0)
    p.nprod = std::max(1, std::min(ctx->opt_gi_nprod, p.nprod));

  const size_t reduce_bytes = size_t(kMmaConsumerWarps) * best->ni * best->nj * 64 * sizeof(double);
  const size_t smem_cap = size_t(ctx->max_smem_optin) - 2048;
  if (reduce_bytes > smem_cap)
    return 0;
  // copies shrink with the number of vectors (one per vector per tile); beyond 64 vectors two deeper stages beat three
  int stages = ctx->opt_gi_stages > 0 ? ctx->opt_gi_stages : (p.nvec > 64 ? 2 : 3);
  stages = std::max(1, std::min(stages, kMaxStages));
  int rows = ctx->opt_gi_rows > 0 ? ctx->opt_gi_rows
                                  : int(std::min<size_t>(smem_cap / (size_t(stages) * p.nvec * sizeof(double)), 1024));
  rows = std::max(16, (rows / 16) * 16);
  // stride = rows + 4: consecutive vectors start 32 bytes apart modulo 128, so the 8 vectors x 4 rows that one fragment
  // load touches fall into distinct banks
  while (size_t(stages) * p.nvec * (rows + 4) * sizeof(double) > smem_cap) {
    if (rows > 16)
      rows -= 16;
    else if (stages > 1)
      --stages;
    else
      return 0;
  }
  p.rows = rows;
  p.stride = rows + 4;
  p.stages = stages;
  p.chunk_rows = 1 << 20;
  p.nfull = (long long)(n / size_t(rows));
  const size_t smem_bytes = std::max(size_t(stages) * p.nvec * p.stride * sizeof(double), reduce_bytes);
  const long long total_tiles = p.nfull + ((n % size_t(rows)) ? 1 : 0);
  const int grid = int(std::min<long long>(total_tiles, (long long)ctx->num_sms));
  const int km = k * m;
  if (ensure_partials(ctx, size_t(grid) * km))
    return 1;
  fill_finalize(ctx, ctx->num_sms, km, &p.fin, host_direct);
  if (ensure_dynamic_smem(ctx, reinterpret_cast<const void*>(best->kernel), smem_bytes))
    return 1;
  mark_launch(ctx);
  best->kernel<<<grid, 32 * kMmaConsumerWarps + 32 * p.nprod, smem_bytes, ctx->stream>>>(p);
  ITSOLV_CUDA(cudaGetLastError());
  ctx->counters.launches += 1;
  if (!p.fin.fused) {
    if (launch_reduce_partials(ctx, grid, km))
      return 1;
    if (finish_with_peers(ctx, km, host_direct))
      return 1;
  }
  *handled = true;
  return 0;
}

} // namespace itsolv
