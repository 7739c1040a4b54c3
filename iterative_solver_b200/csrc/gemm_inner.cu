// gemm_inner: the tall-skinny FP64 contraction  M(i,j) = sum_r x_i[r] * y_j[r]   (k x m, k,m <= 128)
// that builds the overlap / action / rhs blocks of the subspace problem
// (contract: reference src/molpro/linalg/array/ArrayHandler.h:200; CPU path: array/util/gemm.h:157-184, 268-279, which
// runs k*m separate std::inner_product sweeps, i.e. 2km vector passes; here every HBM byte is read once and feeds all
// k*m dot products).
//
// Kernel structure (sm_100a):
//   * the k+m vectors are separate allocations (reference itsolv/subspace/QSpace.h:157), so a "tile" is T rows of each
//     distinct vector; one producer warp moves a tile into shared memory with one 1-D TMA bulk copy per vector
//     (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP), S stages deep, signalled through full/empty mbarriers;
//   * consumer threads own a TI x TJ register tile of the k x m accumulators (strided assignment i = ib + a*KB,
//     j = jb + b*MB so that neighbouring lanes read neighbouring vectors -> conflict-free, broadcast LDS.128) and,
//     when the output needs fewer than all threads, split the rows of the tile between G row groups;
//   * persistent grid, tile t -> CTA t mod grid (static, so the summation order is fixed for a given shape and n);
//   * per-CTA partial k x m sums go to a workspace, a second small kernel adds them in CTA order: run-to-run
//     reproducible, no FP64 atomics.
// Vectors that appear on both sides (overlap(x,x), dot(x,x)) are loaded once.
#include <algorithm>

#include "common.cuh"
#include "gi_finalize.cuh"
#include "gi_pipeline.cuh"

namespace itsolv {

template <int TI, int TJ>
__device__ __forceinline__ void consume_tile(const double* __restrict__ st, int npairs, int g, int G, const int (&xoff)[TI],
                                             const int (&yoff)[TJ], double (&acc)[TI][TJ]) {
#pragma unroll 2
  for (int rp = g; rp < npairs; rp += G) {
    double2 xv[TI], yv[TJ];
#pragma unroll
    for (int a = 0; a < TI; ++a)
      xv[a] = *reinterpret_cast<const double2*>(st + xoff[a] + 2 * rp);
#pragma unroll
    for (int b = 0; b < TJ; ++b)
      yv[b] = *reinterpret_cast<const double2*>(st + yoff[b] + 2 * rp);
#pragma unroll
    for (int a = 0; a < TI; ++a)
#pragma unroll
      for (int b = 0; b < TJ; ++b) {
        acc[a][b] = fma(xv[a].x, yv[b].x, acc[a][b]);
        acc[a][b] = fma(xv[a].y, yv[b].y, acc[a][b]);
      }
  }
}

// Thread budget: BIG kernels run one CTA per SM (large tiles for panels of many vectors, up to 4 producer warps);
// the others run two CTAs per SM.
template <int TI, int TJ, bool BIG>
struct GiShape {
  static constexpr bool heavy = TI * TJ >= 32; // 32 or 64 accumulators per thread
  static constexpr int max_consumers = 256;
  static constexpr int max_producers = BIG ? (heavy ? 4 : kMaxProducerWarps) : 2;
  static constexpr int max_threads = max_consumers + 32 * max_producers;
  static constexpr int min_ctas = BIG ? 1 : 2;
};

template <int TI, int TJ, int LOADER, bool BIG>
__global__ void __launch_bounds__(GiShape<TI, TJ, BIG>::max_threads, GiShape<TI, TJ, BIG>::min_ctas)
    gemm_inner_kernel(const __grid_constant__ GiParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* tiles = reinterpret_cast<double*>(smem_raw);
  __shared__ uint64_t full_bar[kMaxStages];
  __shared__ uint64_t empty_bar[kMaxStages];
  __shared__ int s_is_last;

  const int tid = threadIdx.x;
  const int nconsumers = blockDim.x - 32 * p.nprod; // the last nprod warps are producers
  const int nconsumer_warps = nconsumers / 32;
  const bool is_producer = tid >= nconsumers;
  const int NT = p.KB * p.MB;
  const bool active = !is_producer && tid < NT * p.G;
  const int tile = tid % NT, g = tid / NT;
  const int ib = tile / p.MB, jb = tile % p.MB;

  int xoff[TI], yoff[TJ];
#pragma unroll
  for (int a = 0; a < TI; ++a) {
    const int i = ib + a * p.KB;
    xoff[a] = int(p.xslot[i < p.k ? i : 0]) * p.stride;
  }
#pragma unroll
  for (int b = 0; b < TJ; ++b) {
    const int j = jb + b * p.MB;
    yoff[b] = int(p.yslot[j < p.m ? j : 0]) * p.stride;
  }
  double acc[TI][TJ];
#pragma unroll
  for (int a = 0; a < TI; ++a)
#pragma unroll
    for (int b = 0; b < TJ; ++b)
      acc[a][b] = 0.0;

  const long long my_tiles =
      p.nfull > (long long)blockIdx.x ? (p.nfull - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const size_t stage_doubles = size_t(p.nvec) * p.stride;
  const int npairs = p.rows / 2;

  if (tid == 0) {
    // full: TMA -> one arrive.expect_tx per producer warp (+ bytes); cp.async -> one deferred arrive per producer thread
    const uint32_t full_count = LOADER == LOAD_TMA ? uint32_t(p.nprod) : uint32_t(32 * p.nprod);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], full_count);
      mbar_init(&empty_bar[s], nconsumer_warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if constexpr (LOADER == LOAD_TMA) {
    if (is_producer)
      gi_tma_producer(p, tiles, full_bar, empty_bar, nconsumers, my_tiles, stage_doubles);
  } else {
    if (is_producer) {
      // Every producer lane moves 16-byte (or 8-byte) pieces: piece q of a tile is rows [q % PV * W, +W) of vector q / PV,
      // so a warp instruction covers 512 contiguous bytes of one vector. Rows past the end of the vectors are
      // zero-filled by the copy itself (src-size operand), which also takes care of the last, partial tile.
      constexpr int PB = LOADER == LOAD_CPASYNC16 ? 16 : 8; // bytes per piece
      constexpr int W = PB / 8;                             // rows per piece
      const int pt = tid - nconsumers;                      // producer thread index
      const int NP = 32 * p.nprod;
      const int PV = p.rows / W;                            // pieces per vector per tile
      const int total = p.nvec * PV;
      const uint32_t tiles_u32 = smem_u32(tiles);
      const long long ntiles_all = p.nfull + ((size_t(p.nfull) * size_t(p.rows) < p.n) ? 1 : 0);
      const long long mine = ntiles_all > (long long)blockIdx.x ? (ntiles_all - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      for (long long s = 0; s < mine; ++s) {
        const int stage = int(s % p.stages);
        const long long use = s / p.stages;
        if (use > 0)
          mbar_wait(&empty_bar[stage], uint32_t((use - 1) & 1));
        const size_t row0 = size_t(blockIdx.x + s * gridDim.x) * size_t(p.rows);
        const size_t left = p.n - row0; // rows of the vectors from row0 to their end (>= 1)
        const uint32_t st = tiles_u32 + uint32_t(size_t(stage) * stage_doubles * 8);
        int v = pt / PV, r = pt % PV;
#pragma unroll 4
        for (int q = pt; q < total; q += NP) {
          const size_t row = size_t(r) * W;
          const long long valid = (long long)left - (long long)row; // rows available from this piece on
          const uint32_t nbytes = valid >= W ? uint32_t(PB) : (valid > 0 ? uint32_t(valid) * 8u : 0u);
          const double* src = p.vec[v] + row0 + (valid > 0 ? row : 0);
          cp_async_zfill<PB>(st + (uint32_t(v) * uint32_t(p.stride) + uint32_t(row)) * 8u, src, nbytes);
          r += NP;
          while (r >= PV) {
            r -= PV;
            ++v;
          }
        }
        cp_async_arrive(&full_bar[stage]);
      }
    }
  }

  if (!is_producer) {
    const int lane = tid & 31;
    // the cp.async loaders also bring the last, partial tile (zero-filled); the TMA loader leaves it to the code below
    long long ntiles_c = my_tiles;
    if constexpr (LOADER != LOAD_TMA) {
      const long long ntiles_all = p.nfull + ((size_t(p.nfull) * size_t(p.rows) < p.n) ? 1 : 0);
      ntiles_c = ntiles_all > (long long)blockIdx.x ? (ntiles_all - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    }
    for (long long s = 0; s < ntiles_c; ++s) {
      const int stage = int(s % p.stages);
      mbar_wait(&full_bar[stage], uint32_t((s / p.stages) & 1));
      if (active)
        consume_tile<TI, TJ>(tiles + size_t(stage) * stage_doubles, npairs, g, p.G, xoff, yoff, acc);
      __syncwarp();
      if (lane == 0)
        mbar_arrive(&empty_bar[stage]);
    }
  }
  __syncthreads();

  if constexpr (LOADER == LOAD_TMA) {
    // the last, partial tile belongs to the CTA that would own tile number nfull
    const size_t tail0 = size_t(p.nfull) * size_t(p.rows);
    if (tail0 < p.n && int(p.nfull % gridDim.x) == int(blockIdx.x)) {
      cooperative_fill(p, tiles, tail0, int(p.n - tail0));
      __syncthreads();
      if (active)
        consume_tile<TI, TJ>(tiles, npairs, g, p.G, xoff, yoff, acc);
      __syncthreads();
    }
  }

  // reduce the G row groups in group order; red[(a*TJ+b)][g][tile]
  double* red = tiles;
  const int nact = NT * p.G;
  if (active) {
#pragma unroll
    for (int a = 0; a < TI; ++a)
#pragma unroll
      for (int b = 0; b < TJ; ++b)
        red[size_t(a * TJ + b) * nact + tid] = acc[a][b];
  }
  __syncthreads();
  const int km = p.k * p.m;
  double* out = p.fin.partials + size_t(blockIdx.x) * km;
  {
    // one warp per output element: lanes take the row groups in turn, then a fixed shuffle tree
    const int warp_id = tid >> 5, lane_id = tid & 31, nwarps_cta = blockDim.x >> 5;
    for (int e = warp_id; e < km; e += nwarps_cta) {
      const int i = e / p.m, j = e % p.m;
      const int a = i / p.KB, tb_i = i % p.KB;
      const int b = j / p.MB, tb_j = j % p.MB;
      const double* src = red + size_t(a * TJ + b) * nact + (tb_i * p.MB + tb_j);
      double sum = 0.0;
      for (int gg = lane_id; gg < p.G; gg += 32)
        sum += src[size_t(gg) * NT];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
        sum += __shfl_down_sync(0xffffffffu, sum, off);
      if (lane_id == 0)
        out[e] = sum;
    }
  }
  gi_finalize(p.fin, km, &s_is_last);
}

//! out[e] = sum over CTAs (in CTA order within each lane, then a fixed shuffle tree) of partials[c][e]; one warp per element
__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* __restrict__ partials, int nparts, int km,
                                                              double* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= km)
    return;
  double sum = 0.0;
  for (int c = lane; c < nparts; c += 32)
    sum += partials[size_t(c) * km + warp];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
    sum += __shfl_down_sync(0xffffffffu, sum, off);
  if (lane == 0)
    out[warp] = sum;
}

//! decide where the per-CTA partial sums go and how they are finished (fused last-CTA reduction for small outputs)
void fill_finalize(itsolv_ctx* ctx, int grid_bound, int km, GiFinalize* f, bool* host_direct) {
  // grid_bound = the largest grid this kernel shape can have on this device: the decision must not depend on the
  // local vector length, all ranks of a communicator have to take the same path
  f->partials = ctx->d_partials;
  f->fused = (size_t(grid_bound) * km <= size_t(128) * 1024) ? 1 : 0;
  f->counter = ctx->d_counter;
  f->flag = nullptr;
  f->seq = 0;
  f->out = ctx->d_result;
  f->local = ctx->d_result;
  f->peers.nranks = 1;
  f->peers.rank = 0;
  f->peers.slot_doubles = 0;
  f->peers.error = nullptr;
  f->peers.timeout_ns = 0;
  f->dev_sums = nullptr;
  f->chain_out = nullptr;
  f->chain_offset = f->chain_count = 0;
  f->chain_thresh = 0.0;
  *host_direct = false;
  const int nranks = itsolv_comm_size(ctx);
  if (f->fused && (nranks == 1 || comm_peers(ctx, &f->peers))) {
    // mapped pinned memory: the kernel delivers the result (all-reduced over peer memory when there are several ranks),
    // no NCCL launch, no copy, no stream synchronisation
    f->out = ctx->h_result + ctx->result_offset;
    f->flag = ctx->h_flag;
    f->seq = ++ctx->flag_seq;
    *host_direct = true;
    if (ctx->chain_out) { // chained Gram-Schmidt step: the tail also prepares the next step's coefficients
      f->chain_out = ctx->chain_out;
      f->chain_offset = ctx->chain_offset;
      f->chain_count = ctx->chain_count;
      f->chain_thresh = ctx->chain_thresh;
      f->dev_sums = ctx->d_result + 8192; // the upper half of the device result buffer
    }
  }
  ctx->chain_out = nullptr; // requests hold for one launch
  ctx->result_offset = 0;
}

int launch_reduce_partials(itsolv_ctx* ctx, int grid, int km);
int gemm_inner_direct_device(itsolv_ctx* ctx, const double* const* xx, int k, const double* const* yy, int m, size_t n,
                             bool* host_direct, bool* handled);
int gemm_inner_mma_device(itsolv_ctx* ctx, const double* const* xx, int k, const double* const* yy, int m, size_t n,
                          bool* host_direct, bool* handled);

//! all-reduce over peer memory of sums that are already complete on this rank (f.local), delivered like the fused tail
__global__ void __launch_bounds__(256) peer_exchange_kernel(const __grid_constant__ GiFinalize f, int km) {
  const bool ok = gi_peer_allreduce(f.peers, f.local, km, f.seq, f.out);
  __syncthreads();
  if (threadIdx.x == 0 && f.flag) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(f.flag) = ok ? f.seq : (f.seq | (1ull << 63));
  }
}

//! after a non-fused reduction (or for an empty shard): exchange ctx->d_result with the peers when they are mapped
int finish_with_peers(itsolv_ctx* ctx, int km, bool* host_direct) {
  GiFinalize f{};
  if (itsolv_comm_size(ctx) == 1 || !comm_peers(ctx, &f.peers))
    return 0; // single rank or no peer mapping: finish_result() copies (after ncclAllReduce when needed)
  f.local = ctx->d_result;
  f.out = ctx->h_result;
  f.flag = ctx->h_flag;
  f.seq = ++ctx->flag_seq;
  peer_exchange_kernel<<<1, 256, 0, ctx->stream>>>(f, km);
  ITSOLV_CUDA(cudaGetLastError());
  ctx->counters.launches += 1;
  *host_direct = true;
  return 0;
}

using GiKernel = void (*)(const GiParams);

template <int TI, int TJ>
static GiKernel pick_variant(int loader, bool big) {
  constexpr bool heavy = TI * TJ >= 32;
  if (loader == LOAD_CPASYNC8)
    return gemm_inner_kernel<TI, TJ, LOAD_CPASYNC8, heavy>;
  if constexpr (heavy) {
    return loader == LOAD_TMA ? gemm_inner_kernel<TI, TJ, LOAD_TMA, true> : gemm_inner_kernel<TI, TJ, LOAD_CPASYNC16, true>;
  } else {
    if (loader == LOAD_TMA)
      return big ? gemm_inner_kernel<TI, TJ, LOAD_TMA, true> : gemm_inner_kernel<TI, TJ, LOAD_TMA, false>;
    return big ? gemm_inner_kernel<TI, TJ, LOAD_CPASYNC16, true> : gemm_inner_kernel<TI, TJ, LOAD_CPASYNC16, false>;
  }
}

static GiKernel pick_kernel(int ti, int tj, int loader, bool big) {
#define CASE(I, J)                                                                                                     \
  if (ti == I && tj == J)                                                                                              \
  return pick_variant<I, J>(loader, big)
  CASE(1, 1);
  CASE(1, 2);
  CASE(2, 1);
  CASE(2, 2);
  CASE(1, 4);
  CASE(4, 1);
  CASE(2, 4);
  CASE(4, 2);
  CASE(4, 4);
  CASE(4, 8);
  CASE(8, 4);
  CASE(8, 8);
#undef CASE
  return nullptr;
}

static int pow2_at_most(int v, int cap) {
  int r = 1;
  while (r * 2 <= v && r * 2 <= cap)
    r *= 2;
  return r;
}

/*!
 * Launch the contraction. Returns in *host_direct whether the result is delivered straight into ctx->h_result by the
 * kernel itself (single rank, fused final reduction: the host then waits on ctx->h_flag == ctx->flag_seq), otherwise the
 * result is left in ctx->d_result. No synchronisation here.
 */
int gemm_inner_device(itsolv_ctx* ctx, const double* const* xx, int k, const double* const* yy, int m, size_t n,
                      bool* host_direct) {
  ITSOLV_REQUIRE(k >= 1 && m >= 1 && k <= ITSOLV_MAX_PANEL && m <= ITSOLV_MAX_PANEL, "gemm_inner: panel size out of range");
  const int km = k * m;
  *host_direct = false;
  if (n == 0) { // an empty shard still takes part in the all-reduce
    ITSOLV_CUDA(cudaMemsetAsync(ctx->d_result, 0, size_t(km) * sizeof(double), ctx->stream));
    return finish_with_peers(ctx, km, host_direct);
  }
  {
    bool handled = false;
    if (gemm_inner_direct_device(ctx, xx, k, yy, m, n, host_direct, &handled))
      return 1;
    if (handled)
      return 0;
    if (gemm_inner_mma_device(ctx, xx, k, yy, m, n, host_direct, &handled))
      return 1;
    if (handled)
      return 0;
  }
  GiParams p;
  p.nvec = 0;
  bool async = true;
  auto slot_of = [&](const double* ptr) {
    for (int v = 0; v < p.nvec; ++v)
      if (p.vec[v] == ptr)
        return v;
    p.vec[p.nvec] = ptr;
    return p.nvec++;
  };
  for (int i = 0; i < k; ++i) {
    ITSOLV_REQUIRE(xx[i] != nullptr, "gemm_inner: null vector");
    p.xslot[i] = (unsigned char)slot_of(xx[i]);
    async = async && aligned16(xx[i]);
  }
  for (int j = 0; j < m; ++j) {
    ITSOLV_REQUIRE(yy[j] != nullptr, "gemm_inner: null vector");
    p.yslot[j] = (unsigned char)slot_of(yy[j]);
    async = async && aligned16(yy[j]);
  }
  for (int i = k; i < ITSOLV_MAX_PANEL; ++i)
    p.xslot[i] = 0;
  for (int j = m; j < ITSOLV_MAX_PANEL; ++j)
    p.yslot[j] = 0;
  p.k = k;
  p.m = m;
  p.n = n;

  // ---- thread tile: 4 x 4 accumulators per thread unless the panel is small or needs more than the CTA's threads
  int ti = pow2_at_most(k, 4), tj = pow2_at_most(m, 4);
  if (ctx->opt_gi_tile > 0) {
    ti = ctx->opt_gi_tile / 16;
    tj = ctx->opt_gi_tile % 16;
  }
  auto ntiles = [&](int a, int b) { return ((k + a - 1) / a) * ((m + b - 1) / b); };
  // ---- CTA shape: panels of many vectors run one large CTA per SM (bigger tiles -> bigger TMA copies, more producers)
  // ---- loader: TMA bulk copies pay ~200 issue cycles per copy and suit few, large copies; cp.async pieces suit the
  // rest and are the only choice for vectors that are not 16-byte aligned
  int loader = !async ? LOAD_CPASYNC8 : LOAD_TMA;
  if (async && ctx->opt_gi_loader == 2)
    loader = LOAD_CPASYNC16;
  // ---- CTA shape: panels of more than a few vectors run ONE CTA per SM with up to 8 producer warps (TMA copies
  // cost ~200 issue cycles each per issuing warp, so the issue rate, not the byte rate, has to be spread)
  bool big = p.nvec > 10;
  if (ctx->opt_gi_ctas == 1)
    big = true;
  else if (ctx->opt_gi_ctas >= 2)
    big = false;
  if (loader == LOAD_CPASYNC8)
    big = false; // instantiated for the two-CTA shape only (heavy tiles excepted, below)
  auto grow = [&](int cap) {
    while (ntiles(ti, tj) > cap) {
      if (tj <= ti && tj < 8)
        tj *= 2;
      else if (ti < 8)
        ti *= 2;
      else if (tj < 8)
        tj *= 2;
      else
        break;
    }
  };
  int max_consumers = 256;
  grow(max_consumers);
  const bool heavy = ti * tj >= 32;
  if (heavy) { // 32 or 64 accumulators per thread: 256 consumers, one CTA per SM
    big = true;
    max_consumers = 256;
    grow(max_consumers);
  }
  if (ctx->opt_gi_threads > 0)
    max_consumers = std::min(max_consumers, std::max(32, (ctx->opt_gi_threads / 32) * 32));
  ITSOLV_REQUIRE(ntiles(ti, tj) <= max_consumers, "gemm_inner: no thread tile fits this panel");
  p.KB = (k + ti - 1) / ti;
  p.MB = (m + tj - 1) / tj;
  const int NT = p.KB * p.MB;
  p.G = std::max(1, max_consumers / NT);
  const int nconsumers = ((NT * p.G + 31) / 32) * 32;
  const int ctas_per_sm = big ? 1 : 2;
  const int max_prod = big ? (heavy ? 4 : kMaxProducerWarps) : 2;
  p.nprod = max_prod;
  if (ctx->opt_gi_nprod > 0)
    p.nprod = std::min(ctx->opt_gi_nprod, max_prod);
  if (loader == LOAD_TMA)
    p.nprod = std::max(1, std::min(p.nprod, p.nvec)); // one vector per producer warp at least

  // ---- shared-memory tile: rows per stage and stages
  const size_t reduce_bytes = size_t(ti) * tj * NT * p.G * sizeof(double);
  const size_t smem_cap = size_t(ctx->max_smem_optin) - 2048;
  const size_t budget = (ctas_per_sm == 1 ? smem_cap : (smem_cap - 2048) / 2) & ~size_t(127);
  ITSOLV_REQUIRE(reduce_bytes <= budget, "gemm_inner: reduction scratch does not fit");
  int stages = ctx->opt_gi_stages > 0 ? ctx->opt_gi_stages : (p.nvec > 64 ? 2 : 3);
  stages = std::max(1, std::min(stages, kMaxStages));
  int rows;
  if (ctx->opt_gi_rows > 0) {
    rows = ctx->opt_gi_rows;
  } else {
    rows = int(std::min<size_t>(budget / (size_t(stages) * p.nvec * sizeof(double)), 2048));
    // keep enough tiles per CTA for a balanced static schedule
    const size_t want_tiles = size_t(ctx->num_sms) * ctas_per_sm * 16;
    while (rows > 64 && n / size_t(rows) < want_tiles)
      rows /= 2;
  }
  rows = std::max(16, (rows / 16) * 16);
  while (size_t(stages) * p.nvec * (rows + 2) * sizeof(double) > budget) {
    if (rows > 16)
      rows -= 16;
    else if (stages > 1)
      --stages;
    else
      break;
  }
  p.rows = rows;
  {
    const int chunk_bytes = ctx->opt_gi_chunk > 0 ? ctx->opt_gi_chunk : (1 << 20);
    p.chunk_rows = std::max(2, (chunk_bytes / 8 / 2) * 2);
  }
  p.stride = rows + 2;
  p.stages = stages;
  p.nfull = (long long)(n / size_t(rows));
  const size_t smem_bytes = std::max(size_t(stages) * p.nvec * p.stride * sizeof(double), reduce_bytes);
  ITSOLV_REQUIRE(smem_bytes <= smem_cap, "gemm_inner: shared-memory tile does not fit");

  const long long total_tiles = p.nfull + ((n % size_t(rows)) ? 1 : 0);
  const int grid = int(std::min<long long>(total_tiles, (long long)ctx->num_sms * ctas_per_sm));
  if (ensure_partials(ctx, size_t(grid) * km))
    return 1;
  fill_finalize(ctx, ctx->num_sms * ctas_per_sm, km, &p.fin, host_direct);

  GiKernel kernel = pick_kernel(ti, tj, loader, big);
  ITSOLV_REQUIRE(kernel != nullptr, "gemm_inner: thread tile not instantiated");
  if (ensure_dynamic_smem(ctx, reinterpret_cast<const void*>(kernel), smem_bytes))
    return 1;
  mark_launch(ctx);
  kernel<<<grid, nconsumers + 32 * p.nprod, smem_bytes, ctx->stream>>>(p);
  ITSOLV_CUDA(cudaGetLastError());
  ctx->counters.launches += 1;
  if (!p.fin.fused) {
    if (launch_reduce_partials(ctx, grid, km))
      return 1;
    return finish_with_peers(ctx, km, host_direct);
  }
  return 0;
}

int launch_reduce_partials(itsolv_ctx* ctx, int grid, int km) {
  const int rblocks = (km * 32 + 255) / 256;
  reduce_partials_kernel<<<rblocks, 256, 0, ctx->stream>>>(ctx->d_partials, grid, km, ctx->d_result);
  ITSOLV_CUDA(cudaGetLastError());
  ctx->counters.launches += 1;
  return 0;
}

//! wait for a result that the kernel writes into mapped host memory: spin on the sequence word, watch the stream for errors
static int wait_host_flag(itsolv_ctx* ctx) {
  volatile unsigned long long* flag = ctx->h_flag;
  const unsigned long long want = ctx->flag_seq;
  for (unsigned long long spins = 0;; ++spins) {
    const unsigned long long seen = *flag;
    if (seen == want)
      return 0;
    if (seen == (want | (1ull << 63))) {
      set_error("gemm_inner: a peer rank did not arrive at the all-reduce (timeout)");
      return 1;
    }
    if ((spins & 0x3FFF) == 0x3FFF) {
      const cudaError_t q = cudaStreamQuery(ctx->stream);
      if (q == cudaSuccess) {
        if (*flag == want)
          return 0;
        set_error("gemm_inner: kernel finished without delivering its result");
        return 1;
      }
      if (q != cudaErrorNotReady) {
        set_error(std::string("gemm_inner: ") + cudaGetErrorString(q));
        return 1;
      }
    }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
}

int wait_host_result(itsolv_ctx* ctx) { return wait_host_flag(ctx); }

//! all-reduce over ranks, copy to the pinned buffer, synchronise, hand to the caller
int finish_result(itsolv_ctx* ctx, int count, double* out, bool host_direct) {
  if (host_direct) {
    if (wait_host_flag(ctx))
      return 1;
  } else {
    if (comm_allreduce_device(ctx, ctx->d_result, size_t(count), false))
      return 1;
    ITSOLV_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->d_result, size_t(count) * sizeof(double), cudaMemcpyDeviceToHost,
                                ctx->stream));
    ITSOLV_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  for (int e = 0; e < count; ++e)
    out[e] = ctx->h_result[e];
  return 0;
}

static double distinct_bytes(const double* const* xx, int k, const double* const* yy, int m, size_t n) {
  int nd = 0;
  const double* seen[2 * ITSOLV_MAX_PANEL];
  auto add = [&](const double* ptr) {
    for (int v = 0; v < nd; ++v)
      if (seen[v] == ptr)
        return;
    seen[nd++] = ptr;
  };
  for (int i = 0; i < k; ++i)
    add(xx[i]);
  for (int j = 0; j < m; ++j)
    add(yy[j]);
  return 8.0 * double(n) * nd;
}

} // namespace itsolv

using namespace itsolv;

extern "C" {

int itsolv_gemm_inner_f64(itsolv_ctx* ctx, const double* const* xx, int k, const double* const* yy, int m, size_t n,
                          double* out) {
  if (k <= 0 || m <= 0)
    return 0;
  ctx->counters.n_gemm_inner++;
  // panels wider than ITSOLV_MAX_PANEL are processed block by block
  for (int i0 = 0; i0 < k; i0 += ITSOLV_MAX_PANEL) {
    const int kb = std::min(ITSOLV_MAX_PANEL, k - i0);
    for (int j0 = 0; j0 < m; j0 += ITSOLV_MAX_PANEL) {
      const int mb = std::min(ITSOLV_MAX_PANEL, m - j0);
      CallScope scope(ctx, OP_GEMM_INNER, distinct_bytes(xx + i0, kb, yy + j0, mb, n));
      bool direct = false;
      if (gemm_inner_device(ctx, xx + i0, kb, yy + j0, mb, n, &direct))
        return 1;
      scope.stop();
      if (kb == k && mb == m) {
        if (finish_result(ctx, kb * mb, out, direct))
          return 1;
      } else {
        std::vector<double> block(size_t(kb) * mb);
        if (finish_result(ctx, kb * mb, block.data(), direct))
          return 1;
        for (int i = 0; i < kb; ++i)
          for (int j = 0; j < mb; ++j)
            out[size_t(i0 + i) * m + (j0 + j)] = block[size_t(i) * mb + j];
      }
    }
  }
  return 0;
}

int itsolv_dot_f64(itsolv_ctx* ctx, const double* x, const double* y, size_t n, double* result) {
  ctx->counters.n_dot++;
  // a dot is the 1 x 1 case of the panel kernel and is accounted with it
  CallScope scope(ctx, OP_GEMM_INNER, (x == y ? 8.0 : 16.0) * double(n));
  bool direct = false;
  if (gemm_inner_device(ctx, &x, 1, &y, 1, n, &direct))
    return 1;
  scope.stop();
  return finish_result(ctx, 1, result, direct);
}

} // extern "C"
