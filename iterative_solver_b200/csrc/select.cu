// select / select_max_dot: the n entries of a device vector that the reference's heap selection keeps
// (reference src/molpro/linalg/array/util/select.h:28-55, util/select_max_dot.h:22-46; distributed merge
// array/DistrArray.cpp:191-224). The reference keeps the n largest (key, index) PAIRS under lexicographic order,
// key = v or |v| (max) / -v or -|v| (min), so among equal keys the higher index wins. That rule is reproduced exactly
// by a most-significant-digit radix select over the 128-bit composite (order-preserving image of key, global index):
// histogram passes over the key bytes + the index bytes that can be non-zero, each finished by its last CTA (which picks
// the boundary bin), all on the device with no host round trip. After every key byte the last CTA looks at the size of
// the boundary bucket; as soon as it fits (2^20 elements) the next pass, while it reads the vector for its own digit,
// moves the bucket into a candidate buffer on which the remaining passes and the final compaction run. A vector whose
// values spread over many binary exponents is read twice (8n bytes each); a shard of a sorted diagonal, whose values
// share exponent and leading mantissa bits, four times. Runs once or twice per solve (initial guess, P-space choice).
#include <algorithm>
#include <cstring>
#include <utility>
#include <vector>

#include "common.cuh"

namespace itsolv {

// layout of ctx->d_select (unsigned long long words)
// SEL_GATHER_PASS: 1 + index of the pass that moves the boundary bucket into the candidate buffer (0 = none yet)
enum { SEL_KEY = 0, SEL_IDX = 1, SEL_REMAINING = 2, SEL_COUNT = 3, SEL_NCAND = 4, SEL_TICKET = 5, SEL_GATHER_PASS = 6, SEL_HIST = 16 };

constexpr unsigned long long kSelCandCapacity = 1ull << 20; // largest boundary bucket that is gathered
constexpr int kSelUnroll = 4;                              // trips of a pass whose loads are in flight together

struct SelParams {
  const double* x;
  const double* y; // nullptr unless select_max_dot
  size_t n;
  unsigned long long offset; // global index of x[0]
  int max, ignore_sign;
  // the gathered boundary bucket (key image, global index, value)
  const unsigned long long* cand_key;
  const unsigned long long* cand_idx;
  const double* cand_val;
};

__device__ __forceinline__ double sel_value(const SelParams& p, size_t i) {
  double v = p.x[i];
  if (p.y)
    return fabs(__dmul_rn(v, p.y[i]));
  return p.ignore_sign ? fabs(v) : v;
}
//! order-preserving map of the reference's heap key onto unsigned integers; -0.0 and +0.0 compare equal there, so both map to +0.0
__device__ __forceinline__ unsigned long long sel_key(const SelParams& p, double value) {
  double key = (p.max || p.y) ? value : -value;
  key = key + 0.0;
  const long long b = __double_as_longlong(key);
  return static_cast<unsigned long long>(b) ^ (static_cast<unsigned long long>(b >> 63) | 0x8000000000000000ull);
}

/*!
 * What a pass asks of every element, prepared once per thread: does it agree with the digits fixed so far, and what is
 * its digit `d` (0..7 key bytes, 8..15 index bytes). Masks and a 32-bit word select instead of 64-bit shifts by a
 * run-time amount per element (the pass was bound by instruction issue at ~220 instructions per element).
 */
struct SelDigit {
  unsigned long long kmask, kval; // key bytes fixed so far (all of them for an index digit)
  unsigned long long imask, ival; // index bytes fixed so far (none for a key digit)
  bool on_index, high_word;
  int shift; // of the digit within its 32-bit word
  __device__ __forceinline__ SelDigit(int d, unsigned long long kpre, unsigned long long ipre) {
    const int dd = d < 8 ? d : d - 8;
    on_index = d >= 8;
    const unsigned long long fixed = dd == 0 ? 0ull : ~0ull << (64 - 8 * dd);
    kmask = on_index ? ~0ull : fixed;
    kval = kpre & kmask;
    imask = on_index ? fixed : 0ull;
    ival = ipre & imask;
    high_word = dd < 4;
    shift = 24 - 8 * (dd & 3);
  }
  __device__ __forceinline__ bool match(unsigned long long u, unsigned long long gi, unsigned& digit) const {
    const unsigned long long src = on_index ? gi : u;
    const unsigned word = high_word ? unsigned(src >> 32) : unsigned(src);
    digit = (word >> shift) & 0xFFu;
    return (u & kmask) == kval && (gi & imask) == ival;
  }
  //! above the fixed key prefix: selected for certain (asked by the gathering pass)
  __device__ __forceinline__ bool above(unsigned long long u) const { return (u & kmask) > kval; }
};

/*!
 * One pass of the selection: histogram of digit `d` over the elements that agree with the digits fixed so far; the last
 * CTA to finish then picks the bin that holds the boundary element (walk from the top bin until `remaining` elements are
 * covered), fixes the digit and clears the histogram for the next pass.
 *
 * The pass reads the vector until the boundary bucket has been gathered, then the candidates. The last CTA of a key-byte
 * pass that finds the bucket small enough (kSelCandCapacity) names the next pass (`next_d`) as the gathering pass
 * (SEL_GATHER_PASS): that pass, still reading the vector, also sends the elements above the boundary prefix straight to
 * the output (they are selected for certain) and the elements with the boundary prefix to the candidate buffer.
 * A warp whose elements all fall into one bin (every warp, on a shard of a sorted diagonal) spends one atomic on them;
 * the appends to the output and to the candidates cost one atomic per warp.
 */
__global__ void __launch_bounds__(256)
    select_pass_kernel(const __grid_constant__ SelParams p, int d, int next_d, unsigned long long* __restrict__ state,
                       unsigned long long* __restrict__ cand_key, unsigned long long* __restrict__ cand_idx,
                       double* __restrict__ cand_val, long long* __restrict__ out_idx, double* __restrict__ out_val,
                       unsigned long long capacity) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long suffix[257];
  __shared__ int s_last;
  const unsigned long long kpre = state[SEL_KEY], ipre = state[SEL_IDX];
  const unsigned long long gpass = state[SEL_GATHER_PASS]; // 1 + index of the gathering pass, 0: none named yet
  const bool gather_now = gpass == (unsigned long long)(d + 1);
  const bool from_candidates = gpass != 0 && !gather_now;
  const size_t count = from_candidates ? size_t(state[SEL_NCAND]) : p.n;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned full = 0xFFFFFFFFu, below = (1u << lane) - 1u;
  // the grid is sized for the vector; on the candidates only the CTAs that have elements take part (and CTA 0 always)
  const size_t blocks = (count + blockDim.x - 1) / blockDim.x;
  const unsigned participating = unsigned(blocks < 1 ? 1 : (blocks < gridDim.x ? blocks : gridDim.x));
  if (blockIdx.x >= participating)
    return;
  if (size_t(blockIdx.x) * blockDim.x < count) {
    hist[threadIdx.x] = 0;
    __syncthreads();
    const SelDigit spec(d, kpre, ipre);
    // Warp-uniform trip count: every lane of a warp takes part in the votes below. The loads of kSelUnroll trips are
    // issued before any vote: the votes are convergence points the compiler does not move loads across, and with one
    // trip in flight the pass is bound by DRAM latency (65 us for 80 MB) instead of bandwidth.
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t base0 = size_t(blockIdx.x) * blockDim.x + (threadIdx.x - lane); base0 < count;
         base0 += kSelUnroll * stride) {
      double value[kSelUnroll];
      unsigned long long key[kSelUnroll], gidx[kSelUnroll];
      bool valid[kSelUnroll];
#pragma unroll
      for (int t = 0; t < kSelUnroll; ++t) {
        const size_t i = base0 + size_t(t) * stride + lane;
        valid[t] = i < count;
        value[t] = 0;
        key[t] = 0;
        gidx[t] = 0;
        if (valid[t]) {
          if (from_candidates) {
            key[t] = p.cand_key[i];
            gidx[t] = p.cand_idx[i];
          } else {
            value[t] = sel_value(p, i);
            gidx[t] = p.offset + i;
          }
        }
      }
#pragma unroll
      for (int t = 0; t < kSelUnroll; ++t) {
        if (base0 + size_t(t) * stride >= count) // warp-uniform
          break;
        unsigned digit = 0;
        bool ok = false, certain = false;
        const unsigned long long u = from_candidates ? key[t] : sel_key(p, value[t]);
        const unsigned long long gi = gidx[t];
        if (valid[t]) {
          ok = spec.match(u, gi, digit);
          certain = gather_now && spec.above(u);
        }
        const unsigned okmask = __ballot_sync(full, ok);
        if (gather_now) {
          const unsigned cmask = __ballot_sync(full, certain);
          if (cmask) {
            const int leader = __ffs(int(cmask)) - 1;
            unsigned long long pos = 0;
            if (int(lane) == leader)
              pos = atomicAdd(&state[SEL_COUNT], (unsigned long long)__popc(cmask));
            pos = __shfl_sync(full, pos, leader) + __popc(cmask & below);
            if (certain && pos < capacity) {
              out_idx[pos] = (long long)gi;
              out_val[pos] = value[t];
            }
          }
          if (okmask) {
            const int leader = __ffs(int(okmask)) - 1;
            unsigned long long pos = 0;
            if (int(lane) == leader)
              pos = atomicAdd(&state[SEL_NCAND], (unsigned long long)__popc(okmask));
            pos = __shfl_sync(full, pos, leader) + __popc(okmask & below);
            if (ok && pos < kSelCandCapacity) {
              cand_key[pos] = u;
              cand_idx[pos] = gi;
              cand_val[pos] = value[t];
            }
          }
        }
        if (okmask) {
          const int leader = __ffs(int(okmask)) - 1;
          const unsigned first = __shfl_sync(full, digit, leader);
          if (__all_sync(full, !ok || digit == first)) {
            if (int(lane) == leader)
              atomicAdd(&hist[first], unsigned(__popc(okmask)));
          } else if (ok) {
            atomicAdd(&hist[digit], 1u);
          }
        }
      }
    }
    __syncthreads();
    if (hist[threadIdx.x])
      atomicAdd(&state[SEL_HIST + threadIdx.x], (unsigned long long)hist[threadIdx.x]);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0)
    s_last = atomicAdd(&state[SEL_TICKET], 1ull) == participating - 1 ? 1 : 0;
  __syncthreads();
  if (!s_last)
    return;
  __threadfence();
  // inclusive suffix sums: suffix[b] = number of elements in bins >= b
  const int t = threadIdx.x;
  const unsigned long long remaining = state[SEL_REMAINING]; // read by all before the barriers below, written after them
  suffix[t] = __ldcg(&state[SEL_HIST + t]);
  if (t == 0)
    suffix[256] = 0;
  __syncthreads();
  for (int step = 1; step < 256; step <<= 1) {
    const unsigned long long add = t + step < 256 ? suffix[t + step] : 0;
    __syncthreads();
    suffix[t] += add;
    __syncthreads();
  }
  if (suffix[t] >= remaining && suffix[t + 1] < remaining) {
    state[SEL_REMAINING] = remaining - suffix[t + 1];
    if (d < 8 && next_d >= 0 && gpass == 0 && suffix[t] - suffix[t + 1] <= kSelCandCapacity)
      state[SEL_GATHER_PASS] = (unsigned long long)(next_d + 1);
    if (d < 8)
      state[SEL_KEY] = kpre | ((unsigned long long)t << (56 - 8 * d));
    else
      state[SEL_IDX] = ipre | ((unsigned long long)t << (56 - 8 * (d - 8)));
  }
  state[SEL_HIST + t] = 0;
  if (t == 0)
    state[SEL_TICKET] = 0;
}

//! the elements (of the candidates, if they were gathered; elements above their prefix are in the output already) that
//! lie at or above the boundary (key, index) pair
__global__ void __launch_bounds__(256)
    select_compact_kernel(const __grid_constant__ SelParams p, unsigned long long* __restrict__ state,
                          long long* __restrict__ out_idx, double* __restrict__ out_val, unsigned long long capacity) {
  const unsigned long long kthr = state[SEL_KEY], ithr = state[SEL_IDX];
  const bool from_candidates = state[SEL_GATHER_PASS] != 0;
  const size_t count = from_candidates ? size_t(state[SEL_NCAND]) : p.n;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += size_t(gridDim.x) * blockDim.x) {
    const double value = from_candidates ? p.cand_val[i] : sel_value(p, i);
    const unsigned long long u = from_candidates ? p.cand_key[i] : sel_key(p, value);
    const unsigned long long gi = from_candidates ? p.cand_idx[i] : p.offset + i;
    if (u > kthr || (u == kthr && gi >= ithr)) {
      const unsigned long long pos = atomicAdd(&state[SEL_COUNT], 1ull);
      if (pos < capacity) {
        out_idx[pos] = (long long)gi;
        out_val[pos] = value;
      }
    }
  }
}

static double host_key(double value, bool max) {
  double key = max ? value : -value;
  return key + 0.0;
}

} // namespace itsolv

using namespace itsolv;

extern "C" {

int itsolv_select_merge(const int64_t* idx, const double* val, size_t ncand, size_t nsel, int max, int /*ignore_sign*/,
                        int64_t* out_idx, double* out_val) {
  // values arrive already as |v| when ignore_sign was requested; order by the reference's (key, index) pairs
  std::vector<std::pair<std::pair<double, int64_t>, double>> c;
  c.reserve(ncand);
  for (size_t i = 0; i < ncand; ++i)
    if (idx[i] >= 0)
      c.push_back({{host_key(val[i], max != 0), idx[i]}, val[i]});
  std::sort(c.begin(), c.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
  if (c.size() > nsel)
    c.resize(nsel);
  std::sort(c.begin(), c.end(), [](const auto& a, const auto& b) { return a.first.second < b.first.second; });
  for (size_t i = 0; i < c.size(); ++i) {
    out_idx[i] = c[i].first.second;
    out_val[i] = c[i].second;
  }
  return int(c.size());
}

int itsolv_select_f64(itsolv_ctx* ctx, const double* x, const double* y, size_t n, size_t global_offset, size_t nsel,
                      int max, int ignore_sign, int64_t* indices, double* values, int* nfound) {
  ctx->counters.n_select++;
  const int nranks = itsolv_comm_size(ctx);
  ITSOLV_REQUIRE(nsel * 16 * size_t(nranks) <= ctx->stage_slot_bytes, "itsolv_select_f64: too many entries requested");
  *nfound = 0;
  if (nsel == 0)
    return 0;
  const size_t nloc = std::min(nsel, n);
  CallScope scope(ctx, OP_OTHER, 8.0 * double(n) * (y ? 2 : 1));
  // candidate buffer on the device: [idx(nsel) | val(nsel)] per rank, gathered rank after rank
  char *h = nullptr, *d = nullptr;
  int slot = 0;
  if (stage_acquire(ctx, nsel * 16 * size_t(nranks), &h, &d, &slot))
    return 1;
  const int myrank = itsolv_comm_rank(ctx);
  char* dmine = d + size_t(myrank) * nsel * 16;
  long long* d_idx = reinterpret_cast<long long*>(dmine);
  double* d_val = reinterpret_cast<double*>(dmine + nsel * 8);
  ITSOLV_CUDA(cudaMemsetAsync(dmine, 0xFF, nsel * 8, ctx->stream)); // idx = -1: empty candidate
  if (nloc > 0) {
    if (!ctx->d_select_cand)
      ITSOLV_CUDA(cudaMalloc(&ctx->d_select_cand, kSelCandCapacity * 3 * sizeof(unsigned long long)));
    unsigned long long* cand_key = ctx->d_select_cand;
    unsigned long long* cand_idx = cand_key + kSelCandCapacity;
    double* cand_val = reinterpret_cast<double*>(cand_idx + kSelCandCapacity);
    SelParams p{x, y, n, (unsigned long long)global_offset, (max || y) ? 1 : 0, ignore_sign, cand_key, cand_idx, cand_val};
    unsigned long long init[SEL_HIST + 256];
    std::memset(init, 0, sizeof(init));
    init[SEL_REMAINING] = nloc;
    ITSOLV_CUDA(cudaMemcpyAsync(ctx->d_select, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    const int grid = int(std::min<size_t>((n + 255) / 256, size_t(ctx->num_sms) * 8));
    // index bytes above the largest global index are zero for every element: skip those passes
    int idx_bytes = 1;
    while (idx_bytes < 8 && ((global_offset + n - 1) >> (8 * idx_bytes)) != 0)
      ++idx_bytes;
    mark_launch(ctx);
    int passes[16], npass = 0;
    for (int dgt = 0; dgt < 16; ++dgt)
      if (dgt < 8 || dgt - 8 >= 8 - idx_bytes)
        passes[npass++] = dgt;
    // whether a pass runs over the vector or over the candidates is decided on the device: the grid is always the full
    // one, CTAs beyond the candidates return at once
    for (int k = 0; k < npass; ++k) {
      select_pass_kernel<<<grid, 256, 0, ctx->stream>>>(p, passes[k], k + 1 < npass ? passes[k + 1] : -1, ctx->d_select,
                                                        cand_key, cand_idx, cand_val, d_idx, d_val, nloc);
      ctx->counters.launches += 1;
    }
    select_compact_kernel<<<grid, 256, 0, ctx->stream>>>(p, ctx->d_select, d_idx, d_val, nloc);
    ctx->counters.launches += 1;
    ITSOLV_CUDA(cudaGetLastError());
  }
  if (comm_allgather_device(ctx, dmine, d, nsel * 16))
    return 1;
  ITSOLV_CUDA(cudaMemcpyAsync(h, d, nsel * 16 * size_t(nranks), cudaMemcpyDeviceToHost, ctx->stream));
  ITSOLV_CUDA(cudaStreamSynchronize(ctx->stream));
  if (stage_done(ctx, slot))
    return 1;
  std::vector<int64_t> cidx(nsel * nranks);
  std::vector<double> cval(nsel * nranks);
  for (int r = 0; r < nranks; ++r) {
    std::memcpy(cidx.data() + size_t(r) * nsel, h + size_t(r) * nsel * 16, nsel * 8);
    std::memcpy(cval.data() + size_t(r) * nsel, h + size_t(r) * nsel * 16 + nsel * 8, nsel * 8);
  }
  *nfound = itsolv_select_merge(cidx.data(), cval.data(), cidx.size(), nsel, (max || y) ? 1 : 0, ignore_sign, indices, values);
  return 0;
}

} // extern "C"
