// Shared tail of the gemm_inner kernels: the last CTA to publish its partial k x m sums adds all of them in CTA order
// (deterministic: the order does not depend on which CTA happens to be last) and hands the result to the host,
// either in device memory or straight into mapped pinned memory followed by a sequence word the host spins on.
#pragma once
#include <cuda_runtime.h>

namespace itsolv {

constexpr int kMaxPeers = 8; // GPUs of one NVSwitch box

/*!
 * One-shot all-reduce over NVLink peer memory, fused into the tail of the Gram kernels (replaces the reference's
 * MPI_Allreduce of the k x m block, array/util/gemm.h:179-182): every rank's finishing CTA stores its k x m sums into
 * slot [parity][my rank] of EVERY rank's exchange buffer (P2P stores through NVSwitch), publishes a sequence word
 * next to them, waits for the words of all ranks in its own buffer and adds the slots in rank order — so all ranks
 * obtain bitwise identical results (the host subspace problem is solved redundantly per rank) and no NCCL launch,
 * device-to-host copy or stream synchronisation is on the path. Two parities make the buffer reusable: a rank can be
 * at most one call ahead of any other.
 */
struct GiPeers {
  double* data[kMaxPeers];             // rank r's exchange buffer: [2][nranks][slot_doubles]
  unsigned long long* flags[kMaxPeers]; // rank r's sequence words: [2][nranks]
  int nranks, rank;
  int slot_doubles;
  // A rank that gave up waiting sets *error (device memory of THIS rank) and never clears it: every later launch
  // reports the failure instead of exchanging, so no result computed from a half-finished exchange is ever delivered.
  unsigned int* error;
  unsigned long long timeout_ns; // 0: wait for ever, as MPI_Allreduce does (ITSOLV_P2P_TIMEOUT_S, default 600 s)
};

struct GiFinalize {
  double* partials;         // [gridDim.x][km]
  double* out;              // km final sums (device or mapped host memory)
  double* local;            // device scratch for this rank's sums when peers are exchanged (else unused)
  unsigned int* counter;    // CTAs that have published (reset by the last one)
  unsigned long long* flag; // mapped host word that receives `seq` once `out` is complete (or null)
  unsigned long long seq;
  int fused;                // 0: a separate kernel reduces the partials
  GiPeers peers;            // peers.nranks <= 1: single rank
  // Chained Gram-Schmidt steps (mgs_fused.cu): the finishing CTA turns the row {<r,r>, <r,r_j>...} that starts at
  // sums[chain_offset] into the coefficients of the NEXT pivot step and leaves them in device memory, so that step can be
  // launched without a host round trip: chain_out = {1/|r|, null flag, -<r,r_j>/|r| ...}; null (|r| <= thresh): {1, 1, 0...}
  double* dev_sums;         // device copy of the km final sums (needed when chain_out is set)
  double* chain_out;        // null: no chaining
  int chain_offset, chain_count;
  double chain_thresh;
};

//! coefficients of the next pivot step from its row of inner products, exactly the host's arithmetic
//! (FusedDavidson.h, R-R modified Gram-Schmidt): norm = sqrt(|<r,r>|), o_j = <r,r_j> / norm
__device__ __forceinline__ void gi_chain(const GiFinalize& f) {
  const double* row = f.dev_sums + f.chain_offset;
  const double norm = sqrt(fabs(row[0]));
  const bool ok = norm > f.chain_thresh;
  f.chain_out[0] = ok ? __ddiv_rn(1.0, norm) : 1.0;
  f.chain_out[1] = ok ? 0.0 : 1.0;
  for (int t = 1; t < f.chain_count; ++t)
    f.chain_out[1 + t] = ok ? -__ddiv_rn(row[t], norm) : 0.0;
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ unsigned long long ld_volatile_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

/*!
 * Executed by ONE CTA per rank: `local` (km doubles, visible to this CTA) -> all peers, wait, sum in rank order -> out.
 * Returns false when a peer did not arrive within the time limit, or when an earlier call of this rank already failed
 * (sticky); the caller reports it through the host word. The limit only exists so that a dead peer turns into an error
 * instead of a hung GPU: skew between ranks (an unbalanced user callback, I/O) is waited for.
 */
__device__ __forceinline__ bool gi_peer_allreduce(const GiPeers& pr, const double* local, int km, unsigned long long seq,
                                                  double* out, double* dev_copy = nullptr) {
  const int tid = threadIdx.x;
  const int parity = int(seq & 1ull);
  const size_t slot = (size_t(parity) * pr.nranks + pr.rank) * pr.slot_doubles;
  if (*reinterpret_cast<volatile unsigned int*>(pr.error) != 0u)
    return false; // uniform over the CTA: the word is only written by this CTA's predecessors
  for (int r = 0; r < pr.nranks; ++r)
    for (int e = tid; e < km; e += blockDim.x)
      pr.data[r][slot + e] = local[e];
  __threadfence_system();
  __syncthreads();
  if (tid < pr.nranks)
    *reinterpret_cast<volatile unsigned long long*>(pr.flags[tid] + size_t(parity) * pr.nranks + pr.rank) = seq;
  __shared__ int s_ok;
  if (tid == 0)
    s_ok = 1;
  __syncthreads();
  if (tid < pr.nranks) {
    const unsigned long long* mine = pr.flags[pr.rank] + size_t(parity) * pr.nranks + tid;
    const unsigned long long t0 = global_timer_ns();
    unsigned int spins = 0;
    while (ld_volatile_sys(mine) != seq) {
      if (pr.timeout_ns != 0ull && (++spins & 0x3FFu) == 0u && global_timer_ns() - t0 > pr.timeout_ns) {
        s_ok = 0;
        break;
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (!s_ok) {
    if (tid == 0)
      *reinterpret_cast<volatile unsigned int*>(pr.error) = 1u;
    return false;
  }
  const double* base = pr.data[pr.rank] + size_t(parity) * pr.nranks * pr.slot_doubles;
  for (int e = tid; e < km; e += blockDim.x) {
    double sum = 0.0;
    for (int r = 0; r < pr.nranks; ++r)
      sum += __ldcv(base + size_t(r) * pr.slot_doubles + e);
    out[e] = sum;
    if (dev_copy)
      dev_copy[e] = sum;
  }
  return true;
}

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
  // Sum of the per-CTA partials in a fixed order (a function of the grid and of km only, not of which CTA is last):
  // `tpe` lanes share an entry, lane `sub` adds the CTAs c == sub (mod tpe) with four independent running sums - the
  // loads of a lane are in flight together, what used to be ~19 dependent L2 round trips (~10 us) for the usual
  // grid of 592 CTAs - and the lanes are folded by shuffles.
  const int nthreads = int(blockDim.x), grid = int(gridDim.x);
  int tpe = 1;
  while (tpe < 32 && tpe * 2 * km <= nthreads)
    tpe *= 2;
  const int per_pass = nthreads / tpe;
  const int sub = tid & (tpe - 1), slot = tid / tpe;
  for (int e0 = 0; e0 < km; e0 += per_pass) {
    const int e = e0 + slot;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (e < km) {
      const double* col = f.partials + e;
      int c = sub;
      for (; c + 3 * tpe < grid; c += 4 * tpe) {
        const double v0 = __ldcg(col + size_t(c) * km), v1 = __ldcg(col + size_t(c + tpe) * km);
        const double v2 = __ldcg(col + size_t(c + 2 * tpe) * km), v3 = __ldcg(col + size_t(c + 3 * tpe) * km);
        s0 += v0;
        s1 += v1;
        s2 += v2;
        s3 += v3;
      }
      for (; c < grid; c += tpe)
        s0 += __ldcg(col + size_t(c) * km);
    }
    double sum = (s0 + s1) + (s2 + s3);
    for (int off = tpe >> 1; off > 0; off >>= 1)
      sum += __shfl_down_sync(0xffffffffu, sum, off, tpe);
    if (sub == 0 && e < km) {
      (f.peers.nranks > 1 ? f.local : f.out)[e] = sum;
      if (f.chain_out && f.peers.nranks <= 1)
        f.dev_sums[e] = sum;
    }
  }
  __syncthreads();
  bool ok = true;
  if (f.peers.nranks > 1) {
    ok = gi_peer_allreduce(f.peers, f.local, km, f.seq, f.out, f.chain_out ? f.dev_sums : nullptr);
    __syncthreads();
  }
  if (tid == 0) {
    if (f.chain_out) {
      if (ok) {
        gi_chain(f);
      } else { // no sums: the next step of the chain must leave the vectors alone
        f.chain_out[0] = 1.0;
        f.chain_out[1] = 1.0;
        for (int t = 1; t < f.chain_count; ++t)
          f.chain_out[1 + t] = 0.0;
      }
    }
    *f.counter = 0u;
    if (f.flag) {
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long*>(f.flag) = ok ? f.seq : (f.seq | (1ull << 63));
    }
  }
}

} // namespace itsolv
