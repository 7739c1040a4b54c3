// Shared internals of libitsolv_b200: context, error handling, accounting, small device helpers.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include <itsolv_b200.h>

namespace itsolv {

void set_error(const std::string& msg);

#define ITSOLV_CUDA(call)                                                                                              \
  do {                                                                                                                 \
    cudaError_t err__ = (call);                                                                                        \
    if (err__ != cudaSuccess) {                                                                                        \
      ::itsolv::set_error(std::string(#call) + ": " + cudaGetErrorString(err__) + " (" + __FILE__ + ":" +             \
                          std::to_string(__LINE__) + ")");                                                             \
      return 1;                                                                                                        \
    }                                                                                                                  \
  } while (0)

#define ITSOLV_REQUIRE(cond, msg)                                                                                      \
  do {                                                                                                                 \
    if (!(cond)) {                                                                                                     \
      ::itsolv::set_error(std::string(msg));                                                                           \
      return 2;                                                                                                        \
    }                                                                                                                  \
  } while (0)

struct Comm; // NCCL communicator wrapper (comm.cu)

enum OpClass { OP_BLAS1 = 0, OP_GEMM_INNER = 1, OP_GEMM_OUTER = 2, OP_OTHER = 3, OP_RESIDUAL = 4 };

struct CallScope;

struct PendingEvent {
  cudaEvent_t start, stop;
  int cls;
};

} // namespace itsolv

struct itsolv_ctx {
  int device = 0;
  int num_sms = 148;
  int max_smem_optin = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaMemPool_t pool = nullptr;
  std::unordered_map<const void*, size_t> alloc_bytes; // live vectors handed out by itsolv_alloc
  size_t live_bytes = 0, peak_bytes = 0;

  // workspace for per-CTA partial Gram matrices and reduced results (device), pinned host mirror for results
  double* d_partials = nullptr;
  size_t partials_capacity = 0; // doubles
  double* d_result = nullptr;   // ITSOLV_MAX_PANEL^2 doubles
  double* h_result = nullptr;   // pinned + mapped, ITSOLV_MAX_PANEL^2 doubles
  unsigned long long* h_flag = nullptr; // pinned + mapped: sequence number of the last result delivered by a kernel
  unsigned long long flag_seq = 0;
  unsigned int* d_counter = nullptr;    // CTA arrival counter of the fused final reduction (self-resetting)
  // requests picked up by the next fill_finalize() (chained Gram-Schmidt steps, mgs_fused.cu)
  size_t result_offset = 0;             // the kernel delivers its sums at h_result + result_offset
  double* chain_out = nullptr;          // see GiFinalize::chain_out
  int chain_offset = 0, chain_count = 0;
  double chain_thresh = 0.0;
  // staging for small host->device payloads (alphas, sparse maps, pointer tables): pinned ring + device ring
  char* h_stage = nullptr;
  char* d_stage = nullptr;
  size_t stage_slot_bytes = 0;
  int stage_slots = 0;
  int stage_next = 0;
  std::vector<cudaEvent_t> stage_events;
  // select scratch
  unsigned long long* d_select = nullptr;
  unsigned long long* d_select_cand = nullptr; // candidates of the radix select: [key | global index | value] x capacity

  itsolv::Comm* comm = nullptr;

  // options
  int opt_gi_rows = 0;    // rows per smem tile of gemm_inner (0 = auto)
  int opt_gi_stages = 0;  // pipeline stages (0 = auto)
  int opt_gi_threads = 0; // max threads per CTA (0 = auto)
  int opt_gi_tile = 0;    // thread tile TI*16+TJ (0 = auto)
  int opt_gi_ctas = 0;    // CTAs per SM (0 = auto)
  int opt_gi_chunk = 0;   // bytes per TMA copy (0 = auto)
  int opt_gi_nprod = 0;   // producer warps (0 = auto)
  int opt_gi_loader = 0;  // 0 auto, 1 TMA bulk copies, 2 cp.async pieces
  int opt_gi_direct = 0;  // 0 auto (register-streaming kernel for k*m <= 4), 2 never
  int opt_gi_direct_ctas = 0; // CTAs per SM of the register-streaming kernel
  int opt_gi_mma = 0;     // FP64 tensor-core path: 0 auto (k*m >= 320), >0 minimum k*m, <0 never
  int opt_go_cols = 0;    // gemm_outer columns per thread (0 = auto)
  int opt_go_ctas = 0;    // gemm_outer CTAs per SM
  int opt_blas1_ctas = 0; // CTAs per SM for streaming kernels
  int opt_ds_ring = 0;    // davidson_residual: <0 never use the cp.async ring kernel
  int opt_mgs_chain = 0;   // R-R Gram-Schmidt steps chained on the device: 0 default (on, see mgs_fused.cu), -1 off
  int opt_project_chain = 0; // <0: the projection of a working set does not carry the first Gram row of the chain
  int opt_p2p_allreduce = 0; // <0: use ncclAllReduce even when the peer buffers are mapped
  int opt_p2p_halo = 0;      // <0: halo rows by ncclSend/ncclRecv even when the peer buffers are mapped

  std::vector<std::pair<const void*, size_t>> smem_optin; // kernel -> largest dynamic shared memory size opted in

  itsolv::CallScope* active_scope = nullptr;
  unsigned long long write_epoch = 1; // advanced by every launch (see mark_launch): host-side result caches key on it
  itsolv_counters counters{};
  bool profiling = false;
  std::vector<itsolv::PendingEvent> pending;
  std::vector<cudaEvent_t> event_pool;
  cudaEvent_t timer_start[ITSOLV_TIMERS] = {}, timer_stop[ITSOLV_TIMERS] = {};
};

namespace itsolv {

//! RAII accounting of one C-ABI call: algorithmic bytes, call class and (when profiling) a CUDA-event pair.
struct CallScope {
  itsolv_ctx* ctx;
  int cls;
  cudaEvent_t start = nullptr;
  bool armed = false; // profiling on: the opening event is recorded by mark_launch() right before the first kernel launch
  CallScope(itsolv_ctx* c, int cls, double bytes);
  //! record the closing event now (after the last kernel launch, before any host synchronisation)
  void stop();
  ~CallScope();
};
void drain_pending(itsolv_ctx* ctx);
//! call immediately before a kernel launch: opens the event pair of the active CallScope (host preparation excluded)
void mark_launch(itsolv_ctx* ctx);

//! Reserve a staging slot (pinned host + matching device region) of at least `bytes`; caller fills host, then commit copies.
int stage_acquire(itsolv_ctx* ctx, size_t bytes, char** host, char** dev, int* slot);
int stage_commit(itsolv_ctx* ctx, int slot, size_t bytes);
//! call after launching the kernel that reads the device side of the slot
int stage_done(itsolv_ctx* ctx, int slot);

int ensure_partials(itsolv_ctx* ctx, size_t doubles);
//! opt a kernel in to `bytes` of dynamic shared memory (cached per kernel: the attribute call is made only when it grows)
int ensure_dynamic_smem(itsolv_ctx* ctx, const void* kernel, size_t bytes);

//! sum in place over ranks (device buffer); no-op without a communicator
int comm_allreduce_device(itsolv_ctx* ctx, double* d, size_t count, bool op_max);
int comm_allgather_device(itsolv_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank);
struct GiPeers;
//! fills the peer table when the exchange buffers of all ranks are mapped (itsolv_comm_p2p_import); false otherwise
bool comm_peers(itsolv_ctx* ctx, GiPeers* peers);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#ifdef __CUDACC__
/*!
 * IEEE round-to-nearest quotient, bit for bit what `/` gives on the host. __ddiv_rn's inline sequence (reciprocal
 * estimate + Newton steps + remainder correction, ~15 instructions) leaves every numerator below 2^-120 - zero
 * included - to a generic routine of ~60 instructions. Residual vectors are full of exact zeros wherever the operator has
 * not coupled a row to the starting vectors yet, so 0 / x (x finite or infinite, not zero, not NaN) is answered here:
 * a zero with the sign of the quotient.
 */
__device__ __forceinline__ double div_rn(double num, double den) {
  if (num == 0.0 && den != 0.0 && den == den)
    return __hiloint2double((__double2hiint(num) ^ __double2hiint(den)) & 0x80000000, 0);
  return __ddiv_rn(num, den);
}
#endif

} // namespace itsolv
