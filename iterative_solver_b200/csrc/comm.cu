// Communicator: NCCL over NVLink/NVSwitch, one rank per GPU. Replaces the reference's MPI calls on the vector path
// (MPI_Allreduce of k*m doubles, reference array/util/gemm.h:179-182; of one double, array/DistrArray.cpp:134-136;
// the gather/broadcast of select candidates, array/DistrArray.cpp:191-224).
//
// libnccl is resolved at run time with dlopen so that the library loads on a box without NCCL and, inside a torch
// process, binds to the NCCL that torch already loaded. There is no host fallback: without NCCL comm_init fails.
#include <dlfcn.h>

#include <cstring>

#include "common.cuh"
#include "gi_finalize.cuh"

#define ITSOLV_MAX_ROOTS_HALO 64

namespace itsolv {

// minimal NCCL surface (ABI-stable since NCCL 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclChar = 0, ncclUint8 = 1, ncclFloat64 = 8, ncclDouble = 8 };
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried)
    return api.handle ? &api : nullptr;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle)
      break;
  }
  if (!api.handle) {
    if (const char* extra = std::getenv("ITSOLV_NCCL_LIBRARY"))
      api.handle = dlopen(extra, RTLD_NOW | RTLD_GLOBAL);
  }
  if (!api.handle)
    return nullptr;
#define LOAD(sym) api.sym = reinterpret_cast<decltype(api.sym)>(dlsym(api.handle, "nccl" #sym))
  LOAD(GetUniqueId);
  LOAD(CommInitRank);
  LOAD(CommDestroy);
  LOAD(AllReduce);
  LOAD(AllGather);
  LOAD(Send);
  LOAD(Recv);
  LOAD(GroupStart);
  LOAD(GroupEnd);
  LOAD(GetErrorString);
#undef LOAD
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.AllGather || !api.Send || !api.Recv ||
      !api.GroupStart || !api.GroupEnd) {
    api.handle = nullptr;
    return nullptr;
  }
  return &api;
}

constexpr size_t kPeerSlotDoubles = size_t(ITSOLV_MAX_PANEL) * ITSOLV_MAX_PANEL;
constexpr size_t kHaloSlotDoubles = 4096; // boundary rows of a whole working set: w * b <= 4096

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, size = 1;
  double* d_small = nullptr; // staging for host all-reduces and halos
  // peer exchange buffers of the fused all-reduce (gi_finalize.cuh): mine + the IPC-mapped ones of the other ranks
  char* d_exchange = nullptr; // [2][size][kPeerSlotDoubles] doubles, then [2][size] sequence words
  void* peer_base[kMaxPeers] = {};
  bool peers_ready = false;
  unsigned int* d_error = nullptr;        // sticky failure word of the fused all-reduce (GiPeers::error)
  unsigned long long timeout_ns = 600ull * 1000000000ull;
  // layout of an exchange buffer: all-reduce slots, all-reduce sequence words, then the halo area of the harness operator
  // ([parity][side: 0 = rows from the rank below, 1 = from the rank above][kHaloSlotDoubles]) and its sequence words
  size_t allreduce_bytes() const { return 2 * size_t(size) * kPeerSlotDoubles * sizeof(double) + 2 * size_t(size) * 8 + 64; }
  size_t halo_data_offset() const { return (allreduce_bytes() + 255) & ~size_t(255); }
  size_t halo_flag_offset() const { return halo_data_offset() + 4 * kHaloSlotDoubles * sizeof(double); }
  size_t exchange_bytes() const { return halo_flag_offset() + 4 * 8 + 64; }
  unsigned long long halo_seq = 0;
};

#define ITSOLV_NCCL(call)                                                                                              \
  do {                                                                                                                 \
    int r__ = (call);                                                                                                  \
    if (r__ != ncclSuccess) {                                                                                          \
      NcclApi* a__ = nccl_api();                                                                                       \
      set_error(std::string(#call) + ": " + (a__ && a__->GetErrorString ? a__->GetErrorString(r__) : "nccl error"));  \
      return 1;                                                                                                        \
    }                                                                                                                  \
  } while (0)

bool comm_peers(itsolv_ctx* ctx, GiPeers* peers) {
  Comm* c = ctx->comm;
  if (!c || c->size <= 1 || !c->peers_ready || ctx->opt_p2p_allreduce < 0)
    return false;
  peers->nranks = c->size;
  peers->rank = c->rank;
  peers->slot_doubles = int(kPeerSlotDoubles);
  peers->error = c->d_error;
  peers->timeout_ns = c->timeout_ns;
  const size_t data_bytes = 2 * size_t(c->size) * kPeerSlotDoubles * sizeof(double);
  for (int r = 0; r < c->size; ++r) {
    peers->data[r] = reinterpret_cast<double*>(c->peer_base[r]);
    peers->flags[r] = reinterpret_cast<unsigned long long*>(static_cast<char*>(c->peer_base[r]) + data_bytes);
  }
  return true;
}

struct HaloParams {
  const double* x[ITSOLV_MAX_ROOTS_HALO];
  double* out;              // [w][2 b]: per vector the b rows from below, then the b rows from above
  double* below_data;       // halo area of rank - 1 (null at the lower end): my first rows go to its side 1
  double* above_data;       // halo area of rank + 1 (null at the upper end): my last rows go to its side 0
  unsigned long long* below_flags;
  unsigned long long* above_flags;
  double* my_data;
  unsigned long long* my_flags;
  unsigned int* error;
  unsigned long long timeout_ns, seq;
  size_t nloc;
  int w, b;
};

/*!
 * Boundary rows of all w vectors of a working set to the two neighbouring shards in one launch: stores into the
 * neighbours' halo areas over NVLink peer memory, a sequence word per side, then the rows the neighbours stored here are
 * copied to `out`. One CTA; replaces w grouped ncclSend/ncclRecv pairs per operator application.
 */
__global__ void __launch_bounds__(256) halo_exchange_kernel(const __grid_constant__ HaloParams p) {
  const int tid = threadIdx.x;
  const int parity = int(p.seq & 1ull);
  const int cnt = p.w * p.b;
  __shared__ int s_ok;
  if (tid == 0)
    s_ok = *reinterpret_cast<volatile unsigned int*>(p.error) == 0u ? 1 : 0;
  __syncthreads();
  if (s_ok) {
    for (int e = tid; e < cnt; e += blockDim.x) {
      const int k = e / p.b, j = e % p.b;
      if (p.below_data)
        p.below_data[(size_t(parity) * 2 + 1) * kHaloSlotDoubles + e] = p.x[k][j];
      if (p.above_data)
        p.above_data[(size_t(parity) * 2 + 0) * kHaloSlotDoubles + e] = p.x[k][p.nloc - size_t(p.b) + j];
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0 && p.below_flags)
      *reinterpret_cast<volatile unsigned long long*>(p.below_flags + parity * 2 + 1) = p.seq;
    if (tid == 1 && p.above_flags)
      *reinterpret_cast<volatile unsigned long long*>(p.above_flags + parity * 2 + 0) = p.seq;
    // wait for the rows of the neighbours: side 0 is written by the rank below, side 1 by the rank above
    if (tid < 2 && (tid == 0 ? p.below_data != nullptr : p.above_data != nullptr)) {
      const unsigned long long* mine = p.my_flags + parity * 2 + tid;
      const unsigned long long t0 = global_timer_ns();
      unsigned int spins = 0;
      while (ld_volatile_sys(mine) != p.seq) {
        if (p.timeout_ns != 0ull && (++spins & 0x3FFu) == 0u && global_timer_ns() - t0 > p.timeout_ns) {
          s_ok = 0;
          *reinterpret_cast<volatile unsigned int*>(p.error) = 1u;
          break;
        }
      }
    }
    __threadfence_system();
    __syncthreads();
  }
  for (int e = tid; e < cnt; e += blockDim.x) {
    const int k = e / p.b, j = e % p.b;
    double* o = p.out + size_t(k) * 2 * p.b;
    o[j] = (s_ok && p.below_data) ? __ldcv(p.my_data + (size_t(parity) * 2 + 0) * kHaloSlotDoubles + e) : 0.0;
    o[p.b + j] = (s_ok && p.above_data) ? __ldcv(p.my_data + (size_t(parity) * 2 + 1) * kHaloSlotDoubles + e) : 0.0;
  }
}

void comm_destroy(itsolv_ctx* ctx) {
  if (!ctx->comm)
    return;
  for (int r = 0; r < ctx->comm->size && r < kMaxPeers; ++r)
    if (ctx->comm->peers_ready && r != ctx->comm->rank && ctx->comm->peer_base[r])
      cudaIpcCloseMemHandle(ctx->comm->peer_base[r]);
  cudaFree(ctx->comm->d_exchange);
  cudaFree(ctx->comm->d_error);
  NcclApi* api = nccl_api();
  if (api && ctx->comm->comm && api->CommDestroy)
    api->CommDestroy(ctx->comm->comm);
  cudaFree(ctx->comm->d_small);
  delete ctx->comm;
  ctx->comm = nullptr;
}

int comm_allreduce_device(itsolv_ctx* ctx, double* d, size_t count, bool op_max) {
  if (!ctx->comm || ctx->comm->size == 1 || count == 0)
    return 0;
  NcclApi* api = nccl_api();
  ITSOLV_NCCL(api->AllReduce(d, d, count, ncclDouble, op_max ? ncclMax : ncclSum, ctx->comm->comm, ctx->stream));
  return 0;
}

int comm_allgather_device(itsolv_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank) {
  if (!ctx->comm || ctx->comm->size == 1) {
    if (send != recv)
      ITSOLV_CUDA(cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
  }
  NcclApi* api = nccl_api();
  ITSOLV_NCCL(api->AllGather(send, recv, bytes_per_rank, ncclUint8, ctx->comm->comm, ctx->stream));
  return 0;
}

} // namespace itsolv

using namespace itsolv;

extern "C" {

int itsolv_comm_unique_id(void* id) {
  NcclApi* api = nccl_api();
  ITSOLV_REQUIRE(api, "itsolv_comm_unique_id: libnccl.so.2 could not be loaded");
  static_assert(sizeof(ncclUniqueId) == ITSOLV_UNIQUE_ID_BYTES, "unique id size");
  ncclUniqueId uid;
  ITSOLV_NCCL(api->GetUniqueId(&uid));
  std::memcpy(id, &uid, sizeof(uid));
  return 0;
}

int itsolv_comm_init(itsolv_ctx* ctx, int rank, int nranks, const void* id) {
  ITSOLV_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "itsolv_comm_init: bad rank/size");
  comm_destroy(ctx);
  auto* c = new Comm();
  c->rank = rank;
  c->size = nranks;
  ctx->comm = c;
  ITSOLV_CUDA(cudaSetDevice(ctx->device));
  ITSOLV_CUDA(cudaMalloc(&c->d_small, size_t(ITSOLV_MAX_PANEL) * ITSOLV_MAX_PANEL * sizeof(double)));
  if (nranks == 1)
    return 0;
  NcclApi* api = nccl_api();
  ITSOLV_REQUIRE(api, "itsolv_comm_init: libnccl.so.2 could not be loaded");
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  ITSOLV_NCCL(api->CommInitRank(&c->comm, nranks, uid, rank));
  return 0;
}

int itsolv_comm_p2p_export(itsolv_ctx* ctx, void* handle) {
  ITSOLV_REQUIRE(ctx->comm, "itsolv_comm_p2p_export: no communicator");
  Comm* c = ctx->comm;
  ITSOLV_REQUIRE(c->size <= kMaxPeers, "itsolv_comm_p2p_export: more ranks than one NVSwitch box holds");
  if (!c->d_exchange) {
    ITSOLV_CUDA(cudaMalloc(&c->d_exchange, c->exchange_bytes()));
    ITSOLV_CUDA(cudaMemset(c->d_exchange, 0, c->exchange_bytes()));
    ITSOLV_CUDA(cudaMalloc(&c->d_error, 64));
    ITSOLV_CUDA(cudaMemset(c->d_error, 0, 64));
    if (const char* t = std::getenv("ITSOLV_P2P_TIMEOUT_S")) // seconds a rank waits for its peers; 0 = for ever
      c->timeout_ns = static_cast<unsigned long long>(std::atof(t) * 1e9);
    ITSOLV_CUDA(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  ITSOLV_CUDA(cudaIpcGetMemHandle(&h, c->d_exchange));
  static_assert(sizeof(h) == ITSOLV_IPC_HANDLE_BYTES, "ipc handle size");
  std::memcpy(handle, &h, sizeof(h));
  return 0;
}

int itsolv_comm_p2p_import(itsolv_ctx* ctx, const void* handles) {
  ITSOLV_REQUIRE(ctx->comm && ctx->comm->d_exchange, "itsolv_comm_p2p_import: export first");
  Comm* c = ctx->comm;
  for (int r = 0; r < c->size; ++r) {
    if (r == c->rank) {
      c->peer_base[r] = c->d_exchange;
      continue;
    }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const char*>(handles) + size_t(r) * ITSOLV_IPC_HANDLE_BYTES, sizeof(h));
    const cudaError_t err = cudaIpcOpenMemHandle(&c->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess) {
      // no peer access to rank r (another node, no NVLink/PCIe P2P): unmap what was mapped so far and leave the
      // communicator on its ncclAllReduce path
      cudaGetLastError();
      for (int q = 0; q < r; ++q)
        if (q != c->rank && c->peer_base[q])
          cudaIpcCloseMemHandle(c->peer_base[q]);
      for (int q = 0; q < kMaxPeers; ++q)
        c->peer_base[q] = nullptr;
      set_error(std::string("itsolv_comm_p2p_import: rank ") + std::to_string(r) + ": " + cudaGetErrorString(err));
      return 1;
    }
  }
  c->peers_ready = true;
  return 0;
}

int itsolv_comm_p2p_disable(itsolv_ctx* ctx) {
  Comm* c = ctx->comm;
  if (!c)
    return 0;
  ITSOLV_CUDA(cudaStreamSynchronize(ctx->stream));
  if (c->peers_ready)
    for (int r = 0; r < c->size && r < kMaxPeers; ++r)
      if (r != c->rank && c->peer_base[r])
        cudaIpcCloseMemHandle(c->peer_base[r]);
  for (int q = 0; q < kMaxPeers; ++q)
    c->peer_base[q] = nullptr;
  c->peers_ready = false;
  return 0;
}

int itsolv_comm_rank(itsolv_ctx* ctx) { return ctx->comm ? ctx->comm->rank : 0; }
int itsolv_comm_size(itsolv_ctx* ctx) { return ctx->comm ? ctx->comm->size : 1; }

int itsolv_comm_allreduce_host(itsolv_ctx* ctx, double* values, size_t count, int op_max) {
  if (!ctx->comm || ctx->comm->size == 1 || count == 0)
    return 0;
  ITSOLV_REQUIRE(count <= size_t(ITSOLV_MAX_PANEL) * ITSOLV_MAX_PANEL, "itsolv_comm_allreduce_host: too many values");
  double* d = ctx->comm->d_small;
  ITSOLV_CUDA(cudaMemcpyAsync(d, values, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (comm_allreduce_device(ctx, d, count, op_max != 0))
    return 1;
  ITSOLV_CUDA(cudaMemcpyAsync(values, d, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  ITSOLV_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int itsolv_comm_barrier(itsolv_ctx* ctx) {
  double one = 1;
  return itsolv_comm_allreduce_host(ctx, &one, 1, 0);
}

int itsolv_comm_halo_exchange_multi(itsolv_ctx* ctx, const double* const* x, int w, size_t nloc, int b, double* out) {
  ++ctx->write_epoch;
  if (!ctx->comm || ctx->comm->size == 1 || b <= 0 || w <= 0)
    return 0;
  Comm* c = ctx->comm;
  ITSOLV_REQUIRE(nloc >= size_t(b), "itsolv_comm_halo_exchange_multi: shard shorter than the half bandwidth");
  if (c->peers_ready && ctx->opt_p2p_allreduce >= 0 && ctx->opt_p2p_halo >= 0 && w <= ITSOLV_MAX_ROOTS_HALO &&
      size_t(w) * size_t(b) <= kHaloSlotDoubles) {
    HaloParams p{};
    for (int k = 0; k < w; ++k)
      p.x[k] = x[k];
    p.out = out;
    auto data_of = [&](int r) { return reinterpret_cast<double*>(static_cast<char*>(c->peer_base[r]) + c->halo_data_offset()); };
    auto flags_of = [&](int r) {
      return reinterpret_cast<unsigned long long*>(static_cast<char*>(c->peer_base[r]) + c->halo_flag_offset());
    };
    p.below_data = c->rank > 0 ? data_of(c->rank - 1) : nullptr;
    p.below_flags = c->rank > 0 ? flags_of(c->rank - 1) : nullptr;
    p.above_data = c->rank < c->size - 1 ? data_of(c->rank + 1) : nullptr;
    p.above_flags = c->rank < c->size - 1 ? flags_of(c->rank + 1) : nullptr;
    p.my_data = data_of(c->rank);
    p.my_flags = flags_of(c->rank);
    p.error = c->d_error;
    p.timeout_ns = c->timeout_ns;
    p.seq = ++c->halo_seq;
    p.nloc = nloc;
    p.w = w;
    p.b = b;
    halo_exchange_kernel<<<1, 256, 0, ctx->stream>>>(p);
    ITSOLV_CUDA(cudaGetLastError());
    ctx->counters.launches += 1;
    return 0;
  }
  // no peer mapping: one grouped ncclSend/ncclRecv exchange per vector
  for (int k = 0; k < w; ++k) {
    double* h = out + size_t(k) * 2 * size_t(b);
    if (itsolv_comm_halo_exchange(ctx, x[k], x[k] + (nloc - size_t(b)), h, h + b, size_t(b)))
      return 1;
  }
  return 0;
}

int itsolv_comm_halo_exchange(itsolv_ctx* ctx, const double* send_lo, const double* send_hi, double* recv_lo,
                              double* recv_hi, size_t count) {
  if (!ctx->comm || ctx->comm->size == 1 || count == 0)
    return 0;
  NcclApi* api = nccl_api();
  const int r = ctx->comm->rank, p = ctx->comm->size;
  ITSOLV_NCCL(api->GroupStart());
  if (r > 0) {
    ITSOLV_NCCL(api->Send(send_lo, count, ncclDouble, r - 1, ctx->comm->comm, ctx->stream));
    ITSOLV_NCCL(api->Recv(recv_lo, count, ncclDouble, r - 1, ctx->comm->comm, ctx->stream));
  }
  if (r < p - 1) {
    ITSOLV_NCCL(api->Send(send_hi, count, ncclDouble, r + 1, ctx->comm->comm, ctx->stream));
    ITSOLV_NCCL(api->Recv(recv_hi, count, ncclDouble, r + 1, ctx->comm->comm, ctx->stream));
  }
  ITSOLV_NCCL(api->GroupEnd());
  return 0;
}

} // extern "C"
