// Context, memory pool, staging ring, accounting. No arithmetic here.
#include <cstdlib>
#include <cstring>
#include <string>

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include "common.cuh"

namespace itsolv {

//! ITSOLV_BACKTRACE=1: print the native call stack when the process receives SIGSEGV / SIGABRT (diagnostics on boxes
//! without a debugger), then let the default action run
static void backtrace_handler(int sig) {
  void* frames[64];
  const int n = backtrace(frames, 64);
  const char msg[] = "itsolv_b200: fatal signal, native backtrace:\n";
  if (write(2, msg, sizeof(msg) - 1) < 0) {
  }
  backtrace_symbols_fd(frames, n, 2);
  signal(sig, SIG_DFL);
  raise(sig);
}
static void install_backtrace_handler() {
  static bool done = false;
  const char* e = std::getenv("ITSOLV_BACKTRACE");
  if (done || !e || !*e || *e == '0')
    return;
  done = true;
  signal(SIGSEGV, backtrace_handler);
  signal(SIGABRT, backtrace_handler);
  signal(SIGBUS, backtrace_handler);
}

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }

static int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v && *v ? std::atoi(v) : dflt;
}

CallScope::CallScope(itsolv_ctx* c, int cls_, double bytes) : ctx(c), cls(cls_) {
  ctx->counters.bytes += bytes;
  switch (cls) {
  case OP_GEMM_INNER:
    ctx->counters.bytes_gemm_inner += bytes;
    ctx->counters.calls_gemm_inner++;
    break;
  case OP_GEMM_OUTER:
    ctx->counters.bytes_gemm_outer += bytes;
    ctx->counters.calls_gemm_outer++;
    break;
  case OP_BLAS1:
    ctx->counters.bytes_blas1 += bytes;
    ctx->counters.calls_blas1++;
    break;
  case OP_RESIDUAL:
    ctx->counters.bytes_residual += bytes;
    ctx->counters.calls_residual++;
    break;
  default:
    break;
  }
  if (ctx->profiling && !ctx->active_scope) {
    armed = true;
    ctx->active_scope = this;
  }
}

void mark_launch(itsolv_ctx* ctx) {
  ++ctx->write_epoch; // conservative: any launch may have written a vector
  CallScope* s = ctx->active_scope;
  if (!s || !s->armed || s->start)
    return;
  if (ctx->event_pool.empty()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess)
      return;
    ctx->event_pool.push_back(e);
  }
  s->start = ctx->event_pool.back();
  ctx->event_pool.pop_back();
  cudaEventRecord(s->start, ctx->stream);
}

CallScope::~CallScope() { stop(); }

void CallScope::stop() {
  if (armed && ctx->active_scope == this)
    ctx->active_scope = nullptr;
  armed = false;
  if (!start)
    return;
  cudaEvent_t stop;
  if (ctx->event_pool.empty()) {
    if (cudaEventCreate(&stop) != cudaSuccess)
      return;
  } else {
    stop = ctx->event_pool.back();
    ctx->event_pool.pop_back();
  }
  cudaEventRecord(stop, ctx->stream);
  ctx->pending.push_back({start, stop, cls});
  start = nullptr;
  if (ctx->pending.size() > 4096)
    drain_pending(ctx);
}

void drain_pending(itsolv_ctx* ctx) {
  if (ctx->pending.empty())
    return;
  cudaEventSynchronize(ctx->pending.back().stop);
  for (auto& p : ctx->pending) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, p.start, p.stop) == cudaSuccess) {
      const double s = ms * 1e-3;
      ctx->counters.device_seconds += s;
      if (p.cls == OP_GEMM_INNER)
        ctx->counters.seconds_gemm_inner += s;
      else if (p.cls == OP_GEMM_OUTER)
        ctx->counters.seconds_gemm_outer += s;
      else if (p.cls == OP_BLAS1)
        ctx->counters.seconds_blas1 += s;
      else if (p.cls == OP_RESIDUAL)
        ctx->counters.seconds_residual += s;
    }
    ctx->event_pool.push_back(p.start);
    ctx->event_pool.push_back(p.stop);
  }
  ctx->pending.clear();
}

int stage_acquire(itsolv_ctx* ctx, size_t bytes, char** host, char** dev, int* slot) {
  ITSOLV_REQUIRE(bytes <= ctx->stage_slot_bytes, "itsolv: staged payload too large");
  const int s = ctx->stage_next;
  ctx->stage_next = (s + 1) % ctx->stage_slots;
  // the slot's previous copy (and the kernel that consumed the device side) must have finished before the host overwrites it
  ITSOLV_CUDA(cudaEventSynchronize(ctx->stage_events[s]));
  *host = ctx->h_stage + size_t(s) * ctx->stage_slot_bytes;
  *dev = ctx->d_stage + size_t(s) * ctx->stage_slot_bytes;
  *slot = s;
  return 0;
}

int stage_commit(itsolv_ctx* ctx, int slot, size_t bytes) {
  char* host = ctx->h_stage + size_t(slot) * ctx->stage_slot_bytes;
  char* dev = ctx->d_stage + size_t(slot) * ctx->stage_slot_bytes;
  ITSOLV_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}

int stage_done(itsolv_ctx* ctx, int slot) {
  ITSOLV_CUDA(cudaEventRecord(ctx->stage_events[slot], ctx->stream));
  return 0;
}

int ensure_partials(itsolv_ctx* ctx, size_t doubles) {
  if (doubles <= ctx->partials_capacity)
    return 0;
  if (ctx->d_partials) {
    ITSOLV_CUDA(cudaStreamSynchronize(ctx->stream));
    ITSOLV_CUDA(cudaFree(ctx->d_partials));
    ctx->d_partials = nullptr;
    ctx->partials_capacity = 0;
  }
  ITSOLV_CUDA(cudaMalloc(&ctx->d_partials, doubles * sizeof(double)));
  ctx->partials_capacity = doubles;
  return 0;
}

int ensure_dynamic_smem(itsolv_ctx* ctx, const void* kernel, size_t bytes) {
  if (bytes <= 48 * 1024)
    return 0;
  for (auto& e : ctx->smem_optin)
    if (e.first == kernel) {
      if (e.second >= bytes)
        return 0;
      ITSOLV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
      e.second = bytes;
      return 0;
    }
  ITSOLV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
  ctx->smem_optin.emplace_back(kernel, bytes);
  return 0;
}

static int ctx_init(itsolv_ctx* ctx, int device, cudaStream_t stream, bool own) {
  install_backtrace_handler();
  ITSOLV_CUDA(cudaSetDevice(device));
  ctx->device = device;
  cudaDeviceProp prop;
  ITSOLV_CUDA(cudaGetDeviceProperties(&prop, device));
  ITSOLV_REQUIRE(prop.major >= 10, "itsolv_b200 requires a Blackwell (sm_100a) device");
  ctx->num_sms = prop.multiProcessorCount;
  ctx->max_smem_optin = int(prop.sharedMemPerBlockOptin);
  if (own) {
    ITSOLV_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  } else {
    ctx->stream = stream;
  }
  ITSOLV_CUDA(cudaDeviceGetDefaultMemPool(&ctx->pool, device));
  uint64_t keep = UINT64_MAX; // never trim: freed Q/D vectors are reused by the next iteration
  ITSOLV_CUDA(cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep));
  const size_t panel2 = size_t(ITSOLV_MAX_PANEL) * ITSOLV_MAX_PANEL;
  ITSOLV_CUDA(cudaMalloc(&ctx->d_result, panel2 * sizeof(double)));
  ITSOLV_CUDA(cudaHostAlloc(&ctx->h_result, panel2 * sizeof(double), cudaHostAllocMapped));
  ITSOLV_CUDA(cudaHostAlloc(&ctx->h_flag, 64, cudaHostAllocMapped));
  *ctx->h_flag = 0;
  ITSOLV_CUDA(cudaMalloc(&ctx->d_counter, sizeof(unsigned int)));
  ITSOLV_CUDA(cudaMemset(ctx->d_counter, 0, sizeof(unsigned int)));
  ctx->stage_slot_bytes = panel2 * sizeof(double) + 8192;
  ctx->stage_slots = 8;
  ITSOLV_CUDA(cudaHostAlloc(&ctx->h_stage, ctx->stage_slot_bytes * ctx->stage_slots, cudaHostAllocDefault));
  ITSOLV_CUDA(cudaMalloc(&ctx->d_stage, ctx->stage_slot_bytes * ctx->stage_slots));
  ctx->stage_events.resize(ctx->stage_slots);
  for (auto& e : ctx->stage_events) {
    ITSOLV_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ITSOLV_CUDA(cudaEventRecord(e, ctx->stream));
  }
  ITSOLV_CUDA(cudaMalloc(&ctx->d_select, 4096 * sizeof(unsigned long long)));
  for (int t = 0; t < ITSOLV_TIMERS; ++t) {
    ITSOLV_CUDA(cudaEventCreate(&ctx->timer_start[t]));
    ITSOLV_CUDA(cudaEventCreate(&ctx->timer_stop[t]));
  }
  if (ensure_partials(ctx, size_t(4) * ctx->num_sms * 4096))
    return 1;
  ctx->opt_gi_rows = env_int("ITSOLV_GI_ROWS", 0);
  ctx->opt_gi_stages = env_int("ITSOLV_GI_STAGES", 0);
  ctx->opt_gi_threads = env_int("ITSOLV_GI_THREADS", 0);
  ctx->opt_gi_tile = env_int("ITSOLV_GI_TILE", 0);
  ctx->opt_gi_ctas = env_int("ITSOLV_GI_CTAS", 0);
  ctx->opt_gi_chunk = env_int("ITSOLV_GI_CHUNK", 0);
  ctx->opt_gi_nprod = env_int("ITSOLV_GI_NPROD", 0);
  ctx->opt_gi_loader = env_int("ITSOLV_GI_LOADER", 0);
  ctx->opt_gi_direct = env_int("ITSOLV_GI_DIRECT", 0);
  ctx->opt_gi_direct_ctas = env_int("ITSOLV_GI_DIRECT_CTAS", 0);
  ctx->opt_gi_mma = env_int("ITSOLV_GI_MMA", 0);
  ctx->opt_go_cols = env_int("ITSOLV_GO_COLS", 0);
  ctx->opt_go_ctas = env_int("ITSOLV_GO_CTAS", 0);
  ctx->opt_blas1_ctas = env_int("ITSOLV_BLAS1_CTAS", 0);
  ctx->opt_p2p_allreduce = env_int("ITSOLV_P2P_ALLREDUCE", 0);
  ctx->opt_p2p_halo = env_int("ITSOLV_P2P_HALO", 0);
  ctx->opt_project_chain = env_int("ITSOLV_PROJECT_CHAIN", 0);
  ctx->opt_ds_ring = env_int("ITSOLV_DS_RING", 0);
  ctx->opt_mgs_chain = env_int("ITSOLV_MGS_CHAIN", 0);
  return 0;
}

void comm_destroy(itsolv_ctx* ctx);

} // namespace itsolv

using namespace itsolv;

extern "C" {

const char* itsolv_last_error(void) { return g_error.c_str(); }

int itsolv_ctx_create(int device, itsolv_ctx** out) {
  auto* ctx = new itsolv_ctx();
  if (int rc = ctx_init(ctx, device, nullptr, true)) {
    delete ctx;
    return rc;
  }
  *out = ctx;
  return 0;
}

int itsolv_ctx_create_on_stream(int device, void* cuda_stream, itsolv_ctx** out) {
  auto* ctx = new itsolv_ctx();
  if (int rc = ctx_init(ctx, device, static_cast<cudaStream_t>(cuda_stream), false)) {
    delete ctx;
    return rc;
  }
  *out = ctx;
  return 0;
}

void itsolv_ctx_destroy(itsolv_ctx* ctx) {
  if (!ctx)
    return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  drain_pending(ctx);
  comm_destroy(ctx);
  for (auto e : ctx->event_pool)
    cudaEventDestroy(e);
  for (auto e : ctx->stage_events)
    cudaEventDestroy(e);
  for (int t = 0; t < ITSOLV_TIMERS; ++t) {
    cudaEventDestroy(ctx->timer_start[t]);
    cudaEventDestroy(ctx->timer_stop[t]);
  }
  cudaFree(ctx->d_partials);
  cudaFree(ctx->d_result);
  cudaFreeHost(ctx->h_result);
  cudaFreeHost(ctx->h_flag);
  cudaFree(ctx->d_counter);
  cudaFreeHost(ctx->h_stage);
  cudaFree(ctx->d_stage);
  cudaFree(ctx->d_select);
  cudaFree(ctx->d_select_cand);
  if (ctx->own_stream)
    cudaStreamDestroy(ctx->stream);
  delete ctx;
}

unsigned long long itsolv_ctx_write_epoch(itsolv_ctx* ctx) { return ctx->write_epoch; }
void itsolv_ctx_note_write(itsolv_ctx* ctx) { ++ctx->write_epoch; }
void* itsolv_ctx_stream(itsolv_ctx* ctx) { return ctx->stream; }
int itsolv_ctx_device(itsolv_ctx* ctx) { return ctx->device; }

int itsolv_ctx_synchronize(itsolv_ctx* ctx) {
  ITSOLV_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int itsolv_ctx_set_option(itsolv_ctx* ctx, const char* name, int value) {
  struct {
    const char* n;
    int* p;
  } table[] = {{"GI_ROWS", &ctx->opt_gi_rows},       {"GI_STAGES", &ctx->opt_gi_stages},
               {"GI_THREADS", &ctx->opt_gi_threads}, {"GI_TILE", &ctx->opt_gi_tile},
               {"GI_CTAS", &ctx->opt_gi_ctas},       {"GI_CHUNK", &ctx->opt_gi_chunk},
               {"GI_NPROD", &ctx->opt_gi_nprod},     {"GI_LOADER", &ctx->opt_gi_loader},
               {"GI_DIRECT", &ctx->opt_gi_direct},   {"GI_DIRECT_CTAS", &ctx->opt_gi_direct_ctas},
               {"GI_MMA", &ctx->opt_gi_mma},       {"GO_COLS", &ctx->opt_go_cols},
               {"GO_CTAS", &ctx->opt_go_ctas},       {"BLAS1_CTAS", &ctx->opt_blas1_ctas},
               {"P2P_ALLREDUCE", &ctx->opt_p2p_allreduce}, {"P2P_HALO", &ctx->opt_p2p_halo}, {"PROJECT_CHAIN", &ctx->opt_project_chain}, {"DS_RING", &ctx->opt_ds_ring},
               {"MGS_CHAIN", &ctx->opt_mgs_chain}};
  for (auto& t : table)
    if (std::strcmp(t.n, name) == 0) {
      const int prev = *t.p;
      *t.p = value;
      return prev;
    }
  return -1;
}

void itsolv_ctx_counters(itsolv_ctx* ctx, itsolv_counters* out) {
  drain_pending(ctx);
  *out = ctx->counters;
}
void itsolv_ctx_reset_counters(itsolv_ctx* ctx) {
  drain_pending(ctx);
  ctx->counters = itsolv_counters{};
}
void itsolv_ctx_set_profiling(itsolv_ctx* ctx, int enabled) {
  drain_pending(ctx);
  ctx->profiling = enabled != 0;
}

int itsolv_ctx_timer_start(itsolv_ctx* ctx, int id) {
  ITSOLV_REQUIRE(id >= 0 && id < ITSOLV_TIMERS, "itsolv_ctx_timer_start: bad timer id");
  ITSOLV_CUDA(cudaEventRecord(ctx->timer_start[id], ctx->stream));
  return 0;
}
int itsolv_ctx_timer_stop(itsolv_ctx* ctx, int id, double* milliseconds) {
  ITSOLV_REQUIRE(id >= 0 && id < ITSOLV_TIMERS, "itsolv_ctx_timer_stop: bad timer id");
  ITSOLV_CUDA(cudaEventRecord(ctx->timer_stop[id], ctx->stream));
  ITSOLV_CUDA(cudaEventSynchronize(ctx->timer_stop[id]));
  float ms = 0;
  ITSOLV_CUDA(cudaEventElapsedTime(&ms, ctx->timer_start[id], ctx->timer_stop[id]));
  *milliseconds = ms;
  return 0;
}

int itsolv_alloc(itsolv_ctx* ctx, size_t n, double** out) {
  ++ctx->write_epoch;
  void* p = nullptr;
  const size_t bytes = (n ? n : 1) * sizeof(double);
  ITSOLV_CUDA(cudaMallocAsync(&p, bytes, ctx->stream));
  *out = static_cast<double*>(p);
  ctx->alloc_bytes[p] = bytes;
  ctx->live_bytes += bytes;
  ctx->peak_bytes = ctx->live_bytes > ctx->peak_bytes ? ctx->live_bytes : ctx->peak_bytes;
  return 0;
}
int itsolv_free(itsolv_ctx* ctx, double* p) {
  ++ctx->write_epoch;
  if (p) {
    auto it = ctx->alloc_bytes.find(p);
    if (it != ctx->alloc_bytes.end()) {
      ctx->live_bytes -= it->second;
      ctx->alloc_bytes.erase(it);
    }
    ITSOLV_CUDA(cudaFreeAsync(p, ctx->stream));
  }
  return 0;
}
int itsolv_mem_usage(itsolv_ctx* ctx, size_t* live_bytes, size_t* peak_bytes, int reset_peak) {
  *live_bytes = ctx->live_bytes;
  *peak_bytes = ctx->peak_bytes;
  if (reset_peak)
    ctx->peak_bytes = ctx->live_bytes;
  return 0;
}
int itsolv_upload(itsolv_ctx* ctx, double* dst, const double* src, size_t n) {
  ++ctx->write_epoch;
  ITSOLV_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  ITSOLV_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int itsolv_upload_bytes(itsolv_ctx* ctx, void* dst, const void* src, size_t bytes) {
  ++ctx->write_epoch;
  ITSOLV_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ITSOLV_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int itsolv_download(itsolv_ctx* ctx, double* dst, const double* src, size_t n) {
  // cudaMemcpyDefault: dst may be host memory or, for callers that keep results on the GPU, device memory
  ITSOLV_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDefault, ctx->stream));
  ITSOLV_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int itsolv_mem_trim(itsolv_ctx* ctx) {
  ITSOLV_CUDA(cudaStreamSynchronize(ctx->stream));
  ITSOLV_CUDA(cudaMemPoolTrimTo(ctx->pool, 0));
  return 0;
}
int itsolv_mem_info(itsolv_ctx* ctx, size_t* free_bytes, size_t* total_bytes) {
  ITSOLV_CUDA(cudaSetDevice(ctx->device));
  ITSOLV_CUDA(cudaMemGetInfo(free_bytes, total_bytes));
  return 0;
}

void itsolv_distribution(size_t n, int nranks, int64_t* borders) {
  // reference array/util/Distribution.h:99-110: block = n / P, the first n % P chunks hold one more element
  const size_t block = n / size_t(nranks), extra = n % size_t(nranks);
  borders[0] = 0;
  for (int r = 0; r < nranks; ++r)
    borders[r + 1] = borders[r] + int64_t(block + (size_t(r) < extra ? 1 : 0));
}

} // extern "C"
