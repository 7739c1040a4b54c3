// FP64 peak of the device this runs on: DFMA (the FMA pipe the register-tile Gram kernels and the expansion kernels
// use) and DMMA (mma.sync.aligned.m8n8k4.f64, the FP64 tensor-core path of gemm_inner_mma.cu). Gives the compute-bound
// panels (64 x 64 Gram blocks, 16-root residuals over 40 vector pairs) a roofline denominator: BASELINE.md section 2 lists
// it as "not measured yet". Prints one JSON line.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu && tools/fp64_peak
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

constexpr int kIters = 4096;
constexpr int kChains = 16; // independent accumulators per thread: covers the DFMA latency at 8+ warps per scheduler

__global__ void __launch_bounds__(256) dfma_kernel(double* out, double a, double b) {
  double acc[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c)
    acc[c] = double(threadIdx.x + c);
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int c = 0; c < kChains; ++c)
      acc[c] = fma(acc[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < kChains; ++c)
    s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dmma_kernel(double* out, double a, double b) {
  constexpr int kTiles = 8; // independent 8x8 accumulator tiles per warp
  double c0[kTiles], c1[kTiles];
#pragma unroll
  for (int t = 0; t < kTiles; ++t) {
    c0[t] = double(threadIdx.x + t);
    c1[t] = 0.5;
  }
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int t = 0; t < kTiles; ++t)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[t]), "+d"(c1[t])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < kTiles; ++t)
    s += c0[t] + c1[t];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
static double time_ms(K launch, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  return best;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  const int blocks = sms * 8, threads = 256;
  double* out;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  const double t_fma = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, 0.999999, 1e-9); }, 10);
  const double t_mma = time_ms([&] { dmma_kernel<<<blocks, threads>>>(out, 0.999999, 1e-9); }, 10);
  const double flop_fma = 2.0 * kIters * kChains * double(blocks) * threads;
  const double flop_mma = 2.0 * 8 * 8 * 4 * 8.0 * kIters * double(blocks) * (threads / 32); // m8n8k4, 8 tiles per warp
  int clock_khz = 0;
  cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
  std::printf("{\"device\": \"%s\", \"sms\": %d, \"sm_clock_mhz_max\": %.0f, \"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, "
              "\"dfma_ms\": %.3f, \"dmma_ms\": %.3f, \"dfma_per_sm_per_clk_at_max_clock\": %.1f}\n",
              prop.name, sms, clock_khz / 1e3, flop_fma / t_fma / 1e9, flop_mma / t_mma / 1e9, t_fma, t_mma,
              flop_fma / 2.0 / (t_fma * 1e-3) / sms / (clock_khz * 1e3));
  cudaFree(out);
  return 0;
}
