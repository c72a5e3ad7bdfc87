// fp64_peaks.cu -- measures the FP64 roofline denominators MEASURED_PEAKS.json lacks:
//   (i)   DFMA peak: register-resident FMA chains on every SM
//   (ii)  DMMA peak: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) chains on every SM
//   (iii) both interleaved in one kernel -- tells whether DMMA and DFMA share a pipe on B200
// Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <algorithm>

template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) acc[k] = threadIdx.x * 1e-9 + k;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = fma(acc[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) { c0[k] = threadIdx.x * 1e-9 + k; c1[k] = k; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < ILP; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[k]), "+d"(c1[k]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += c0[k] + c1[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// interleave: per iteration ILP DMMAs and ILP*RATIO DFMAs
template <int ILP, int RATIO>
__global__ void __launch_bounds__(256) mixed_kernel(double* out, int iters, double a, double b) {
  double c0[ILP], c1[ILP], f[ILP * RATIO];
#pragma unroll
  for (int k = 0; k < ILP; ++k) { c0[k] = threadIdx.x * 1e-9 + k; c1[k] = k; }
#pragma unroll
  for (int k = 0; k < ILP * RATIO; ++k) f[k] = threadIdx.x * 1e-9 + k;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[k]), "+d"(c1[k]) : "d"(a), "d"(b));
#pragma unroll
      for (int q = 0; q < RATIO; ++q) f[k * RATIO + q] = fma(f[k * RATIO + q], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += c0[k] + c1[k];
#pragma unroll
  for (int k = 0; k < ILP * RATIO; ++k) s += f[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_best(F launch, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  return best;
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  const int blocks = sms * 8, threads = 256, iters = 4096;
  double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  constexpr int ILP = 8;
  auto l_dfma = [&]() { dfma_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); };
  auto l_dmma = [&]() { dmma_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); };
  auto l_mix = [&]() { mixed_kernel<4, 8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); };
  for (int w = 0; w < 3; ++w) { l_dfma(); l_dmma(); l_mix(); }
  cudaDeviceSynchronize();
  const float t_dfma = time_best(l_dfma, 10), t_dmma = time_best(l_dmma, 10), t_mix = time_best(l_mix, 10);
  // sustained: back to back for ~2 s
  auto sustained = [&](auto launch, double flop_per_launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int n = 0; cudaEventRecord(e0);
    float ms = 0;
    do { for (int k = 0; k < 20; ++k) launch(); n += 20; cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); } while (ms < 2000.f);
    return flop_per_launch * n / (ms * 1e-3) / 1e12;
  };
  const double total_threads = (double)blocks * threads;
  const double f_dfma = total_threads * iters * ILP * 2.0;
  const double f_dmma = (total_threads / 32.0) * iters * ILP * (8 * 8 * 4 * 2.0);
  const double f_mix_dmma = (total_threads / 32.0) * iters * 4 * (8 * 8 * 4 * 2.0);
  const double f_mix_dfma = total_threads * iters * 4 * 8 * 2.0;
  const double s_dfma = sustained(l_dfma, f_dfma), s_dmma = sustained(l_dmma, f_dmma);
  cudaError_t e = cudaDeviceSynchronize();
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, \"dfma_tflops\": %.3f, \"dmma_tflops\": %.3f, "
         "\"dfma_tflops_sustained\": %.3f, \"dmma_tflops_sustained\": %.3f, "
         "\"mixed_ms\": %.4f, \"mixed_dmma_alone_ms\": %.4f, \"mixed_dfma_alone_ms\": %.4f, \"mixed_total_tflops\": %.3f, "
         "\"cuda_error\": \"%s\"}\n",
         prop.name, sms, prop.clockRate, f_dfma / (t_dfma * 1e-3) / 1e12, f_dmma / (t_dmma * 1e-3) / 1e12, s_dfma, s_dmma,
         t_mix, f_mix_dmma / (f_dmma / t_dmma), f_mix_dfma / (f_dfma / t_dfma),
         (f_mix_dmma + f_mix_dfma) / (t_mix * 1e-3) / 1e12, cudaGetErrorString(e));
  return 0;
}
