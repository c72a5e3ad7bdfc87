// The in-CTA 64 x 64 factorisation (gingr_b200/csrc/chol_potrf.cuh) in isolation: cycles per call, checked against a host
// Cholesky.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/potrf_bench tools/potrf_bench.cu
#include <cmath>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
namespace gingr { namespace {
constexpr int TB = 64, TP = 68, DF_THREADS = 256;
#define DF_CLOCK(t, k) ((void)0)
#ifdef POTRF_PROFILE
__device__ long long g_prof[32 * 16];
#define PF_MARK(s, k) do { if ((threadIdx.x & 31) == 0) g_prof[(s) * 16 + (k)] = clock64(); } while (0)
#else
#define PF_MARK(s, k) ((void)0)
#endif
#include "../gingr_b200/csrc/chol_potrf.cuh"
__global__ void __launch_bounds__(256, 1) bench(const double* A, double* Lout, double* Zout, long long* cyc, int reps) {
  extern __shared__ __align__(16) double sm[];
  double* sT = sm;
  double2* G2 = reinterpret_cast<double2*>(sm + TB * TP);
  double* sDiag = sm + 2 * TB * TP;
  double* sSub = sDiag + 64;
  unsigned long long* sBar = reinterpret_cast<unsigned long long*>(sSub + 32);
  const int tid = threadIdx.x;
  if (tid < 32) mbar_init(&sBar[tid], 1);
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  __syncthreads();
  long long total = 0;
  for (int it = 0; it < reps; ++it) {
    for (int e = tid; e < 64 * 64; e += 256) sT[(e >> 6) * TP + (e & 63)] = A[e];
    __syncthreads();
    const long long t0 = clock64();
    potrf64(sT, G2, sDiag, sSub, sBar, (unsigned)(it & 1), tid, 1);
    const long long t1 = clock64();
    total += t1 - t0;
  }
  if (tid == 0 && blockIdx.x == 0) cyc[0] = total / reps;
  if (blockIdx.x == 0)
    for (int e = tid; e < 64 * 64; e += 256) {
      const int r = e >> 6, c = e & 63;
      Lout[e] = r >= c ? potrf_L(G2, sDiag, sSub, r, c) : 0.0;
      Zout[e] = potrf_Z(G2, r, c);
    }
}
} }
int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 1;
  std::vector<double> A(4096), L(4096, 0.0);
  for (int i = 0; i < 64; ++i)
    for (int j = 0; j < 64; ++j) A[i * 64 + j] = 1.0 / (1.0 + std::abs(i - j)) + (i == j ? 2.0 : 0.0);
  for (int j = 0; j < 64; ++j) {
    for (int i = j; i < 64; ++i) {
      double s = A[i * 64 + j];
      for (int k = 0; k < j; ++k) s -= L[i * 64 + k] * L[j * 64 + k];
      L[i * 64 + j] = (i == j) ? std::sqrt(s) : s / L[j * 64 + j];
    }
  }
  double *dA, *dL, *dZ; long long* dc;
  cudaMalloc(&dA, 4096 * 8); cudaMalloc(&dL, 4096 * 8); cudaMalloc(&dZ, 4096 * 8); cudaMalloc(&dc, 8);
  cudaMemcpy(dA, A.data(), 4096 * 8, cudaMemcpyHostToDevice);
  const size_t smem = (2 * 64 * 68 + 64 + 32 + 32) * 8 + 64;
  cudaFuncSetAttribute(gingr::bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  gingr::bench<<<grid, 256, smem>>>(dA, dL, dZ, dc, 20);
  gingr::bench<<<grid, 256, smem>>>(dA, dL, dZ, dc, 200);
  std::vector<double> hL(4096), hZ(4096); long long cyc = 0;
  cudaMemcpy(hL.data(), dL, 4096 * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(hZ.data(), dZ, 4096 * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
  double eL = 0, eZ = 0;
  for (int i = 0; i < 64; ++i)
    for (int j = 0; j <= i; ++j) eL = std::fmax(eL, std::fabs(hL[i * 64 + j] - L[i * 64 + j]));
  for (int i = 0; i < 64; ++i)
    for (int j = 0; j < 64; ++j) {   // Z L = I
      double s = 0;
      for (int k = 0; k < 64; ++k) s += hZ[i * 64 + k] * L[k * 64 + j];
      eZ = std::fmax(eZ, std::fabs(s - (i == j ? 1.0 : 0.0)));
    }
#ifdef POTRF_PROFILE
  {
    long long h[32 * 16];
    cudaMemcpyFromSymbol(h, gingr::g_prof, sizeof(h));
    for (int s2 = 8; s2 < 18; ++s2) {
      printf("step %d:", s2);
      for (int k = 0; k < 8; ++k) printf(" %lld", h[s2 * 16 + k] - h[8 * 16]);
      printf("\n");
    }
  }
#endif
  printf("{\"grid\": %d, \"cycles_per_potrf\": %lld, \"cycles_per_step\": %.1f, \"err_L\": %.2e, \"err_ZL\": %.2e, \"cuda\": \"%s\"}\n", grid, cyc,
         cyc / 32.0, eL, eZ, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
