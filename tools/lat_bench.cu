// Dependent-issue latencies that bound the Cholesky critical path (one warp, clock64 around a chain).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/lat_bench tools/lat_bench.cu && tools/bin/lat_bench
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
__device__ __forceinline__ long long clk(double& dep) { long long t; asm volatile("{\n\t.reg .f64 tmp;\n\tmov.f64 tmp, %1;\n\tmov.u64 %0, %%clock64;\n\t}" : "=l"(t), "+d"(dep) :: "memory"); return t; }
__device__ __forceinline__ double rsq_approx(double x) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
__device__ __forceinline__ double rsq_nr(double d) {      // MUFU seed + one cubic step (what rsqrt() does, without its slow path)
  const double y0 = rsq_approx(d);
  const double e = fma(-d * y0, y0, 1.0);
  const double p = fma(e, 0.375, 0.5);
  return fma(y0 * e, p, y0);
}
__global__ void k(double* out, long long* cyc, double seed) {
  __shared__ double sx[64];
  const int lane = threadIdx.x;
  double x = seed + lane * 1e-3, y = 1.0000001;
  long long t0, t1;
  int s = 0;
  // 0: DFMA chain
  t0 = clk(x);
#pragma unroll
  for (int i = 0; i < N; ++i) x = fma(x, y, 1e-9);
  t1 = clk(x); cyc[s++] = t1 - t0;
  // 1: DMUL chain
  t0 = clk(x);
#pragma unroll
  for (int i = 0; i < N; ++i) x = x * y;
  t1 = clk(x); cyc[s++] = t1 - t0;
  // 2: 64-bit shuffle chain
  t0 = clk(x);
#pragma unroll
  for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
  t1 = clk(x); cyc[s++] = t1 - t0;
  // 3: STS + syncwarp + LDS round trip
  t0 = clk(x);
#pragma unroll
  for (int i = 0; i < N; ++i) { sx[lane] = x; __syncwarp(); x = sx[(lane + 1) & 31]; __syncwarp(); }
  t1 = clk(x); cyc[s++] = t1 - t0;
  // 4: library rsqrt chain
  x = fabs(x) + 1.0;
  t0 = clk(x);
#pragma unroll
  for (int i = 0; i < N; ++i) x = rsqrt(x) + 1.0;
  t1 = clk(x); cyc[s++] = t1 - t0;
  // 5: MUFU.RSQ64H seed only chain (+ DADD)
  t0 = clk(x);
#pragma unroll
  for (int i = 0; i < N; ++i) x = rsq_approx(x) + 1.0;
  t1 = clk(x); cyc[s++] = t1 - t0;
  // 6: branch-free Newton rsqrt chain (+ DADD)
  t0 = clk(x);
#pragma unroll
  for (int i = 0; i < N; ++i) x = rsq_nr(x) + 1.0;
  t1 = clk(x); cyc[s++] = t1 - t0;
  // 7: dependent DMMA chain
  double c0 = x, c1 = y;
  t0 = clk(x);
#pragma unroll
  for (int i = 0; i < N; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(1e-3), "d"(1e-3));
  t1 = clk(c0); cyc[s++] = t1 - t0;
  // 8: 4 independent DFMA chains (throughput check)
  double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3;
  t0 = clk(x);
#pragma unroll
  for (int i = 0; i < N; ++i) { a0 = fma(a0, y, 1e-9); a1 = fma(a1, y, 1e-9); a2 = fma(a2, y, 1e-9); a3 = fma(a3, y, 1e-9); }
  a0 += a1 + a2 + a3;
  t1 = clk(a0); cyc[s++] = t1 - t0;
  // 9: LDS broadcast latency chain (address depends on the loaded value)
  sx[lane] = 0.0; sx[lane + 32] = 0.0; __syncwarp();
  int idx = 0; double acc = 0;
  t0 = clk(x);
#pragma unroll
  for (int i = 0; i < N; ++i) { const double v = sx[idx]; idx = (int)v; acc += v; }
  t1 = clk(acc); cyc[s++] = t1 - t0;
  // 10: FP32 rsqrt seed + 2 Newton steps in FP64
  t0 = clk(x);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double yy = (double)rsqrtf((float)x);
    yy = yy * fma(-0.5 * x * yy, yy, 1.5);
    yy = yy * fma(-0.5 * x * yy, yy, 1.5);
    x = yy + 1.0;
  }
  t1 = clk(x); cyc[s++] = t1 - t0;
  out[lane] = x + c0 + c1 + a0 + a1 + a2 + a3 + acc;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 16 * 8);
  for (int r = 0; r < 2; ++r) k<<<1, 32>>>(out, cyc, 1.5);
  long long h[16];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[] = {"DFMA dependent", "DMUL dependent", "SHFL.64 dependent", "STS+syncwarp+LDS+syncwarp", "rsqrt() + DADD", "MUFU.RSQ64H + DADD",
                         "branch-free NR rsqrt + DADD", "DMMA.884 dependent", "4 independent DFMA (per 4)", "LDS dependent (+cvt)", "rsqrtf seed + 2 NR + DADD"};
  printf("{");
  for (int s = 0; s < 11; ++s) printf("\"%s\": %.1f%s", names[s], (double)h[s] / N, s < 10 ? ", " : "");
  printf("}\n%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
