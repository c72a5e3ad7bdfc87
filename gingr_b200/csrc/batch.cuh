// batch.cuh -- one kernel launch serving many chains (BASELINE config 5: independent Metropolis-Hastings chains).
//
// A chain's MH step is a fixed sequence of ~50 small kernels; 1024 chains as 1024 captured graphs are bound by the node
// rate of the graph front end (profiles/r02_small_problems.md), not by the math.  Here every kernel on that path exists
// in two forms generated from ONE body:
//     name<<<grid, block>>>(args)                     the chain alone (arguments in the parameter bank), and
//     name_batched<<<(grid.x, grid.y, chains), block>>>(args[chains])   blockIdx.z = chain, arguments from an array
// so that a batch of chains is the same kernel sequence with one launch per step of the sequence: the same blocks, the same
// arithmetic in the same order per chain (bit-identical to the chain alone), 1/chains of the launches.
//
// GINGR_KERNEL defines the pair; GINGR_LAUNCH launches the single form or, while a LaunchRecorder is installed on the
// ctx, records (batched entry point, grid, block, argument tuple) instead.  gingr_mcmc_batch (mcmc.cuh) records every
// chain's step, checks that the sequences agree, uploads the argument arrays once and replays the batched sequence as
// one CUDA graph.  The recording runs inside a stream capture: any launch / copy on the path that is not batch-aware lands
// in the (discarded) capture graph instead of executing, is detected by its node count, and the batch falls back to the
// per-chain graphs.
#pragma once
#include <cuda/std/tuple>
#include <string.h>

#include "common.cuh"

struct LaunchRecord {
  const void* fn;      // name_batched
  const char* name;
  dim3 grid, block;
  unsigned smem;
  unsigned arg_bytes;
  size_t arg_off;      // into LaunchRecorder::args
};

struct LaunchRecorder {
  std::vector<LaunchRecord> recs;
  std::vector<unsigned char> args;
  void add(const void* fn, const char* name, dim3 grid, dim3 block, size_t smem, const void* a, size_t bytes) {
    LaunchRecord r;
    r.fn = fn; r.name = name; r.grid = grid; r.block = block; r.smem = (unsigned)smem; r.arg_bytes = (unsigned)bytes;
    r.arg_off = args.size();
    args.resize(args.size() + bytes);
    memcpy(args.data() + r.arg_off, a, bytes);
    recs.push_back(r);
  }
};

template <typename F> struct gingr_args_of;
// by-value kernel parameters; a body may take a large struct as `const T&` and then reads it in place (parameter bank of
// the single form -- __grid_constant__ --, the argument array of the batched form) instead of copying it to local memory
template <typename T> struct gingr_decay { using type = T; };
template <typename T> struct gingr_decay<const T&> { using type = T; };
template <typename... A> struct gingr_args_of<void (*)(A...)> { using type = cuda::std::tuple<typename gingr_decay<A>::type...>; };

#define GINGR_KERNEL_IMPL(bounds_attr, name, ...)                                                                       \
  static __device__ __forceinline__ void name##_body(__VA_ARGS__);                                                      \
  using name##_args = typename gingr_args_of<decltype(&name##_body)>::type;                                             \
  __global__ void bounds_attr name(const __grid_constant__ name##_args a) { cuda::std::apply(name##_body, a); }         \
  __global__ void bounds_attr name##_batched(const name##_args* __restrict__ a) { cuda::std::apply(name##_body, a[blockIdx.z]); } \
  static __device__ __forceinline__ void name##_body(__VA_ARGS__)

// bounds in parentheses: GINGR_KERNEL((256), foo_kernel, int n, const double* x) { ... }
#define GINGR_KERNEL(bounds, name, ...) GINGR_KERNEL_IMPL(__launch_bounds__ bounds, name, __VA_ARGS__)
#define GINGR_KERNEL_NB(name, ...) GINGR_KERNEL_IMPL(, name, __VA_ARGS__)

#define GINGR_UNPAREN(...) __VA_ARGS__
// template kernels: GINGR_KERNEL_T((int K), (K), (256), foo_kernel, int n, ...) { ... }
#define GINGR_KERNEL_T(tdecl, targs, bounds, name, ...)                                                                 \
  template <GINGR_UNPAREN tdecl> static __device__ __forceinline__ void name##_body(__VA_ARGS__);                       \
  template <GINGR_UNPAREN tdecl>                                                                                        \
  using name##_args = typename gingr_args_of<decltype(&name##_body<GINGR_UNPAREN targs>)>::type;                        \
  template <GINGR_UNPAREN tdecl>                                                                                        \
  __global__ void __launch_bounds__ bounds name(const __grid_constant__ name##_args<GINGR_UNPAREN targs> a) {           \
    cuda::std::apply(name##_body<GINGR_UNPAREN targs>, a);                                                              \
  }                                                                                                                     \
  template <GINGR_UNPAREN tdecl>                                                                                        \
  __global__ void __launch_bounds__ bounds name##_batched(const name##_args<GINGR_UNPAREN targs>* __restrict__ a) {     \
    cuda::std::apply(name##_body<GINGR_UNPAREN targs>, a[blockIdx.z]);                                                  \
  }                                                                                                                     \
  template <GINGR_UNPAREN tdecl> static __device__ __forceinline__ void name##_body(__VA_ARGS__)

template <typename Args, typename... B>
static inline void gingr_launch(gingr_ctx* ctx, void (*single)(Args), void (*batched)(const Args*), const char* name, dim3 grid,
                                dim3 block, size_t smem, cudaStream_t st, B&&... b) {
  const Args a(static_cast<B&&>(b)...);
  if (ctx->rec) ctx->rec->add((const void*)batched, name, grid, block, smem, &a, sizeof(Args));
  else single<<<grid, block, smem, st>>>(a);
}

#define GINGR_LAUNCH(ctx, name, grid, block, smem, st, ...) \
  gingr_launch<name##_args>(ctx, name, name##_batched, #name, grid, block, smem, st, __VA_ARGS__)
#define GINGR_LAUNCH_T(ctx, name, targs, grid, block, smem, st, ...)                                                      \
  gingr_launch<name##_args<GINGR_UNPAREN targs>>(ctx, name<GINGR_UNPAREN targs>, name##_batched<GINGR_UNPAREN targs>, #name, \
                                                 grid, block, smem, st, __VA_ARGS__)

// device-to-device copy / byte fill on the path of a batched sequence: the driver's copy engines cannot be recorded per chain
namespace {
GINGR_KERNEL_NB(batch_copy_kernel, unsigned long long* __restrict__ dst, const unsigned long long* __restrict__ src, size_t words) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
GINGR_KERNEL_NB(batch_fill_kernel, unsigned char* __restrict__ dst, int value, size_t bytes) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < bytes; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = (unsigned char)value;
}
}  // namespace

// cudaMemcpyAsync(device -> device) of a multiple of 8 bytes (both 8-byte aligned)
static inline cudaError_t gingr_copy_d2d(gingr_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (!ctx->rec || (bytes & 7)) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st);
  const size_t words = bytes / 8;
  const int blocks = (int)(words < 256 * 64 ? (words + 255) / 256 : 64);
  GINGR_LAUNCH(ctx, batch_copy_kernel, blocks > 0 ? blocks : 1, 256, 0, st, (unsigned long long*)dst, (const unsigned long long*)src, words);
  return cudaSuccess;
}

static inline cudaError_t gingr_fill_bytes(gingr_ctx* ctx, void* dst, int value, size_t bytes, cudaStream_t st) {
  if (!ctx->rec) return cudaMemsetAsync(dst, value, bytes, st);
  const int blocks = (int)(bytes < 256 * 64 ? (bytes + 255) / 256 : 64);
  GINGR_LAUNCH(ctx, batch_fill_kernel, blocks > 0 ? blocks : 1, 256, 0, st, (unsigned char*)dst, value, bytes);
  return cudaSuccess;
}
