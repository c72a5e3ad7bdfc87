// nccl_dl.cu -- NCCL bound at run time with dlopen, so libgingr_cuda.so has no link-time dependency
// on a particular libnccl (inside a PyTorch process the bundled libnccl.so.2 is reused; a JVM host
// picks up the system one).  Only the few entry points the library needs are resolved.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "nccl_dl.cuh"

namespace gingr {

// Minimal NCCL ABI (stable since 2.x): ncclUniqueId is 128 bytes, ncclComm_t is an opaque pointer,
// ncclDataType_t ncclFloat64 = 8, ncclRedOp_t ncclSum = 0.
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef int (*fn_ncclGetUniqueId)(ncclUniqueId_t*);
typedef int (*fn_ncclCommInitRank)(void**, int, ncclUniqueId_t, int);
typedef int (*fn_ncclCommDestroy)(void*);
typedef int (*fn_ncclAllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_ncclAllGather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*fn_ncclGetErrorString)(int);

struct NcclFns {
  void* handle = nullptr;
  fn_ncclGetUniqueId GetUniqueId = nullptr;
  fn_ncclCommInitRank CommInitRank = nullptr;
  fn_ncclCommDestroy CommDestroy = nullptr;
  fn_ncclAllReduce AllReduce = nullptr;
  fn_ncclAllGather AllGather = nullptr;
  fn_ncclGetErrorString GetErrorString = nullptr;
};

static NcclFns g_nccl;

static bool nccl_load(std::string* err) {
  if (g_nccl.handle) return true;
  const char* env = getenv("GINGR_NCCL_LIB");
  const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n) continue;
    // RTLD_NOLOAD first: reuse a copy that is already mapped (e.g. torch's bundled NCCL)
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);
    if (!h) h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    *err = std::string("cannot dlopen libnccl.so.2 (set GINGR_NCCL_LIB): ") + (dlerror() ? dlerror() : "");
    return false;
  }
  g_nccl.GetUniqueId = (fn_ncclGetUniqueId)dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (fn_ncclCommInitRank)dlsym(h, "ncclCommInitRank");
  g_nccl.CommDestroy = (fn_ncclCommDestroy)dlsym(h, "ncclCommDestroy");
  g_nccl.AllReduce = (fn_ncclAllReduce)dlsym(h, "ncclAllReduce");
  g_nccl.AllGather = (fn_ncclAllGather)dlsym(h, "ncclAllGather");
  g_nccl.GetErrorString = (fn_ncclGetErrorString)dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce ||
      !g_nccl.AllGather) {
    *err = "libnccl is missing required symbols";
    return false;
  }
  g_nccl.handle = h;
  return true;
}

static int32_t nccl_fail(gingr_ctx* ctx, const char* what, int rc) {
  std::string m = std::string(what) + " failed: " +
                  (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "nccl error");
  return gingr_fail(ctx, GINGR_ERR_NCCL, m.c_str());
}

int32_t comm_unique_id(char id[128]) {
  std::string err;
  if (!nccl_load(&err)) return gingr_fail(nullptr, GINGR_ERR_NCCL, err.c_str());
  ncclUniqueId_t u;
  int rc = g_nccl.GetUniqueId(&u);
  if (rc != 0) return nccl_fail(nullptr, "ncclGetUniqueId", rc);
  memcpy(id, u.internal, 128);
  return GINGR_OK;
}

int32_t comm_init(gingr_ctx* ctx, int nranks, int rank, const char id[128]) {
  if (nranks < 1 || rank < 0 || rank >= nranks) return gingr_fail(ctx, GINGR_ERR_ARG, "gingr_comm_init: bad rank");
  if (ctx->nccl_comm) return gingr_fail(ctx, GINGR_ERR_ARG, "gingr_comm_init: communicator already initialised");
  ctx->nranks = nranks;
  ctx->rank = rank;
  if (nranks == 1) return GINGR_OK;
  std::string err;
  if (!nccl_load(&err)) return gingr_fail(ctx, GINGR_ERR_NCCL, err.c_str());
  GINGR_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId_t u;
  memcpy(u.internal, id, 128);
  int rc = g_nccl.CommInitRank(&ctx->nccl_comm, nranks, u, rank);
  if (rc != 0) {
    ctx->nccl_comm = nullptr;
    ctx->nranks = 1;
    ctx->rank = 0;
    return nccl_fail(ctx, "ncclCommInitRank", rc);
  }
  return GINGR_OK;
}

void comm_destroy(gingr_ctx* ctx) {
  if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
}

int32_t comm_allreduce_sum(gingr_ctx* ctx, double* d_buf, size_t count) {
  if (ctx->nranks <= 1) return GINGR_OK;
  int rc = g_nccl.AllReduce(d_buf, d_buf, count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, ctx->nccl_comm, ctx->stream);
  if (rc != 0) return nccl_fail(ctx, "ncclAllReduce", rc);
  return GINGR_OK;
}

int32_t comm_allgather(gingr_ctx* ctx, const double* d_send, double* d_recv, size_t count_per_rank) {
  if (ctx->nranks <= 1) return GINGR_OK;
  int rc = g_nccl.AllGather(d_send, d_recv, count_per_rank, /*ncclFloat64*/ 8, ctx->nccl_comm, ctx->stream);
  if (rc != 0) return nccl_fail(ctx, "ncclAllGather", rc);
  return GINGR_OK;
}

}  // namespace gingr

extern "C" {
int32_t gingr_comm_unique_id(char id[128]) {
  if (!id) return GINGR_ERR_ARG;
  return gingr::comm_unique_id(id);
}
int32_t gingr_comm_init(gingr_ctx* ctx, int32_t nranks, int32_t rank, const char id[128]) {
  if (!ctx || !id) return GINGR_ERR_ARG;
  return gingr::comm_init(ctx, nranks, rank, id);
}
}
