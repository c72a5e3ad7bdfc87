// nccl_dl.cuh -- run-time bound NCCL collectives used inside gingr_update (nccl_dl.cu).
#pragma once
#include "common.cuh"

namespace gingr {
int32_t comm_unique_id(char id[128]);
int32_t comm_init(gingr_ctx* ctx, int nranks, int rank, const char id[128]);
void comm_destroy(gingr_ctx* ctx);
// in-place sum all-reduce of `count` doubles on ctx->stream (no-op when nranks == 1)
int32_t comm_allreduce_sum(gingr_ctx* ctx, double* d_buf, size_t count);
// all-gather of equal-sized blocks (count_per_rank doubles each) on ctx->stream
int32_t comm_allgather(gingr_ctx* ctx, const double* d_send, double* d_recv, size_t count_per_rank);
}  // namespace gingr
