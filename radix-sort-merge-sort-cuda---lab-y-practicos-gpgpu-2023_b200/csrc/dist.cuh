// dist.cuh -- internal interface of the one-box multi-GPU path (dist.cu).
#pragma once
#include "common.cuh"

namespace b200sort {

// Device-resident plan of one distributed sort (written by dist_plan_kernel; B200SORT_DIST_PLAN_BYTES).
struct DistPlanDev {
    unsigned long long recv_count[B200SORT_DIST_MAX_WORLD];   // keys every rank owns after the exchange
    unsigned long long send_count[B200SORT_DIST_MAX_WORLD];   // keys THIS rank sends to every rank
    unsigned long long dst_offset[B200SORT_DIST_MAX_WORLD];   // where this rank's block starts in every receive buffer
    unsigned int m;                                           // recv_count[this rank], for the local sort
    unsigned int error;                                       // 1: some rank would receive more than its buffer holds
    unsigned int pad[2];
    unsigned int top_hist[256];                               // top-byte histogram of the keys this rank will own
};
static_assert(sizeof(DistPlanDev) == B200SORT_DIST_PLAN_BYTES, "include/b200sort.h states the size of the plan record");

size_t dist_workspace_bytes(size_t n, int bits);
int dist_histogram(const int32_t *d_keys, size_t n, int bits, unsigned long long *d_hist, cudaStream_t s);
int dist_plan(const unsigned long long *all_hist, int world, int rank, int bits, int *bin_owner,
              unsigned long long *recv_count, unsigned long long *send_count,
              unsigned long long *dst_offset);
int dist_partition(const int32_t *d_keys, size_t n, int bits, int world, int32_t *const *h_dst_base,
                   const int *d_bin_owner, const unsigned long long *h_dst_offset, void *d_ws,
                   size_t ws_bytes, cudaStream_t s);

int dist_plan_device(const unsigned long long *d_all_hist, int world, int rank, int bits, unsigned long long cap,
                     int *d_bin_owner, void *d_plan, void *d_ws, size_t ws_bytes, cudaStream_t s);
int dist_partition_planned(const int32_t *d_keys, size_t n, int bits, int world, int32_t *const *h_dst_base,
                           const int *d_bin_owner, const void *d_plan, unsigned int *d_src_hist, void *d_ws, size_t ws_bytes,
                           cudaStream_t s);

}  // namespace b200sort
