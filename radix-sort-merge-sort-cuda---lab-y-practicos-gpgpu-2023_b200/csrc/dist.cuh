// dist.cuh -- internal interface of the one-box multi-GPU path (dist.cu).
#pragma once
#include "common.cuh"

namespace b200sort {

size_t dist_workspace_bytes(size_t n, int bits);
int dist_histogram(const int32_t *d_keys, size_t n, int bits, unsigned long long *d_hist, cudaStream_t s);
int dist_plan(const unsigned long long *all_hist, int world, int rank, int bits, int *bin_owner,
              unsigned long long *recv_count, unsigned long long *send_count,
              unsigned long long *dst_offset);
int dist_partition(const int32_t *d_keys, size_t n, int bits, int world, int32_t *const *h_dst_base,
                   const int *d_bin_owner, const unsigned long long *h_dst_offset, void *d_ws,
                   size_t ws_bytes, cudaStream_t s);

}  // namespace b200sort
