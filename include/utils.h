/*
 * utils.h -- the two conveniences the lab's drivers expect next to lab.h
 * (reference: SRM/include/utils.h:18-36): CUDA_CHK for abort-on-error and MS(call, name) which
 * declares `double name` holding the wall-clock milliseconds `call` took.
 */
#ifndef B200SORT_UTILS_H
#define B200SORT_UTILS_H

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

static inline void b200sort_cuda_chk(cudaError_t code, const char *file, int line)
{
    if (code == cudaSuccess) return;
    fprintf(stderr, "GPUassert: %s %s %d\n", cudaGetErrorString(code), file, line);
    exit((int)code);
}
#define CUDA_CHK(ans) b200sort_cuda_chk((ans), __FILE__, __LINE__);

static inline double b200sort_now_ms(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return 1000.0 * (double)t.tv_sec + (double)t.tv_nsec / 1.0e6;
}
#ifndef MS
#define MS(f, elap)                                  \
    double elap = 0;                                 \
    {                                                \
        const double b200sort_t0 = b200sort_now_ms(); \
        f;                                           \
        elap = b200sort_now_ms() - b200sort_t0;      \
    }
#endif

#endif /* B200SORT_UTILS_H */
