// lab_shim.cu -- the C++ symbols the lab's drivers link against (SRM/include/lab.h:9-10),
// implemented over the C-ABI.  Error convention of SRM/include/utils.h:18-26: print
// "GPUassert: ..." and exit; these two functions never return an error.
#include <cstdio>
#include <cstdlib>

#include "../../include/b200sort.h"
#include "../../include/lab.h"

static void lab_fail(int status, const char *file, int line) {
    const char *msg = (status == B200SORT_ERR_CUDA) ? b200sort_last_cuda_error_string()
                                                    : b200sort_status_string(status);
    std::fprintf(stderr, "GPUassert: %s %s %d\n", msg, file, line);
    const int code = (status == B200SORT_ERR_CUDA) ? b200sort_last_cuda_error() : status;
    std::exit(code != 0 ? code : 1);
}

void order_array(int *srcCpu, int length) {
    const int status = (length < 0) ? B200SORT_ERR_INVALID
                                    : b200sort_order_array_host(srcCpu, (size_t)length, B200SORT_ALGO_RADIX);
    if (status != B200SORT_OK) lab_fail(status, __FILE__, __LINE__);
}

void order_with_trust(int *src, int length) {
    const int status = (length < 0) ? B200SORT_ERR_INVALID
                                    : b200sort_order_with_trust_host(src, (size_t)length);
    if (status != B200SORT_OK) lab_fail(status, __FILE__, __LINE__);
}
