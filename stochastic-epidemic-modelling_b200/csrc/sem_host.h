// sem_host.h -- host-side helpers shared by the translation units of libsem_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

#include "../../include/sem_b200.h"

namespace sem {

inline char *err_buf() {
    static thread_local char buf[512] = "";
    return buf;
}
inline void set_error(const char *fmt, const char *a = "", const char *b = "") { snprintf(err_buf(), 512, fmt, a, b); }

#define SEM_CUDA(call)                                                                                       \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) { sem::set_error("%s: %s", #call, cudaGetErrorString(e_)); return SEM_ERR_CUDA; } \
    } while (0)

inline int model_cols(int model, int G) { return model == SEM_MODEL_SIR ? 3 : model == SEM_MODEL_SEIR ? 4 : 3 * G; }
inline int model_ntheta(int model, int G) { return model == SEM_MODEL_SIR ? 2 : model == SEM_MODEL_SEIR ? 3 : G * G + 1; }

inline int sm_count() {                                      // of the CURRENT device (cached per device)
    static int n_sm[64] = {0};
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!n_sm[dev]) n_sm[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
    return n_sm[dev];
}

}  // namespace sem
