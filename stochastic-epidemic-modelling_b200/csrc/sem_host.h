// sem_host.h -- host-side helpers shared by the translation units of libsem_b200.so
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <vector>

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

// ---------------------------------------------------------------------------------------------- candidate-count tables
// Alias tables (Vose) of Poisson(mu) for the 641 grid means in [4, 4096] with six mantissa bits (see "candidate-count
// tables" in sem_common.cuh, which reads them).  Support [mu - 10 sigma - 4, mu + 10 sigma + 12], probabilities by the
// recurrence p(k+1) = p(k) mu / (k+1) from the mode, renormalised; built once per process in fp64.
struct KTabHostEntry { double prob; int32_t alias, pad; };
struct KTabHost { std::vector<KTabHostEntry> e; std::vector<int32_t> meta; };   // meta: 4 ints per table

inline double ktab_grid_mean(int id) {
    const uint64_t bits = (uint64_t)(((uint32_t)(0x40100000u >> 14) + (uint32_t)id) << 14) << 32;
    double mu; memcpy(&mu, &bits, 8);
    return mu;
}

inline void ktab_build_one(double mu, std::vector<KTabHostEntry> &out, int32_t *meta) {
    const double sd = sqrt(mu);
    long long k_lo = (long long)floor(mu - 10.0 * sd) - 4, k_hi = (long long)ceil(mu + 10.0 * sd) + 12;
    if (k_lo < 0) k_lo = 0;
    const int n = (int)(k_hi - k_lo + 1);
    std::vector<double> p(n);
    const long long mode = (long long)floor(mu);
    const double pm = exp((double)mode * log(mu) - mu - lgamma((double)mode + 1.0));
    p[mode - k_lo] = pm;
    for (long long k = mode; k < k_hi; k++) p[k + 1 - k_lo] = p[k - k_lo] * mu / (double)(k + 1);
    for (long long k = mode; k > k_lo; k--) p[k - 1 - k_lo] = p[k - k_lo] * (double)k / mu;
    double sum = 0.0;
    for (int i = 0; i < n; i++) sum += p[i];
    // Vose: scaled probabilities q = n p / sum; small (< 1) columns are topped up by large ones, in index order
    std::vector<double> q(n);
    std::vector<int> small, large;
    for (int i = 0; i < n; i++) { q[i] = p[i] * (double)n / sum; (q[i] < 1.0 ? small : large).push_back(i); }
    const size_t base = out.size();
    out.resize(base + n);
    for (int i = 0; i < n; i++) { out[base + i].prob = 1.0; out[base + i].alias = i; out[base + i].pad = 0; }
    size_t si = 0, li = 0;
    while (si < small.size() && li < large.size()) {
        const int a = small[si++], g = large[li];
        out[base + a].prob = q[a]; out[base + a].alias = g;
        q[g] = (q[g] + q[a]) - 1.0;
        if (q[g] < 1.0) { small.push_back(g); li++; }
    }
    meta[0] = (int32_t)base; meta[1] = n; meta[2] = (int32_t)k_lo; meta[3] = 0;
}

inline const KTabHost &ktab_host() {
    static KTabHost t;
    static std::once_flag once;
    std::call_once(once, [] {
        const int count = (int)((0x40B00000u >> 14) - (0x40100000u >> 14)) + 1;
        t.meta.resize(4 * (size_t)count);
        for (int id = 0; id < count; id++) ktab_build_one(ktab_grid_mean(id), t.e, &t.meta[4 * (size_t)id]);
    });
    return t;
}

// device copy of the tables for the CURRENT device (made on first use, kept for the life of the process)
inline int ktab_device(const void **entries, const void **meta) {
    static const void *d_e[64] = {nullptr}, *d_m[64] = {nullptr};
    static std::mutex mu;
    int dev = 0;
    SEM_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device index out of range"); return SEM_ERR_INVALID; }
    std::lock_guard<std::mutex> lock(mu);
    if (!d_e[dev]) {
        const KTabHost &h = ktab_host();
        void *e = nullptr, *m = nullptr;
        SEM_CUDA(cudaMalloc(&e, h.e.size() * sizeof(KTabHostEntry)));
        SEM_CUDA(cudaMalloc(&m, h.meta.size() * sizeof(int32_t)));
        SEM_CUDA(cudaMemcpy(e, h.e.data(), h.e.size() * sizeof(KTabHostEntry), cudaMemcpyHostToDevice));
        SEM_CUDA(cudaMemcpy(m, h.meta.data(), h.meta.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        d_e[dev] = e; d_m[dev] = m;
    }
    *entries = d_e[dev]; *meta = d_m[dev];
    return SEM_OK;
}

}  // namespace sem
