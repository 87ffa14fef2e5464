// micro-benchmark of the uniformized candidate loop (SIR): cycles per group of four candidates per scheduler as a
// function of the resident warps per scheduler, for variants of the candidate test.  Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --fmad=false -o tools/micro/cand_loop tools/micro/cand_loop.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../stochastic-epidemic-modelling_b200/csrc/sem_common.cuh"

using namespace sem;

// V0: fp64 compares of u*B against the propensities (the round-1 loop)
// V1: fixed-point thresholds, integer compares (production, unif32_candidate)
// V2: Philox only (the words are xor-ed into a checksum)
// V3: V1's candidates fed by a counter hash instead of Philox (candidate cost alone)
// V4: V1 with Philox4x32 cut to 7 rounds (what the generator costs)
// V5: V1 with the next group's Philox call inside the current group's straight-line block (software pipeline)
template <int ROUNDS>
__device__ __forceinline__ uint4 philox_r(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKey &key) {
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ key.rk[2 * r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ key.rk[2 * r + 1];
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    }
    return make_uint4(c0, c1, c2, c3);
}

template <int V>
__global__ void __launch_bounds__(768) k_cand(const double *theta, int n, const __grid_constant__ PhiloxKey key, int32_t *Xout, uint32_t groups) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double x[3] = {8000.0, 1000.0, 1000.0};
    SirModel m;
    m.setup(theta, x);
    double r[2];
    double a0 = ssa_total<SirModel, SEM_ARITH_FAST>(m, x, r);
    const double B = 1.6 * a0;
    SirModel::Scaled sc;
    m.scale(sc, __ddiv_rn(4294967296.0, B));
    double T[2];
    m.thresholds(sc, x, T);
    uint32_t cand = 0, chk = 0;
    const uint32_t last = 4u * groups;
    bool stop = false;
    if constexpr (V == 5) {                                  // V1 with the next group's Philox call INSIDE the straight-line block of the current group
        uint4 w = philox4x32_10(0u, (uint32_t)j, 5u, stream_word(DOM_SSA, 0), key);
        while (cand < last && !stop) {
            const uint32_t k = cand >> 2;
            if ((cand & 3u) == 0u && last - cand >= 4u) {
                const uint4 wn = philox4x32_10(k + 1u, (uint32_t)j, 5u, stream_word(DOM_SSA, 0), key);
                const uint32_t words[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const bool live = !stop;
                    cand += live ? 1u : 0u;
                    unif32_candidate<SirModel, false>(m, sc, x, T, words[q], live, stop);
                }
                w = wn;
            } else {
                const uint32_t base = cand & ~3u;
#pragma unroll 1
                for (uint32_t q = 0; q < 4u; q++) {
                    const uint32_t c = base + q;
                    const bool live = !stop && c >= cand && c < last;
                    const uint32_t wq = q == 0u ? w.x : q == 1u ? w.y : q == 2u ? w.z : w.w;
                    unif32_candidate<SirModel, false>(m, sc, x, T, wq, live, stop);
                    cand += live ? 1u : 0u;
                }
                w = philox4x32_10((cand >> 2), (uint32_t)j, 5u, stream_word(DOM_SSA, 0), key);
            }
        }
    }
    if constexpr (V == 1) {                                  // the production loop
        PairSource<false> loc; loc.init(key, (uint32_t)j, 5u, stream_word(DOM_SSA, 0));
        unif32_serve<SirModel, false>(m, sc, x, T, loc, cand, last, stop);
    }
    while (V != 1 && V != 5 && cand < last && !stop) {
        uint4 w;
        if constexpr (V == 3) {
            uint32_t h = (cand >> 2) * 0x9E3779B9u + (uint32_t)j;
            w = make_uint4(h, h * 0x85EBCA6Bu, h ^ 0xC2B2AE35u, h * 0x27D4EB2Fu);
        } else if constexpr (V == 4) w = philox_r<7>(cand >> 2, (uint32_t)j, 5u, stream_word(DOM_SSA, 0), key);
        else w = philox4x32_10(cand >> 2, (uint32_t)j, 5u, stream_word(DOM_SSA, 0), key);
        const uint32_t words[4] = {w.x, w.y, w.z, w.w};
        if constexpr (V == 2) { chk ^= w.x ^ w.y ^ w.z ^ w.w; cand += 4; }
        else if constexpr (V == 0) {
            const uint32_t base = cand & ~3u;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t c = base + q;
                const bool live = !stop && c >= cand && c < last;
                const double v = __fma_rn(word_to_d12(words[q]), B, -B);
                const bool hit = live && v < a0;
                const int jj = (r[0] <= v) ? 1 : 0;
                if (hit) m.template apply<false>(x, jj);
                a0 = ssa_total<SirModel, SEM_ARITH_FAST>(m, x, r);
                stop = stop || (hit && !(a0 > 0 && a0 <= B));
                cand += live ? 1u : 0u;
            }
        } else {
            if ((cand & 3u) == 0u && last - cand >= 4u) {
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const bool live = !stop;
                    cand += live ? 1u : 0u;
                    unif32_candidate<SirModel, false>(m, sc, x, T, words[q], live, stop);
                }
            } else {
                const uint32_t base = cand & ~3u;
#pragma unroll 1
                for (uint32_t q = 0; q < 4u; q++) {
                    const uint32_t c = base + q;
                    const bool live = !stop && c >= cand && c < last;
                    const uint32_t wq = q == 0u ? w.x : q == 1u ? w.y : q == 2u ? w.z : w.w;
                    unif32_candidate<SirModel, false>(m, sc, x, T, wq, live, stop);
                    cand += live ? 1u : 0u;
                }
            }
        }
    }
    Xout[j] = (int32_t)x[0] + 3 * (int32_t)x[1] + (int32_t)chk + (int32_t)cand;
}

template <int V>
static void run(int warps_per_smsp, const double *d_theta, int32_t *d_out, uint32_t groups) {
    const int threads = 32 * 4 * warps_per_smsp, blocks = 148, n = threads * blocks;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int it = 0; it < 4; it++) {
        cudaEventRecord(a);
        k_cand<V><<<blocks, threads>>>(d_theta, n, make_philox_key(1234ull), d_out, groups);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (it && ms < best) best = ms;
    }
    std::vector<int32_t> h(n);
    cudaMemcpy(h.data(), d_out, sizeof(int32_t) * n, cudaMemcpyDeviceToHost);
    long long cs = 0;
    for (int i = 0; i < n; i++) cs += h[i];
    const double cyc = best * 1e-3 * 1.965e9;
    printf("variant %d  warps/SMSP %d  %.3f ms  cycles per group: %.0f per warp, %.1f per SMSP  checksum %lld  %s\n", V, warps_per_smsp, best,
           cyc / groups, cyc / groups / warps_per_smsp, cs, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char **argv) {
    const int maxn = 148 * 768;
    double theta[2] = {0.4, 0.2};
    double *d_theta; int32_t *d_out;
    cudaMalloc(&d_theta, 16); cudaMalloc(&d_out, sizeof(int32_t) * maxn);
    cudaMemcpy(d_theta, theta, 16, cudaMemcpyHostToDevice);
    const uint32_t groups = argc > 1 ? (uint32_t)atoi(argv[1]) : 250u;
    for (int w = 1; w <= 6; w++) {
        if (w == 3 || w == 4) continue;
        run<0>(w, d_theta, d_out, groups);
        run<1>(w, d_theta, d_out, groups);
        run<2>(w, d_theta, d_out, groups);
        run<3>(w, d_theta, d_out, groups);
        run<4>(w, d_theta, d_out, groups);
        run<5>(w, d_theta, d_out, groups);
    }
    return 0;
}
