// micro-benchmark of the SSA event loop (SIR, FAST arithmetic): cycles per event per warp as a function of the number
// of resident warps per scheduler, for loop variants.  Not part of the product; used to choose the production loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --fmad=false -o tools/micro/ssa_loop tools/micro/ssa_loop.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../stochastic-epidemic-modelling_b200/csrc/sem_common.cuh"

using namespace sem;

#ifndef VARIANTS_EXTRA
#define VARIANTS_EXTRA
#endif

// variant 0: one event per iteration (ssa_run_fast); 1,2: speculative blocks of 2,4 events, 52-bit uniforms;
// 3,4,5: blocks of 2,4,6 events with 32-bit uniforms (one Philox call per two events)
template <int V>
__global__ void __launch_bounds__(768) k_loop(const double *theta, const int32_t *X0, int n, double dt, const PhiloxKey key, int32_t *Xout,
                                              unsigned long long *events, int reps) {
    __shared__ double2 s_tab[kLogTabSize];
    load_logtab(s_tab);
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    unsigned long long my = 0;
    double xo[3] = {0, 0, 0};
    for (int rep = 0; rep < reps; rep++) {
        double x[3] = {(double)X0[j], (double)X0[n + j], (double)X0[2 * n + j]};
        SirModel m;
        m.setup(theta, x);
        PairSource<false> src;
        src.init(key, (uint32_t)j, (uint32_t)rep, stream_word(DOM_SSA, 0));
        long long pairs;
        if constexpr (V == 0) pairs = ssa_run_fast<SirModel, false>(m, x, dt, src, s_tab, NoRec());
        else if constexpr (V == 1) pairs = ssa_run_spec<SirModel, 2, false, false>(m, x, dt, src, s_tab, NoRec());
        else if constexpr (V == 2) pairs = ssa_run_spec<SirModel, 4, false, false>(m, x, dt, src, s_tab, NoRec());
        else if constexpr (V == 3) pairs = ssa_run_spec<SirModel, 2, true, false>(m, x, dt, src, s_tab, NoRec());
        else if constexpr (V == 4) pairs = ssa_run_spec<SirModel, 4, true, false>(m, x, dt, src, s_tab, NoRec());
        else if constexpr (V == 5) pairs = ssa_run_spec<SirModel, 6, true, false>(m, x, dt, src, s_tab, NoRec());
        else {
            PairSource<false> aux; aux.init(key, (uint32_t)j, (uint32_t)rep, stream_word(DOM_AUX, 0));
            if constexpr (V == 6) pairs = ssa_run_unif<SirModel, false>(m, x, dt, src, aux, s_tab);
            else pairs = ssa_run_unif32<SirModel, false>(m, x, dt, src, aux, s_tab);
        }
        my += (unsigned long long)pairs;
        xo[0] += x[0]; xo[1] += x[1]; xo[2] += x[2];
    }
    Xout[j] = (int32_t)xo[0]; Xout[n + j] = (int32_t)xo[1]; Xout[2 * n + j] = (int32_t)xo[2];
#pragma unroll
    for (int d = 16; d; d >>= 1) my += __shfl_xor_sync(0xffffffffu, my, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(events, my);
}

template <int V>
static void run(int warps_per_smsp, const double *d_theta, const int32_t *d_X0, int32_t *d_out, unsigned long long *d_ev, int reps,
                long long *checksum) {
    const int threads = 32 * 4 * warps_per_smsp, blocks = 148, n = threads * blocks;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    unsigned long long ev = 0;
    for (int it = 0; it < 4; it++) {
        cudaMemset(d_ev, 0, 8);
        cudaEventRecord(a);
        k_loop<V><<<blocks, threads>>>(d_theta, d_X0, n, 1.0, make_philox_key(1234ull), d_out, d_ev, reps);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (it && ms < best) best = ms;
        cudaMemcpy(&ev, d_ev, 8, cudaMemcpyDeviceToHost);
    }
    std::vector<int32_t> h(3 * n);
    cudaMemcpy(h.data(), d_out, sizeof(int32_t) * 3 * n, cudaMemcpyDeviceToHost);
    long long cs = 0;
    for (int i = 0; i < n; i++) cs += (long long)h[i] * 3 + h[n + i];
    const double ev_per_lane = (double)ev / n;                   // mean events per lane (a warp iterates to its slowest lane)
    const double cyc = best * 1e-3 * 1.965e9;
    printf("variant %d  warps/SMSP %d  %.3f ms  events %llu  %.1f Gev/s  cycles per lane-event: %.0f per warp, %.1f per SMSP  checksum %lld  %s\n",
           V, warps_per_smsp, best, ev, ev / (best * 1e-3) / 1e9, cyc / ev_per_lane, cyc / ev_per_lane / warps_per_smsp, cs,
           cudaGetErrorString(cudaGetLastError()));
    if (checksum) *checksum = cs;
}

int main(int argc, char **argv) {
    const int maxn = 148 * 768;
    std::vector<int32_t> X0(3 * maxn);
    for (int i = 0; i < maxn; i++) { X0[i] = 8000; X0[maxn + i] = 1000; X0[2 * maxn + i] = 1000; }
    // kernels index X0 with stride n (not maxn): upload per-n layouts lazily -> simply make all rows constant
    double theta[2] = {0.4, 0.2};
    double *d_theta; int32_t *d_X0, *d_out; unsigned long long *d_ev;
    cudaMalloc(&d_theta, 16); cudaMalloc(&d_X0, sizeof(int32_t) * 3 * maxn); cudaMalloc(&d_out, sizeof(int32_t) * 3 * maxn); cudaMalloc(&d_ev, 8);
    cudaMemcpy(d_theta, theta, 16, cudaMemcpyHostToDevice);
    const int reps = argc > 1 ? atoi(argv[1]) : 8;
    for (int w = 1; w <= 6; w++) {
        // constant rows per layout
        const int n = 32 * 4 * w * 148;
        for (int i = 0; i < n; i++) { X0[i] = 8000; X0[n + i] = 1000; X0[2 * n + i] = 1000; }
        cudaMemcpy(d_X0, X0.data(), sizeof(int32_t) * 3 * n, cudaMemcpyHostToDevice);
        long long c0 = 0, c = 0;
        run<0>(w, d_theta, d_X0, d_out, d_ev, reps, &c0);
        run<1>(w, d_theta, d_X0, d_out, d_ev, reps, &c); if (c != c0) printf("   MISMATCH v1\n");
        run<2>(w, d_theta, d_X0, d_out, d_ev, reps, &c); if (c != c0) printf("   MISMATCH v2\n");
        run<3>(w, d_theta, d_X0, d_out, d_ev, reps, &c0);
        run<4>(w, d_theta, d_X0, d_out, d_ev, reps, &c); if (c != c0) printf("   MISMATCH v4\n");
        run<5>(w, d_theta, d_X0, d_out, d_ev, reps, &c); if (c != c0) printf("   MISMATCH v5\n");
        run<6>(w, d_theta, d_X0, d_out, d_ev, reps, &c);      // uniformized, 52-bit candidates (2 per call)
        run<7>(w, d_theta, d_X0, d_out, d_ev, reps, &c);      // uniformized, 32-bit candidates (4 per call), drift-anticipating bound

    }
    return 0;
}
