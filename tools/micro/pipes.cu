// micro-benchmark: per-SMSP issue cost (cycles per warp instruction) of the instruction classes in the SSA loop
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/micro/pipes tools/micro/pipes.cu
#include <cstdio>
#include <cstdint>

constexpr int ITERS = 4096;

template <int OP>
__global__ void k(double *out, uint32_t *outi, double a, double b, uint32_t ua, uint32_t ub) {
    double d[8];
    uint32_t u[8];
    uint64_t q[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { d[i] = a + i + threadIdx.x; u[i] = ua + i * 77 + threadIdx.x; q[i] = u[i]; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) d[i] = fma(d[i], a, b);                                   // DFMA
            if (OP == 1) d[i] = d[i] * a;                                          // DMUL
            if (OP == 2) d[i] = d[i] + b;                                          // DADD
            if (OP == 3) { q[i] = (uint64_t)(uint32_t)q[i] * 0xD2511F53u + (q[i] >> 32); }  // IMAD.WIDE (+ add hi)
            if (OP == 4) u[i] = (u[i] ^ ua ^ (u[(i + 1) & 7])) ;                   // LOP3
            if (OP == 5) u[i] = u[i] * ua + ub;                                    // IMAD
            if (OP == 6) { d[i] = (d[i] > b) ? d[i] - 1.0 : d[i] + a; }            // DSETP + DADD + select-ish
            if (OP == 7) { d[i] = fma(d[i], a, b); u[i] = (u[i] ^ ua ^ (u[(i + 1) & 7])); }   // DFMA + LOP3 co-issue?
            if (OP == 8) { d[i] = fma(d[i], a, b); q[i] = (uint64_t)(uint32_t)q[i] * 0xD2511F53u + (q[i] >> 32); }  // DFMA + IMAD.WIDE
            if (OP == 9) { q[i] = (uint64_t)(uint32_t)q[i] * 0xD2511F53u + (q[i] >> 32); u[i] = (u[i] ^ ua ^ (u[(i + 1) & 7])); }
        }
    }
    double s = 0; uint32_t su = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { s += d[i]; su += u[i] + (uint32_t)q[i] + (uint32_t)(q[i] >> 32); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    outi[blockIdx.x * blockDim.x + threadIdx.x] = su;
}

template <int OP>
void run(const char *name, int n_inst_per_iter, double *d, uint32_t *di) {
    for (int w : {1, 2, 4, 8}) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        const int threads = 128 * w;
        k<OP><<<148, threads>>>(d, di, 1.0000001, 1e-9, 0x9E3779B9u, 12345u);
        cudaEventRecord(a);
        k<OP><<<148, threads>>>(d, di, 1.0000001, 1e-9, 0x9E3779B9u, 12345u);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double cyc = ms * 1e-3 * 1.965e9;
        printf("%-28s warps/SMSP %d: %.2f cycles per warp-instruction-group (%d instr) per SMSP  [%s]\n", name, w,
               cyc / ((double)ITERS * 8 * w), n_inst_per_iter, cudaGetErrorString(cudaGetLastError()));
    }
}

int main() {
    double *d; uint32_t *di;
    cudaMalloc(&d, 8 * 148 * 1024); cudaMalloc(&di, 4 * 148 * 1024);
    run<0>("DFMA", 1, d, di);
    run<1>("DMUL", 1, d, di);
    run<2>("DADD", 1, d, di);
    run<3>("IMAD.WIDE", 1, d, di);
    run<4>("LOP3", 1, d, di);
    run<5>("IMAD", 1, d, di);
    run<6>("DSETP+DADD+DADD+SEL", 4, d, di);
    run<7>("DFMA + LOP3", 2, d, di);
    run<8>("DFMA + IMAD.WIDE", 2, d, di);
    run<9>("IMAD.WIDE + LOP3", 2, d, di);
    return 0;
}
