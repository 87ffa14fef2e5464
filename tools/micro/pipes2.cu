// micro-benchmark 2: which pipe does the 32x32->64 multiply of Philox use on sm_100a, and does it overlap with FP64?
#include <cstdio>
#include <cstdint>

constexpr int ITERS = 4096;

__device__ __forceinline__ uint32_t mulhi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t mullo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint64_t mulwide(uint32_t a, uint32_t b) { uint64_t r; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r; }

template <int OP>
__global__ void k(double *out, uint32_t *outi, double a, double b, uint32_t ua, uint32_t ub) {
    double d[8];
    uint32_t u[8], v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { d[i] = a + i + threadIdx.x; u[i] = ua + i * 77 + threadIdx.x; v[i] = u[i] * 3; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0 || OP == 3) u[i] = mulhi(u[i], 0xD2511F53u) ^ ub;
            if (OP == 1 || OP == 4) u[i] = mullo(u[i], 0xD2511F53u) ^ ub;
            if (OP == 2 || OP == 5) { const uint64_t p = mulwide(u[i], 0xD2511F53u); u[i] = (uint32_t)(p >> 32) ^ ub; v[i] ^= (uint32_t)p; }
            if (OP == 6 || OP == 7) { const uint32_t h = mulhi(u[i], 0xD2511F53u), l = mullo(u[i], 0xD2511F53u); u[i] = h ^ ub; v[i] ^= l; }
            if (OP == 3 || OP == 4 || OP == 5 || OP == 7 || OP == 8) d[i] = fma(d[i], a, b);
            if (OP == 9) { u[i] = (u[i] << 3) ^ (v[i] >> 5) ^ ub; v[i] = v[i] + u[i]; }       // plain ALU
            if (OP == 10) { u[i] = (u[i] << 3) ^ (v[i] >> 5) ^ ub; v[i] = v[i] + u[i]; d[i] = fma(d[i], a, b); }
        }
    }
    double s = 0; uint32_t su = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { s += d[i]; su += u[i] + v[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    outi[blockIdx.x * blockDim.x + threadIdx.x] = su;
}

template <int OP>
void run(const char *name, double *d, uint32_t *di) {
    for (int w : {1, 4, 8}) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        const int threads = 128 * w;
        k<OP><<<148, threads>>>(d, di, 1.0000001, 1e-9, 0x9E3779B9u, 12345u);
        cudaEventRecord(a);
        k<OP><<<148, threads>>>(d, di, 1.0000001, 1e-9, 0x9E3779B9u, 12345u);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double cyc = ms * 1e-3 * 1.965e9;
        printf("%-34s warps/SMSP %d: %.2f cycles per group per SMSP  [%s]\n", name, w, cyc / ((double)ITERS * 8 * w), cudaGetErrorString(cudaGetLastError()));
    }
}

int main() {
    double *d; uint32_t *di;
    cudaMalloc(&d, 8 * 148 * 1024); cudaMalloc(&di, 4 * 148 * 1024);
    run<0>("mul.hi + xor", d, di);
    run<1>("mul.lo + xor", d, di);
    run<2>("mul.wide + 2 xor", d, di);
    run<6>("mul.hi + mul.lo + 2 xor", d, di);
    run<8>("DFMA", d, di);
    run<3>("DFMA + mul.hi + xor", d, di);
    run<4>("DFMA + mul.lo + xor", d, di);
    run<5>("DFMA + mul.wide + 2 xor", d, di);
    run<7>("DFMA + mul.hi + mul.lo + 2 xor", d, di);
    run<9>("4 ALU", d, di);
    run<10>("DFMA + 4 ALU", d, di);
    return 0;
}
