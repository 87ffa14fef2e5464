// micro-benchmark: Philox4x32-10 with the 32x32->64 products done (a) by IMAD.WIDE.U32 and (b) split into
// lo = IMAD (FMA pipe) and hi = one DFMA.RZ on the FP64 pipe:
//   hi(a*M) = low word of fma_rz( bits(0x43300000:a) = 2^52 + a,  M * 2^-32,  2^52 - M * 2^20 )   (exact: the sum is
//   2^52 + a*M/2^32 with one truncating rounding at ulp 1).
#include <cstdio>
#include <cstdint>

constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;

struct Key { uint32_t rk[20]; };

__device__ __forceinline__ uint4 philox_wide(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const Key &key) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ key.rk[2 * r], n2 = (uint32_t)(p0 >> 32) ^ c3 ^ key.rk[2 * r + 1];
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint32_t mulhi_fp64(uint32_t a, double ms, double c) {
    return (uint32_t)__double2loint(__fma_rz(__hiloint2double(0x43300000, (int)a), ms, c));
}

__device__ __forceinline__ uint4 philox_split(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const Key &key) {
    const double ms0 = (double)M0 * (1.0 / 4294967296.0), ms1 = (double)M1 * (1.0 / 4294967296.0);
    const double k0 = 4503599627370496.0 - (double)M0 * 1048576.0, k1 = 4503599627370496.0 - (double)M1 * 1048576.0;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t h0 = mulhi_fp64(c0, ms0, k0), h1 = mulhi_fp64(c2, ms1, k1);
        const uint32_t l0 = c0 * M0, l1 = c2 * M1;
        const uint32_t n0 = h1 ^ c1 ^ key.rk[2 * r], n2 = h0 ^ c3 ^ key.rk[2 * r + 1];
        c1 = l1; c3 = l0; c0 = n0; c2 = n2;
    }
    return make_uint4(c0, c1, c2, c3);
}

template <int V, int ILP>
__global__ void k(const __grid_constant__ Key key, uint32_t *out, int iters, double fa, double fb) {
    uint32_t acc = 0;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    double d[4] = {fa, fa + 1, fa + 2, fa + 3};
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            const uint4 w = V == 0 ? philox_wide(it * ILP + i, tid, 7u, 9u, key) : philox_split(it * ILP + i, tid, 7u, 9u, key);
            acc ^= w.x ^ w.y ^ w.z ^ w.w;
        }
        if (V >= 2 || true) {                 // the fp64 work of an SSA event pair rides along (about 20 FP64 ops per event)
#pragma unroll
            for (int q = 0; q < 10 * ILP; q++) d[q & 3] = fma(d[q & 3], fb, fa);
        }
    }
    out[tid] = acc ^ (uint32_t)__double2loint(d[0] + d[1] + d[2] + d[3]);
}

template <int V, int ILP>
void run(const char *name, const Key &key, uint32_t *d_out) {
    for (int w : {1, 4, 6}) {
        const int threads = 128 * w, iters = 2048;
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        k<V, ILP><<<148, threads>>>(key, d_out, iters, 1.0000001, 0.9999999);
        cudaEventRecord(a);
        k<V, ILP><<<148, threads>>>(key, d_out, iters, 1.0000001, 0.9999999);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        uint32_t h[4]; cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
        printf("%-34s ILP %d warps/SMSP %d: %.1f cycles per (Philox call + 10 DFMA) per SMSP   out %08x %08x  [%s]\n", name, ILP, w,
               ms * 1e-3 * 1.965e9 / ((double)iters * ILP * w), h[0], h[1], cudaGetErrorString(cudaGetLastError()));
    }
}

int main() {
    Key key;
    for (int r = 0; r < 10; r++) { key.rk[2 * r] = 0x12345678u + r * 0x9E3779B9u; key.rk[2 * r + 1] = 0x9abcdef0u + r * 0xBB67AE85u; }
    uint32_t *d; cudaMalloc(&d, 4 * 148 * 1024);
    run<0, 1>("IMAD.WIDE", key, d);
    run<1, 1>("IMAD lo + DFMA.RZ hi", key, d);
    run<0, 2>("IMAD.WIDE", key, d);
    run<1, 2>("IMAD lo + DFMA.RZ hi", key, d);
    return 0;
}
