// micro-benchmark: cost of cooperative_groups grid.sync() for the filter's launch shape
#include <cooperative_groups.h>
#include <cstdio>
namespace cg = cooperative_groups;
__global__ void k(int iters, int *out) {
    cg::grid_group g = cg::this_grid();
    int acc = 0;
    for (int i = 0; i < iters; i++) { acc += i; g.sync(); }
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = acc;
}
__global__ void empty(int *out) { if (threadIdx.x == 0 && blockIdx.x == 0) *out = 1; }
int main() {
    int *d; cudaMalloc(&d, 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int threads : {32, 704}) for (int blocks : {32, 148}) {
        int iters = 2000; void *args[] = {&iters, &d};
        cudaLaunchCooperativeKernel((void *)k, dim3(blocks), dim3(threads), args, 0, 0); cudaDeviceSynchronize();
        cudaEventRecord(a); cudaLaunchCooperativeKernel((void *)k, dim3(blocks), dim3(threads), args, 0, 0); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("grid.sync  blocks=%d threads=%d : %.2f us per sync (%s)\n", blocks, threads, 1e3 * ms / iters, cudaGetErrorString(cudaGetLastError()));
    }
    cudaEventRecord(a); for (int i = 0; i < 2000; i++) empty<<<148, 704>>>(d); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); printf("back-to-back empty launches 148x704: %.2f us per launch\n", 1e3 * ms / 2000);
    return 0;
}
