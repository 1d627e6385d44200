// Event-timed cost of an (almost) empty persistent kernel: ordinary launch vs cooperative launch,
// with and without one grid barrier.  nvcc -gencode arch=compute_100a,code=sm_100a -o launch_overhead launch_overhead.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
__global__ void __launch_bounds__(256, 4) k_empty(int *p) { if (p && threadIdx.x == 9999) *p = 1; }
__global__ void __launch_bounds__(256, 4) k_sync(int *p) { cg::this_grid().sync(); if (p && threadIdx.x == 9999) *p = 1; }
__device__ unsigned int g_bar[2];
__global__ void __launch_bounds__(256, 4) k_own(int *p, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&g_bar[0], 1u);
        while (*(volatile unsigned int *)&g_bar[0] < target) ;
        __threadfence();
    }
    __syncthreads();
    if (p && threadIdx.x == 9999) *p = 1;
}
int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int *p = nullptr;
    void *args[] = {&p};
    auto timeit = [&](const char *name, auto launch) {
        float best = 1e9f, sum = 0;
        for (int i = 0; i < 60; ++i) {
            cudaEventRecord(e0); launch(i); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (i >= 10) { sum += ms; if (ms < best) best = ms; }
        }
        printf("%-34s mean %.2f us  min %.2f us  (%s)\n", name, sum / 50 * 1e3, best * 1e3, cudaGetErrorString(cudaGetLastError()));
    };
    timeit("ordinary launch, empty", [&](int) { k_empty<<<grid, 256, 40 * 1024>>>(p); });
    timeit("cooperative launch, empty", [&](int) { cudaLaunchCooperativeKernel((void *)k_empty, dim3(grid), dim3(256), args, 40 * 1024, 0); });
    timeit("cooperative launch, grid.sync", [&](int) { cudaLaunchCooperativeKernel((void *)k_sync, dim3(grid), dim3(256), args, 40 * 1024, 0); });
    unsigned int target = 0;
    timeit("ordinary launch, own barrier", [&](int) { target += grid; k_own<<<grid, 256, 40 * 1024>>>(p, target); });
    return 0;
}
