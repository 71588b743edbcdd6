// Microbenchmark: bandwidth of a KERNEL reading pinned host memory over PCIe (zero-copy, uint4 loads) next to
// cudaMemcpyAsync from the same buffer: would a device-driven upload (no host round trip per chunk) keep the link busy?
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o pcie_read pcie_read.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int UNROLL>
__global__ void __launch_bounds__(256) copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n16; i += UNROLL * stride) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = __ldcs(src + i + u * stride);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) dst[i + u * stride] = v[u];
    }
    for (; i < n16; i += stride) dst[i] = __ldcs(src + i);
}

int main() {
    const size_t bytes = (size_t)316 << 20;
    void *h, *d;
    CK(cudaMallocHost(&h, bytes));
    CK(cudaMalloc(&d, bytes));
    for (size_t i = 0; i < bytes / 4; i += 1024) ((float*)h)[i] = (float)i;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("cudaMemcpyAsync H2D            : %.3f ms  %.1f GB/s\n", ms, bytes / ms * 1e-6);
    const int grids[] = {4, 8, 16, 32, 64, 148};
    for (int g : grids) {
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            copy_kernel<8><<<g, 256>>>((const uint4*)h, (uint4*)d, bytes / 16);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        printf("kernel, %3d CTAs x 256, 8 x 16 B in flight per thread: %.3f ms  %.1f GB/s\n", g, ms, bytes / ms * 1e-6);
    }
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        copy_kernel<2><<<32, 256>>>((const uint4*)h, (uint4*)d, bytes / 16);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("kernel,  32 CTAs x 256, 2 x 16 B in flight per thread: %.3f ms  %.1f GB/s\n", ms, bytes / ms * 1e-6);
    return 0;
}
