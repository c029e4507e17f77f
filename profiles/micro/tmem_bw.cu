// Micro-benchmark: TMEM -> register bandwidth of tcgen05.ld.32x32b.x32 on one SM (and chip-wide), sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../new-vit_b200/csrc/ptx.cuh"
using namespace mst;

__global__ void __launch_bounds__(512, 1) tmem_read_kernel(int iters, int nwarps, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(&tptr);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t base = tptr + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < nwarps) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(base + ((warp >> 2) * 256 + c * 32) % 512, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) acc ^= r[i];
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678) sink[0] = acc;
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tptr);
}

int main() {
    long long* d_cycles; uint32_t* d_sink;
    cudaMalloc(&d_cycles, 148 * sizeof(long long)); cudaMalloc(&d_sink, 4);
    const int iters = 2000;
    for (int grid : {1, 148})
        for (int nwarps : {1, 4, 8, 16}) {
            tmem_read_kernel<<<grid, 512>>>(iters, nwarps, d_cycles, d_sink);
            cudaDeviceSynchronize();
            tmem_read_kernel<<<grid, 512>>>(iters, nwarps, d_cycles, d_sink);
            cudaError_t e = cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, d_cycles, sizeof(c), cudaMemcpyDeviceToHost);
            const double bytes = double(iters) * 8 * 32 * 32 * 4 * nwarps;  // per SM
            printf("grid %3d warps %2d: %lld cycles, %.1f B/clk/SM (%s)\n", grid, nwarps, c, bytes / c, cudaGetErrorString(e));
        }
    return 0;
}
