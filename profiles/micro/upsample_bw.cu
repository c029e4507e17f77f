// Write-bandwidth probe for the saliency upsampler: how fast can 1.65 GB of fp32 be written on this GPU, and which launch
// shape gets there?  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/micro/upsample_bw profiles/micro/upsample_bw.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int ROWS, int UNROLL, bool STREAM>
__global__ void __launch_bounds__(256) up_kernel(const float* __restrict__ coarse, float* __restrict__ full, int gh, int gw, int H, int W,
                                                   int crow_max) {
    extern __shared__ float hrow[];
    const int s = blockIdx.x, y_begin = blockIdx.y * ROWS, y_end = min(H, y_begin + ROWS);
    const float sy = (float)gh / H, sx = (float)gw / W;
    const int r_first = (int)fmaxf(sy * (y_begin + 0.5f) - 0.5f, 0.f);
    const float* c = coarse + (long long)s * gh * gw;
    for (int idx = threadIdx.x; idx < crow_max * W; idx += blockDim.x) {
        const int rr = idx / W, x = idx - rr * W;
        const int r = min(r_first + rr, gh - 1);
        const float fx = fmaxf(sx * (x + 0.5f) - 0.5f, 0.f);
        const int x0 = (int)fx, x1 = min(x0 + 1, gw - 1);
        const float lx = fx - x0, hx = 1.f - lx;
        hrow[idx] = hx * __ldg(c + r * gw + x0) + lx * __ldg(c + r * gw + x1);
    }
    __syncthreads();
    float* out = full + (long long)s * H * W;
    const int W4 = W >> 2, tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    for (int y = y_begin + ty * UNROLL; y < y_end; y += 4 * UNROLL) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int yy = y + u;
            if (yy >= y_end) break;
            const float fy = fmaxf(sy * (yy + 0.5f) - 0.5f, 0.f);
            const int y0 = (int)fy, y1 = min(y0 + 1, gh - 1);
            const float ly = fy - y0, hy = 1.f - ly;
            const float4* r0 = (const float4*)(hrow + (y0 - r_first) * W);
            const float4* r1 = (const float4*)(hrow + (y1 - r_first) * W);
            for (int cg = tx; cg < W4; cg += 64) {
                const float4 a = r0[cg], b = r1[cg];
                const float4 o = make_float4(hy * a.x + ly * b.x, hy * a.y + ly * b.y, hy * a.z + ly * b.z, hy * a.w + ly * b.w);
                if (STREAM) __stcs((float4*)(out + (long long)yy * W) + cg, o); else ((float4*)(out + (long long)yy * W))[cg] = o;
            }
        }
    }
}
// flat variant: one thread = one float4 of the whole slice, grid-stride over the slice (fully coalesced 512 B per warp)
template <bool STREAM>
__global__ void __launch_bounds__(256) up_flat(const float* __restrict__ coarse, float* __restrict__ full, int gh, int gw, int H, int W) {
    extern __shared__ float hrow[];   // [gh][W]
    const int s = blockIdx.x;
    const float sy = (float)gh / H, sx = (float)gw / W;
    const float* c = coarse + (long long)s * gh * gw;
    for (int idx = threadIdx.x; idx < gh * W; idx += blockDim.x) {
        const int r = idx / W, x = idx - r * W;
        const float fx = fmaxf(sx * (x + 0.5f) - 0.5f, 0.f);
        const int x0 = (int)fx, x1 = min(x0 + 1, gw - 1);
        const float lx = fx - x0, hx = 1.f - lx;
        hrow[idx] = hx * __ldg(c + r * gw + x0) + lx * __ldg(c + r * gw + x1);
    }
    __syncthreads();
    float4* out = (float4*)(full + (long long)s * H * W);
    const int W4 = W >> 2, total = H * W4;
    for (int i = threadIdx.x; i < total; i += 256) {
        const int y = i / W4, cg = i - y * W4;
        const float fy = fmaxf(sy * (y + 0.5f) - 0.5f, 0.f);
        const int y0 = (int)fy, y1 = min(y0 + 1, gh - 1);
        const float ly = fy - y0, hy = 1.f - ly;
        const float4 a = ((const float4*)(hrow + y0 * W))[cg], b = ((const float4*)(hrow + y1 * W))[cg];
        const float4 o = make_float4(hy * a.x + ly * b.x, hy * a.y + ly * b.y, hy * a.z + ly * b.z, hy * a.w + ly * b.w);
        if (STREAM) __stcs(out + i, o); else out[i] = o;
    }
}
__global__ void fill_kernel(float4* p, long long n4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        p[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

template <typename F> float time_ms(F f, int reps = 10) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    const int BD = 8192, gh = 16, gw = 16, H = 224, W = 224;
    const long long n = (long long)BD * H * W;
    float *coarse, *full;
    CK(cudaMalloc(&coarse, (size_t)BD * gh * gw * 4)); CK(cudaMalloc(&full, n * 4));
    CK(cudaMemset(coarse, 0, (size_t)BD * gh * gw * 4));
    const double gb = n * 4 / 1e9;
    float ms = time_ms([&] { cudaMemsetAsync(full, 0, n * 4); });
    printf("cudaMemset            %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = time_ms([&] { fill_kernel<<<148 * 8, 256>>>((float4*)full, n / 4); });
    printf("fill kernel (148x8)   %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = time_ms([&] { fill_kernel<<<148 * 32, 512>>>((float4*)full, n / 4); });
    printf("fill kernel (148x32)  %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
#define RUN(ROWS, UN, ST)                                                                                                   \
    {                                                                                                                       \
        const int crow = (int)((double)ROWS * gh / H) + 3;                                                                  \
        cudaFuncSetAttribute(up_kernel<ROWS, UN, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);             \
        ms = time_ms([&] { up_kernel<ROWS, UN, ST><<<dim3(BD, (H + ROWS - 1) / ROWS), 256, crow * W * 4>>>(coarse, full, gh, gw, H, W, crow); }); \
        printf("rows %3d unroll %d stream %d  %.3f ms  %.0f GB/s\n", ROWS, UN, (int)ST, ms, gb / ms * 1e3);                   \
    }
    RUN(56, 1, true) RUN(56, 1, false) RUN(56, 2, true) RUN(112, 2, true) RUN(224, 2, true) RUN(224, 4, true) RUN(28, 1, true)
    cudaFuncSetAttribute(up_flat<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(up_flat<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    ms = time_ms([&] { up_flat<true><<<BD, 256, gh * W * 4>>>(coarse, full, gh, gw, H, W); });
    printf("flat stream           %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = time_ms([&] { up_flat<false><<<BD, 256, gh * W * 4>>>(coarse, full, gh, gw, H, W); });
    printf("flat plain            %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    CK(cudaDeviceSynchronize());
    return 0;
}
