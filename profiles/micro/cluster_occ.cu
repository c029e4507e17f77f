// How many clusters of 2 / 4 / 8 one-CTA-per-SM blocks (200 KB dynamic smem, 640 threads) are co-resident on this GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_occ cluster_occ.cu && ./cluster_occ
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(640, 1) k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t q{};
        q.gridDim = dim3(sms / cs * cs); q.blockDim = dim3(640); q.dynamicSmemBytes = 200 * 1024;
        cudaLaunchAttribute a[1];
        a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
        q.attrs = a; q.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &q);
        printf("SMs %d cluster %2d: max active clusters %d (= %d CTAs) %s\n", sms, cs, n, n * cs, cudaGetErrorString(e));
    }
    return 0;
}
