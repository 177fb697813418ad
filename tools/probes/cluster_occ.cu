// How many clusters of 2 / 4 / 8 CTAs with ~225 KB of dynamic shared memory each can be co-resident on this GPU?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* out) { extern __shared__ int sm[]; if (threadIdx.x == 0 && out) out[blockIdx.x] = sm[0]; }
int main() {
    const int smem = 225 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs : {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148 / cs * cs); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
        cfg.attrs = a; cfg.numAttrs = 1;
        int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf("cluster %2d: max active clusters %d (%d SMs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
