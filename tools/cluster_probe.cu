// Probe: how many 16-CTA clusters with 222 KB of dynamic shared memory can be co-resident on this GPU?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* x) { extern __shared__ char s[]; if (x) x[0] = s[0]; }
int main() {
    for (int cs : {2, 4, 8, 16}) {
        for (int smem : {100 * 1024, 222 * 1024}) {
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(128); cfg.blockDim = dim3(416); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension;
            a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
            cfg.attrs = a; cfg.numAttrs = 1;
            int n = -1;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
            printf("cluster %2d smem %3d KB: max active clusters %d (%s)\n", cs, smem / 1024, n, cudaGetErrorString(e));
        }
    }
    return 0;
}
