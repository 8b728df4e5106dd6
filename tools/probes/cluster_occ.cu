// How many thread-block clusters of a persistent one-CTA-per-SM kernel (222 KB of shared memory, 512 threads) can be
// resident at once on this GPU, per cluster size — decides whether a 4-CTA weight multicast would idle SMs.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dummy(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
    cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, 226816);
    cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int cs : {1, 2, 4, 8}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(sms / cs * cs); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 226816;
        cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = cs; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
        printf("cluster size %d: max active clusters %d (= %d of %d SMs) %s\n", cs, n, n * cs, sms, cudaGetErrorString(e));
    }
    return 0;
}
