// How many clusters of size 1/2/4/8 with one 200 KB-smem CTA per SM can be co-resident?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { if (p) *p = 0; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148 / cs * cs);
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute a[1];
    a[0].id = cudaLaunchAttributeClusterDimension;
    a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    cfg.attrs = a; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster size %2d: max active clusters %d (%d SMs) %s\n", cs, n, n * cs, e ? cudaGetErrorString(e) : "");
  }
  return 0;
}
