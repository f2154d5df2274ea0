// Probe: how many clusters of size 2 / 4 / 8 with a full-smem 384-thread CTA can be co-resident on this GPU?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(384, 1) dummy(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int smem = (int)prop.sharedMemPerBlockOptin;
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  printf("SMs %d smem optin %d\n", prop.multiProcessorCount, smem);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(148 / cs * cs); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("cluster %2d: max active clusters %d (%d SMs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
