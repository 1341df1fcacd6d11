// How many clusters of size 2 / 4 (one CTA per SM: 200+ KB of dynamic shared memory) can be resident at once?
// Build + run (under gpurun): nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/cp tools/cluster_probe.cu && /tmp/cp
// B200, 225 KB of dynamic shared memory per CTA: cluster 1 -> 148 SMs, 2 -> 148, 4 -> 132 (33 clusters), 8 -> 120, 16 -> 112.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  printf("%s SMs %d\n", pr.name, pr.multiProcessorCount);
  const int smem = 225 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(cs * 148); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster %2d: max active clusters %d -> %d SMs (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}
