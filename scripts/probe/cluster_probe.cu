// How many clusters of 1-CTA-per-SM kernels (226 KB dynamic smem, 576 threads) can be co-resident on this GPU, per cluster size?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(576, 1) stub(int* p) { extern __shared__ char s[]; if (p && threadIdx.x == 0) p[blockIdx.x] = s[0]; }
int main() {
  const int smem = 226560;
  cudaFuncSetAttribute(stub, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(stub, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs = 1; cs <= 8; ++cs) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 40); cfg.blockDim = dim3(576); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, stub, &cfg);
    printf("cluster size %d: max active clusters %d (%d CTAs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
