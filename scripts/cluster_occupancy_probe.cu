// Prints cudaOccupancyMaxActiveClusters for cluster sizes 2..16 of a 352-thread kernel with ~215 KB of dynamic shared memory
// (the footprint of the cluster-resident decode kernel), and checks co-residency for real: every cluster of a launch
// waits until all clusters of the launch have started (a launch that is not co-resident times out instead).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(352, 1) probe(int* counter, int want, int* ok) {
  extern __shared__ char sm[];
  sm[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    atomicAdd(counter, 1);
    long long t0 = clock64();
    while (atomicAdd(counter, 0) < want && clock64() - t0 < 400000000LL) {}
    if (atomicAdd(counter, 0) >= want) atomicAdd(ok, 1);
  }
}
int main() {
  const int smem = 215 * 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  int *d; cudaMalloc(&d, 8);
  for (int cs : {2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 32); cfg.blockDim = dim3(352); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int mc = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&mc, probe, &cfg);
    printf("cluster size %2d: cudaOccupancyMaxActiveClusters = %d (%s)\n", cs, mc, cudaGetErrorString(e));
    for (int n = mc - 1; n <= mc + 1; ++n) {
      if (n < 1) continue;
      cudaMemset(d, 0, 8);
      cfg.gridDim = dim3(cs * n);
      e = cudaLaunchKernelEx(&cfg, probe, d, cs * n, d + 1);
      cudaError_t e2 = cudaDeviceSynchronize();
      int h[2]; cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
      printf("   %2d clusters launched: %d of %d CTAs saw everyone start (%s / %s)\n", n, h[1], cs * n, cudaGetErrorString(e), cudaGetErrorString(e2));
    }
  }
  return 0;
}
