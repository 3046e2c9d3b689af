// How many thread-block clusters of a persistent 512-thread kernel can be co-resident on this GPU, and what does a
// cluster barrier / a DSMEM read cost?  (Input to the batch-1 decode design: a head group as a cluster.)
// nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_probe cluster_probe.cu && ./cluster_probe
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(512, 1) probe_kernel(long long* out, int iters) {
  extern __shared__ float sm[];
  cg::cluster_group cl = cg::this_cluster();
  const unsigned rank = cl.block_rank(), n = cl.num_blocks();
  sm[threadIdx.x] = (float)(blockIdx.x * 1000 + threadIdx.x);
  cl.sync();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) cl.sync();
  long long t1 = clock64();
  float acc = 0.f;
  for (int i = 0; i < iters; ++i) {
    const float* remote = cl.map_shared_rank(sm, (rank + 1 + i) % n);
    acc += remote[(threadIdx.x + i) & 511];
    cl.sync();
  }
  long long t2 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = (t1 - t0) / iters; out[1] = (t2 - t1) / iters; out[2] = (long long)acc; }
  cl.sync();
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s SMs %d\n", p.name, p.multiProcessorCount);
  long long* d;
  cudaMalloc(&d, 64);
  for (int smem_kb : {48, 100, 160}) {
    for (int cs : {2, 4, 8, 16}) {
      cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
      if (cs > 8) cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * 16); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem_kb * 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int ncl = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, probe_kernel, &cfg);
      printf("smem %3d KB cluster %2d: max active clusters %d (%s) -> %d CTAs", smem_kb, cs, ncl, cudaGetErrorString(e), ncl * cs);
      if (e == cudaSuccess && ncl > 0) {
        cfg.gridDim = dim3(cs * ncl);
        e = cudaLaunchKernelEx(&cfg, probe_kernel, d, 200);
        cudaError_t e2 = cudaDeviceSynchronize();
        long long h[3] = {0, 0, 0};
        cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
        printf("  launch %s/%s: cluster.sync %lld clk, dsmem read + sync %lld clk", cudaGetErrorString(e), cudaGetErrorString(e2), h[0], h[1]);
      }
      printf("\n");
      cudaGetLastError();
    }
  }
  return 0;
}
