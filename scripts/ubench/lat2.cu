// Global-memory latencies that matter for the grid path: fences, global atomics, the software grid barrier.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
__global__ void k1(unsigned long long *out, uint32_t *g) {
  uint32_t x = threadIdx.x;
  long long t0, t1;
  int s = 0;
  const int N = 64;
  t0 = clock64();
  for (int i = 0; i < N; i++) { g[1024 + (x & 63)] = x; __threadfence(); x++; }
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  t0 = clock64();
  for (int i = 0; i < N; i++) { g[1024 + (x & 63)] = x; __threadfence_block(); x++; }
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  t0 = clock64();
  for (int i = 0; i < N; i++) x = atomicAdd(g + 4096 + (x & 63), 1u) + 1;
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  t0 = clock64();
  for (int i = 0; i < N; i++) x = atomicCAS(g + 4096 + (x & 63), x, x + 1) + 1;
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  t0 = clock64();
  for (int i = 0; i < N; i++) x = __ldcg(g + (x & 1023)) + 1;
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  t0 = clock64();
  for (int i = 0; i < N; i++) { g[2048 + (x & 63)] = x; asm volatile("fence.acq_rel.gpu;" ::: "memory"); x++; }
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  if (x == 0xFFFFFFFF) out[63] = x;
}
__global__ void k_coop(unsigned long long *out) {
  cg::grid_group grid = cg::this_grid();
  long long t0 = clock64();
  for (int i = 0; i < 200; i++) grid.sync();
  long long t1 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[8] = (t1 - t0) / 200;
}
__device__ unsigned int bar_count, bar_gen;
__global__ void k_soft(unsigned long long *out) {
  long long t0 = clock64();
  for (int i = 0; i < 200; i++) {
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int gen = *(volatile unsigned int *)&bar_gen;
      __threadfence();
      if (atomicAdd(&bar_count, 1u) == gridDim.x - 1) { bar_count = 0; __threadfence(); atomicAdd(&bar_gen, 1u); }
      else while (*(volatile unsigned int *)&bar_gen == gen) __nanosleep(32);
      __threadfence();
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[9] = (t1 - t0) / 200;
}
int main() {
  unsigned long long *o; uint32_t *g;
  cudaMalloc(&o, 64 * 8); cudaMalloc(&g, 16384 * 4);
  cudaMemset(o, 0, 64 * 8); cudaMemset(g, 0, 16384 * 4);
  k1<<<1, 32>>>(o, g);
  cudaDeviceSynchronize();
  void *args[] = {&o};
  cudaLaunchCooperativeKernel((void *)k_coop, dim3(148), dim3(256), args, 0, 0);
  cudaDeviceSynchronize();
  cudaLaunchCooperativeKernel((void *)k_soft, dim3(148), dim3(256), args, 0, 0);
  cudaDeviceSynchronize();
  unsigned long long h[64];
  cudaMemcpy(h, o, sizeof(h), cudaMemcpyDeviceToHost);
  const char *nm[] = {"STG + __threadfence()", "STG + __threadfence_block()", "atomicAdd global dep", "atomicCAS global dep", "ld.cg dep (L2 hit)", "STG + fence.acq_rel.gpu"};
  for (int i = 0; i < 6; i++) printf("%-30s %llu cycles\n", nm[i], h[i]);
  printf("%-30s %llu cycles\n", "cg grid.sync (148x256)", h[8]);
  printf("%-30s %llu cycles\n", "software grid barrier", h[9]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
