// Latency microbenchmarks for the single-warp path (one warp, one SM): cycles per dependent operation.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 256
__device__ __noinline__ uint32_t callee(uint32_t x, uint32_t *p) { return p[x & 63] + 1; }
__global__ void k(unsigned long long *out, uint32_t *g) {
  __shared__ uint32_t sh[4096];
  __shared__ unsigned long long sh64[512];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = (i * 17 + 1) & 4095;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) sh64[i] = 0;
  __syncthreads();
  uint32_t x = threadIdx.x & 31;
  long long t0, t1;
  int s = 0;
  // 0: clock overhead
  t0 = clock64(); t1 = clock64(); if (threadIdx.x == 0) out[s] = t1 - t0; s++;
  // 1: dependent LDS chain
  t0 = clock64();
  for (int i = 0; i < N; i++) x = sh[x];
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  // 2: dependent volatile LDS chain
  t0 = clock64();
  for (int i = 0; i < N; i++) x = *(volatile uint32_t *)&sh[x & 4095];
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  // 3: LDS + membar.cta
  t0 = clock64();
  for (int i = 0; i < N; i++) { x = sh[x & 4095]; __threadfence_block(); }
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  // 4: STS + membar.cta + LDS
  t0 = clock64();
  for (int i = 0; i < N; i++) { sh[(x + 77) & 4095] = x; __threadfence_block(); x = sh[x & 4095]; }
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  // 5: shared atomicAdd dependent
  t0 = clock64();
  for (int i = 0; i < N; i++) x = atomicAdd(&sh[x & 4095], 1u) & 4095;
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  // 6: shared atomicCAS 32 dependent
  t0 = clock64();
  for (int i = 0; i < N; i++) x = atomicCAS(&sh[x & 4095], x, x + 1) & 4095;
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  // 7: shared atomicCAS 64 dependent
  t0 = clock64();
  for (int i = 0; i < N; i++) x = (uint32_t)atomicCAS(&sh64[x & 511], (unsigned long long)x, (unsigned long long)x + 1) & 511;
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  // 8: shared atomicMin 64
  t0 = clock64();
  for (int i = 0; i < N; i++) x = (uint32_t)atomicMin(&sh64[x & 511], (unsigned long long)x) & 511;
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  // 9: noinline call
  t0 = clock64();
  for (int i = 0; i < N; i++) x = callee(x, sh);
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  // 10: dependent IMAD chain
  t0 = clock64();
  for (int i = 0; i < N; i++) x = x * 0x9E3779B1u + 12345u;
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  // 11: ballot + popc dependent
  t0 = clock64();
  for (int i = 0; i < N; i++) x = __popc(__ballot_sync(0xFFFFFFFFu, x & 1)) + x;
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  // 12: shfl dependent
  t0 = clock64();
  for (int i = 0; i < N; i++) x = __shfl_sync(0xFFFFFFFFu, x, x & 31) + 1;
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  // 13: match_any dependent
  t0 = clock64();
  for (int i = 0; i < N; i++) x = __match_any_sync(0xFFFFFFFFu, x & 7) + x;
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  // 14: syncwarp
  t0 = clock64();
  for (int i = 0; i < N; i++) { __syncwarp(); x += 1; }
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  // 15: __syncthreads
  t0 = clock64();
  for (int i = 0; i < N; i++) { __syncthreads(); x += 1; }
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  // 16: FP64 dependent FMA
  double d = x;
  t0 = clock64();
  for (int i = 0; i < N; i++) d = __dmul_rn(d, 1.0000001);
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  // 17: FP64 divide dependent
  t0 = clock64();
  for (int i = 0; i < N; i++) d = __ddiv_rn(d, 1.0000001);
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  // 18: global ld.cg dependent (L2 hit)
  t0 = clock64();
  for (int i = 0; i < 64; i++) x = __ldcg(g + (x & 1023));
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / 64; s++;
  // 19: global store + membar.cta
  t0 = clock64();
  for (int i = 0; i < 64; i++) { g[1024 + (x & 63)] = x; __threadfence_block(); x++; }
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / 64; s++;
  // 20: LDS.128 dependent
  t0 = clock64();
  for (int i = 0; i < N; i++) { uint4 v = *(const uint4 *)&sh[(x & 1023) * 4]; x = v.x ^ v.w; }
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  // 21: floor + f2i
  t0 = clock64();
  for (int i = 0; i < N; i++) { d = (double)(long long)floor(d * 1.5) + 0.25; }
  t1 = clock64(); if (threadIdx.x == 0) out[s] = (t1 - t0) / N; s++;
  if (x == 0xFFFFFFFF || d == 123.0) out[63] = x;
}
int main() {
  unsigned long long *o; uint32_t *g;
  cudaMalloc(&o, 64 * 8); cudaMalloc(&g, 8192 * 4);
  cudaMemset(o, 0, 64 * 8); cudaMemset(g, 0, 8192 * 4);
  const char *nm[] = {"clock pair", "LDS dep", "LDS volatile dep", "LDS+membar.cta", "STS+membar.cta+LDS", "ATOMS.ADD dep", "ATOMS.CAS32 dep", "ATOMS.CAS64 dep",
                      "ATOMS.MIN64 dep", "noinline call+LDS", "IMAD dep", "ballot+popc dep", "shfl dep", "match_any dep", "syncwarp", "syncthreads", "DMUL dep", "DDIV dep",
                      "ld.cg L2 dep", "STG+membar.cta", "LDS.128 dep", "floor+f2i+i2f dep"};
  for (int threads = 32; threads <= 256; threads *= 8) {
    k<<<1, threads>>>(o, g);
    cudaDeviceSynchronize();
    unsigned long long h[64];
    cudaMemcpy(h, o, sizeof(h), cudaMemcpyDeviceToHost);
    printf("threads=%d\n", threads);
    for (int i = 0; i < 22; i++) printf("  %-22s %llu cycles\n", nm[i], h[i]);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
