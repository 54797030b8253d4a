// Which hardware warp slots (and so which of the SM's four schedulers, slot % 4) do the warps of
// co-resident 128-thread CTAs get?  nvcc -O2 -arch=sm_100a -o warpid_ubench warpid_ubench.cu
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k(int* out) {
  extern __shared__ double sm[];
  unsigned smid, wid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
  if ((threadIdx.x & 31) == 0) { out[(blockIdx.x * 4 + (threadIdx.x >> 5)) * 2] = smid; out[(blockIdx.x * 4 + (threadIdx.x >> 5)) * 2 + 1] = wid; }
  sm[threadIdx.x] = 1.0;
  long long t0 = clock64(); while (clock64() - t0 < 200000) { }
}
int main() {
  int* d; cudaMalloc(&d, 592 * 4 * 2 * sizeof(int));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 53840);
  k<<<592, 128, 53840>>>(d);
  static int h[592 * 4 * 2]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int sm = 0; sm < 3; ++sm) {
    printf("SM %d:", sm);
    for (int b = 0; b < 592; ++b) if (h[b * 8] == sm) { printf("  block %d warps->slots", b); for (int w = 0; w < 4; ++w) printf(" %d", h[(b * 4 + w) * 2 + 1]); }
    printf("\n");
  }
  int hist[4][4] = {};
  for (int b = 0; b < 592; ++b) for (int w = 0; w < 4; ++w) hist[w][h[(b * 4 + w) * 2 + 1] % 4]++;
  for (int w = 0; w < 4; ++w) printf("warp %d -> scheduler histogram: %d %d %d %d\n", w, hist[w][0], hist[w][1], hist[w][2], hist[w][3]);
  return 0;
}
