// cost of clock64() / barrier / smem round trip on B200
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, double* d) {
  __shared__ unsigned long long s[16];
  __shared__ double buf[256];
  long long t0 = clock64();
  long long acc = 0;
#pragma unroll
  for (int i = 0; i < 64; ++i) acc += clock64();
  long long t1 = clock64();
  // dependent chain: LDS -> DFMA -> STS -> syncwarp x 32
  double x = d[threadIdx.x];
  buf[threadIdx.x] = x;
  __syncwarp();
  long long t2 = clock64();
#pragma unroll
  for (int i = 0; i < 32; ++i) { x = fma(buf[(threadIdx.x + 1) & 31], 1.0000001, x); buf[threadIdx.x] = x; __syncwarp(); }
  long long t3 = clock64();
  // 64 x __syncthreads with 128 threads
#pragma unroll
  for (int i = 0; i < 64; ++i) __syncthreads();
  long long t4 = clock64();
  // IMAD dependent chain 64
  int y = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 64; ++i) y = y * 3 + i;
  long long t5 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t3 - t2; out[2] = t4 - t3; out[3] = t5 - t4; out[4] = acc + y; s[0] = 0; }
  d[threadIdx.x] = x;
}
int main() {
  long long* out; double* d;
  cudaMallocManaged(&out, 64); cudaMalloc(&d, 4096); cudaMemset(d, 0, 4096);
  k<<<1, 128>>>(out, d); cudaDeviceSynchronize();
  k<<<1, 128>>>(out, d); cudaDeviceSynchronize();
  printf("clock64 x64: %.1f cyc each\n", out[0] / 64.0);
  printf("LDS->DFMA->STS->syncwarp round: %.1f cyc each\n", out[1] / 32.0);
  printf("__syncthreads(128): %.1f cyc each\n", out[2] / 64.0);
  printf("IMAD dependent: %.1f cyc each\n", out[3] / 64.0);
  return 0;
}
