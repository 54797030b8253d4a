// FP64 microbenchmarks on B200: DFMA latency/throughput, DMMA shapes, rcp chain, LDS latency.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_ubench fp64_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)

__global__ void dfma_latency(double* out, long long* cyc, int iters, double a, double b) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) x = fma(x, a, b);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void dfma_tput(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) x[k] = out[threadIdx.x] + k;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += x[k];
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__device__ __forceinline__ double fast_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  return x;
}
__global__ void rcp_latency(double* out, long long* cyc, int iters, int mode) {
  double x = out[threadIdx.x] + 1.5;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (mode == 0) x = fast_rcp(x) + 1.25;
      else if (mode == 1) x = 1.0 / x + 1.25;
      else x = sqrt(x) + 1.25;
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// DMMA m8n8k4
template <int ILP>
__global__ void dmma884(double* out, long long* cyc, int iters) {
  double c[ILP][2];
  double a = out[threadIdx.x], b = a + 1;
#pragma unroll
  for (int k = 0; k < ILP; ++k) { c[k][0] = k; c[k][1] = -k; }
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += c[k][0] + c[k][1];
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// DMMA m16n8k8 : A 4 regs, B 2 regs, C 4 regs
template <int ILP>
__global__ void dmma1688(double* out, long long* cyc, int iters) {
  double c[ILP][4];
  double a0 = out[threadIdx.x], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 - 1, b1 = a0 - 2;
#pragma unroll
  for (int k = 0; k < ILP; ++k) { c[k][0] = k; c[k][1] = -k; c[k][2] = 1; c[k][3] = 2; }
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[k][0]), "+d"(c[k][1]), "+d"(c[k][2]), "+d"(c[k][3])
                   : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lds_latency(double* out, long long* cyc, int iters) {
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 17 + 1) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) p = idx[p];
  }
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void bar_latency(double* out, long long* cyc, int iters) {
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void shfl_latency(double* out, long long* cyc, int iters) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  double* out; long long* cyc;
  CK(cudaMalloc(&out, 1 << 24)); CK(cudaMemset(out, 0, 1 << 24));
  CK(cudaMallocManaged(&cyc, 8 * 4096));
  int iters = 2000;
  dfma_latency<<<1, 32>>>(out, cyc, iters, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
  printf("DFMA dependent latency: %.2f cyc\n", (double)cyc[0] / (iters * 16));
  for (int mode = 0; mode < 3; ++mode) {
    rcp_latency<<<1, 32>>>(out, cyc, iters, mode); CK(cudaDeviceSynchronize());
    printf("%s + DADD dependent latency: %.2f cyc\n", mode == 0 ? "fast_rcp(mufu+2 newton)" : mode == 1 ? "IEEE 1.0/x" : "sqrt", (double)cyc[0] / (iters * 8));
  }
  lds_latency<<<1, 32>>>(out, cyc, iters); CK(cudaDeviceSynchronize());
  printf("LDS dependent latency: %.2f cyc\n", (double)cyc[0] / (iters * 16));
  shfl_latency<<<1, 32>>>(out, cyc, iters); CK(cudaDeviceSynchronize());
  printf("SHFL(f64)+DADD dependent latency: %.2f cyc\n", (double)cyc[0] / (iters * 16));
  for (int th : {32, 128, 256, 512}) {
    bar_latency<<<1, th>>>(out, cyc, iters); CK(cudaDeviceSynchronize());
    printf("__syncthreads (%d thr): %.2f cyc\n", th, (double)cyc[0] / (iters * 16));
  }
  // throughput: per SM, vary warps
  for (int warps : {1, 2, 4, 8, 16, 32}) {
    dfma_tput<8><<<148, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
    double fma_per_clk = (double)iters * 4 * 8 * warps * 32 / (double)cyc[0];
    printf("DFMA ILP8 %2d warps/SM: %.1f FMA/clk/SM\n", warps, fma_per_clk);
  }
  for (int warps : {1, 4}) {
    dfma_tput<2><<<148, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
    printf("DFMA ILP2 %2d warps/SM: %.1f FMA/clk/SM\n", warps, (double)iters * 4 * 2 * warps * 32 / (double)cyc[0]);
    dfma_tput<4><<<148, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
    printf("DFMA ILP4 %2d warps/SM: %.1f FMA/clk/SM\n", warps, (double)iters * 4 * 4 * warps * 32 / (double)cyc[0]);
    dfma_tput<16><<<148, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
    printf("DFMA ILP16 %2d warps/SM: %.1f FMA/clk/SM\n", warps, (double)iters * 4 * 16 * warps * 32 / (double)cyc[0]);
  }
  for (int warps : {1, 4, 8, 16}) {
    dmma884<1><<<148, warps * 32>>>(out, cyc, iters); CK(cudaDeviceSynchronize());
    printf("DMMA m8n8k4 ILP1 %2d warps/SM: %.1f FMA/clk/SM (lat %.1f cyc)\n", warps, (double)iters * 1 * 256 * warps / (double)cyc[0], (double)cyc[0] / iters);
    dmma884<4><<<148, warps * 32>>>(out, cyc, iters); CK(cudaDeviceSynchronize());
    printf("DMMA m8n8k4 ILP4 %2d warps/SM: %.1f FMA/clk/SM\n", warps, (double)iters * 4 * 256 * warps / (double)cyc[0]);
    dmma884<8><<<148, warps * 32>>>(out, cyc, iters); CK(cudaDeviceSynchronize());
    printf("DMMA m8n8k4 ILP8 %2d warps/SM: %.1f FMA/clk/SM\n", warps, (double)iters * 8 * 256 * warps / (double)cyc[0]);
    dmma1688<1><<<148, warps * 32>>>(out, cyc, iters); CK(cudaDeviceSynchronize());
    printf("DMMA m16n8k8 ILP1 %2d warps/SM: %.1f FMA/clk/SM (lat %.1f cyc)\n", warps, (double)iters * 1 * 1024 * warps / (double)cyc[0], (double)cyc[0] / iters);
    dmma1688<4><<<148, warps * 32>>>(out, cyc, iters); CK(cudaDeviceSynchronize());
    printf("DMMA m16n8k8 ILP4 %2d warps/SM: %.1f FMA/clk/SM\n", warps, (double)iters * 4 * 1024 * warps / (double)cyc[0]);
  }
  return 0;
}
