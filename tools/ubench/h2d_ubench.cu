// h2d_ubench.cu -- how fast do narrow strided (2D/3D) DMA copies and SM zero-copy reads move
// pinned host memory to the device?  Informs the host path's "lower triangle of Q only" transfer.
//   nvcc -O2 -arch=sm_100a -o h2d_ubench h2d_ubench.cu && ./h2d_ubench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void zc_read(const double2* __restrict__ src, double2* __restrict__ dst, size_t rows, int row_d2, int width_d2) {
  // warp per row: lanes read consecutive 16-byte chunks of the first `width_d2` chunks of each row
  const int lane = threadIdx.x & 31;
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
  double2 acc = make_double2(0, 0);
  for (size_t r = warp; r < rows; r += nw)
    for (int c = lane; c < width_d2; c += 32) { double2 v = src[r * row_d2 + c]; acc.x += v.x; acc.y += v.y; }
  if (acc.x == 1234.5) dst[0] = acc;
}

int main() {
  const int n = 60; const size_t B = 1 << 15;  // 32768 QPs x 60 x 60 doubles = 944 MB
  const size_t bytes = B * n * n * 8;
  double *h, *d; CK(cudaMallocHost(&h, bytes)); CK(cudaMalloc(&d, bytes)); memset(h, 1, bytes);
  cudaStream_t st; CK(cudaStreamCreate(&st)); cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto timeit = [&](const char* name, size_t moved, auto fn) {
    fn(); CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st)); fn(); CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-44s %8.2f ms  %6.1f GB/s (bytes moved %.0f MB)\n", name, ms, moved / ms * 1e-6, moved * 1e-6);
  };
  timeit("contiguous memcpyAsync", bytes, [&] { CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st)); });
  // 16 chunks like the host path
  timeit("contiguous, 16 chunks", bytes, [&] { for (int c = 0; c < 16; ++c) CK(cudaMemcpyAsync((char*)d + bytes / 16 * c, (char*)h + bytes / 16 * c, bytes / 16, cudaMemcpyHostToDevice, st)); });
  for (int w : {60, 48, 32, 30, 16, 8}) {
    char name[64]; snprintf(name, 64, "2D all rows, width %d B of pitch 480", w * 8);
    timeit(name, B * n * w * 8, [&] { CK(cudaMemcpy2DAsync(d, n * 8, h, n * 8, w * 8, B * n, cudaMemcpyHostToDevice, st)); });
  }
  for (int bands : {2, 4, 8}) {
    const int rb = (n + bands - 1) / bands; size_t moved = 0;
    for (int k = 0; k < bands; ++k) { int r0 = k * rb, r1 = r0 + rb < n ? r0 + rb : n; moved += B * (size_t)(r1 - r0) * r1 * 8; }
    char name[64]; snprintf(name, 64, "3D lower-triangle bands x%d, 16 chunks", bands);
    timeit(name, moved, [&] {
      for (int c = 0; c < 16; ++c)
        for (int k = 0; k < bands; ++k) {
          int r0 = k * rb, r1 = r0 + rb < n ? r0 + rb : n;
          cudaMemcpy3DParms p = {};
          p.srcPtr = make_cudaPitchedPtr(h, n * 8, n * 8, n); p.dstPtr = make_cudaPitchedPtr(d, n * 8, n * 8, n);
          p.srcPos = make_cudaPos(0, r0, B / 16 * c); p.dstPos = p.srcPos;
          p.extent = make_cudaExtent((size_t)r1 * 8, r1 - r0, B / 16); p.kind = cudaMemcpyHostToDevice;
          CK(cudaMemcpy3DAsync(&p, st));
        }
    });
  }
  for (int w : {30, 15, 4}) {   // 16-byte chunks per 480-byte row
    for (int grid : {148, 592, 2368}) {
      char name[64]; snprintf(name, 64, "SM zero-copy read, %d B of each row, grid %d", w * 16, grid);
      timeit(name, B * n * (size_t)w * 16, [&] { zc_read<<<grid, 128, 0, st>>>((const double2*)h, (double2*)d, B * n, n / 2, w); });
    }
  }
  return 0;
}
