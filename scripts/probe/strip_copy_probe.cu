// Probe: does the access PATTERN of the rolling conv kernel (a CTA walks down a 128-pixel column strip of an NCHW tensor:
// per row 32 channel planes x 512 bytes in, 32 x 512 bytes out) cap the achievable DRAM bandwidth, compared with the same
// bytes moved as one contiguous stream or in a channels-last layout (per row one 16 KB piece)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o strip_copy_probe strip_copy_probe.cu && ./strip_copy_probe
#include <cuda_runtime.h>
#include <stdio.h>
constexpr int B = 16, C = 32, H = 436, W = 1024;
// mode 0: NCHW strips (the kernel's pattern)   mode 1: NHWC strips (128 px x 32 ch = 16 KB contiguous per row)   mode 2: flat
__global__ void __launch_bounds__(512) k(const float4* __restrict__ x, float4* __restrict__ y, int mode, int L) {
  const int tid = threadIdx.x;
  const int segs = (H + L - 1) / L, xt = W / 128;
  const int items = B * segs * xt;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / (segs * xt), r = item % (segs * xt), seg = r / xt, x0 = (r % xt) * 128;
    const int ya = seg * L, nr = min(L, H - ya);
    for (int row = ya; row < ya + nr; ++row) {
      // 128 px x 32 ch = 4096 floats = 1024 float4 per row: 512 threads x 2
      for (int u = tid; u < 1024; u += 512) {
        size_t idx;
        if (mode == 0) { const int c = u >> 5, q = u & 31; idx = (((size_t)b * C + c) * H + row) * (W / 4) + x0 / 4 + q; }
        else if (mode == 1) { const int px = u >> 3, q = u & 7; idx = ((((size_t)b * H + row) * W + x0 + px) * C) / 4 + q; }
        else { idx = ((size_t)item * 0 + ((size_t)b * segs * xt + r) ) * 0 + (((size_t)(b * H + row) * xt + (x0 >> 7)) * 1024 + u); }
        y[idx] = __ldg(x + idx);
      }
    }
  }
}
int main() {
  const size_t n = (size_t)B * C * H * W;
  float4 *x, *y;
  cudaMalloc(&x, n * 4); cudaMalloc(&y, n * 4);
  cudaMemset(x, 0, n * 4);
  const char* names[3] = {"NCHW column strips (32 x 512 B per row)", "NHWC column strips (16 KB per row)      ", "flat (item-major contiguous)           "};
  for (int L : {55, 109}) for (int mode = 0; mode < 3; ++mode) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148, 512>>>(x, y, mode, L); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) k<<<148, 512>>>(x, y, mode, L);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    printf("L=%3d %s: %.3f ms  %.0f GB/s (read + write)   %s\n", L, names[mode], ms, 2.0 * n * 4 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  // more bytes in flight: 4 CTAs per SM
  for (int mode = 0; mode < 3; ++mode) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148 * 4, 512>>>(x, y, mode, 55); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) k<<<148 * 4, 512>>>(x, y, mode, 55);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    printf("4 CTA/SM %s: %.3f ms  %.0f GB/s\n", names[mode], ms, 2.0 * n * 4 / ms / 1e6);
  }
  return 0;
}
