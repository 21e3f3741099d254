// Probe: issue rate of 3-register FFMA vs FFMA2 (fma.rn.f32x2) in the correlation kernel's 8 px x 9 dx register tile,
// 1..4 warps per SM sub-partition.  Prints cycles per warp-level FMA instruction per SMSP.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__global__ void k_ffma(const float* in, float* out, long long* clk, int iters) {
  float a[8], b[16], acc[9][8];
  for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x + i];
  for (int i = 0; i < 16; ++i) b[i] = in[threadIdx.x + 8 + i];
  for (int d = 0; d < 9; ++d) for (int p = 0; p < 8; ++p) acc[d][p] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int d = 0; d < 9; ++d)
#pragma unroll
      for (int p = 0; p < 8; ++p) acc[d][p] = fmaf(a[p], b[p + d], acc[d][p]);
    a[it & 7] += 1.0f;   // keep the loop from being hoisted
  }
  long long t1 = clock64();
  float s = 0;
  for (int d = 0; d < 9; ++d) for (int p = 0; p < 8; ++p) s += acc[d][p];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
__device__ __forceinline__ void fma2(uint64_t& d, uint64_t a, uint64_t b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__global__ void k_ffma2(const float* in, float* out, long long* clk, int iters) {
  // row-pair formulation: 4 px x 9 dx x (2 rows packed) -> 36 packed accumulators
  uint64_t a[4], b[12], acc[9][4];
  const uint64_t* in2 = reinterpret_cast<const uint64_t*>(in);
  for (int i = 0; i < 4; ++i) a[i] = in2[threadIdx.x + i];
  for (int i = 0; i < 12; ++i) b[i] = in2[threadIdx.x + 4 + i];
  for (int d = 0; d < 9; ++d) for (int p = 0; p < 4; ++p) acc[d][p] = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int d = 0; d < 9; ++d)
#pragma unroll
      for (int p = 0; p < 4; ++p) fma2(acc[d][p], a[p], b[p + d]);
    a[it & 3] += 1;
  }
  long long t1 = clock64();
  uint64_t s = 0;
  for (int d = 0; d < 9; ++d) for (int p = 0; p < 4; ++p) s ^= acc[d][p];
  reinterpret_cast<uint64_t*>(out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
int main() {
  float *in, *out; long long* clk;
  cudaMalloc(&in, 1 << 20); cudaMemset(in, 0, 1 << 20); cudaMalloc(&out, 64 << 20); cudaMalloc(&clk, 8);
  const int iters = 2000;
  for (int warps = 4; warps <= 16; warps += 4) {
    long long c;
    k_ffma<<<148, warps * 32>>>(in, out, clk, iters); cudaDeviceSynchronize();
    k_ffma<<<148, warps * 32>>>(in, out, clk, iters); cudaDeviceSynchronize();
    cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    printf("FFMA : %2d warps/SM (%d per SMSP): %.2f clk per warp-FFMA per SMSP  (%.1f FMA lanes/clk/SM)\n", warps, warps / 4,
           (double)c / (iters * 72.0 * (warps / 4)), 72.0 * iters * warps * 32 / (double)c);
    k_ffma2<<<148, warps * 32>>>(in, out, clk, iters); cudaDeviceSynchronize();
    k_ffma2<<<148, warps * 32>>>(in, out, clk, iters); cudaDeviceSynchronize();
    cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    printf("FFMA2: %2d warps/SM (%d per SMSP): %.2f clk per warp-FFMA2 per SMSP (%.1f FMA lanes/clk/SM)\n", warps, warps / 4,
           (double)c / (iters * 36.0 * (warps / 4)), 72.0 * iters * warps * 32 / (double)c);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
