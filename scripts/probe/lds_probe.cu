// Probe: shared-memory wavefronts of warp-wide LDS.128 / LDS.64 / LDS.32 under address sharing (broadcast) between lanes
// of DIFFERENT quarter-warps — decides whether the correlation kernel's f2-row-sharing thread layout makes the f2 loads
// cheaper than one wavefront per quarter warp.  Prints clk per warp-level load per SM (1 = one wavefront).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_probe lds_probe.cu && ./lds_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

template <int VEC>
__global__ void k(const int* __restrict__ offs, float* out, long long* clk, int iters) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + (uint32_t)offs[lane];
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t a = base + (uint32_t)(((it + u) & 7) * 2048);
      if (VEC == 4) {
        float x, y, z, w;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a));
        acc += x + y + z + w;
      } else if (VEC == 2) {
        float x, y;
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(a));
        acc += x + y;
      } else {
        float x;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a));
        acc += x;
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

int main() {
  float* out; long long* clk; int* doffs;
  cudaMalloc(&out, 64 << 20); cudaMalloc(&clk, 8); cudaMalloc(&doffs, 32 * 4);
  const int iters = 4000;
  struct Pat { const char* name; int vec; int off[32]; };
  Pat pats[16];
  int np = 0;
  auto add = [&](const char* n, int vec, int (*f)(int)) { pats[np].name = n; pats[np].vec = vec; for (int l = 0; l < 32; ++l) pats[np].off[l] = f(l); ++np; };
  add("v4 distinct 512B            ", 4, [](int l) { return l * 16; });
  add("v4 quarters identical 128B  ", 4, [](int l) { return (l % 8) * 16; });
  add("v4 4 chunks (l%4) 64B       ", 4, [](int l) { return (l % 4) * 16; });
  add("v4 4 chunks (l/8) 64B       ", 4, [](int l) { return (l / 8) * 16; });
  add("v4 all same 16B             ", 4, [](int l) { return 0; });
  add("v4 8 chunks (l/4) 128B      ", 4, [](int l) { return (l / 4) * 16; });
  add("v4 strips*32B (l%4)*32      ", 4, [](int l) { return (l % 4) * 32; });
  add("v4 8 rows pitch 176B x 4 strips 32B", 4, [](int l) { return (l / 4) * 176 + (l % 4) * 32; });
  add("v4 8 rows pitch 144B x 4 strips 32B", 4, [](int l) { return (l / 4) * 144 + (l % 4) * 32; });
  add("v2 distinct 256B            ", 2, [](int l) { return l * 8; });
  add("v2 halves identical 128B    ", 2, [](int l) { return (l % 16) * 8; });
  add("v2 4 chunks 32B             ", 2, [](int l) { return (l % 4) * 8; });
  add("v1 distinct 128B            ", 1, [](int l) { return l * 4; });
  add("v1 all same                 ", 1, [](int l) { return 0; });
  for (int warps = 4; warps <= 16; warps *= 2) {
    for (int p = 0; p < np; ++p) {
      cudaMemcpy(doffs, pats[p].off, 128, cudaMemcpyHostToDevice);
      long long c = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (pats[p].vec == 4) k<4><<<148, warps * 32, 32768>>>(doffs, out, clk, iters);
        else if (pats[p].vec == 2) k<2><<<148, warps * 32, 32768>>>(doffs, out, clk, iters);
        else k<1><<<148, warps * 32, 32768>>>(doffs, out, clk, iters);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
      printf("%2d warps/SM  %s : %.2f clk per warp-load per SM\n", warps, pats[p].name, (double)c / (iters * 8.0 * warps));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
