// Probe: the correlation kernel's inner loop in isolation — what does one channel step of an 8 x 32 tile (72 (row, dy)
// pairs x 4 strips of 8 px; 6 LDS.128 + 72 FFMA per thread) cost on one SM as a function of (a) the lane -> (pair, strip)
// mapping, (b) the load width, (c) the register tile?  No TMA, no barriers, the stage is filled once.
// Ideal FMA time of a channel step of one tile: 72 pairs * 4 strips * 72 FFMA / 128 lanes = 162 clk.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o corr_loop_probe corr_loop_probe.cu && ./corr_loop_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

constexpr int TH = 8, F1_P = 36, F2_H = 16, F2_P = 44, CC = 8;
constexpr int F1_ELEMS = CC * TH * F1_P, F2_ELEMS = CC * F2_H * F2_P;

__device__ __forceinline__ void pair_of(int p, int& r, int& dyi) {
  int h = 0;
  for (;;) {
    const int cnt = min(min(h, TH - 1), min(8, TH + 9 - 2 - h)) + 1;
    if (p < cnt) break;
    p -= cnt;
    ++h;
  }
  r = max(h - 8, 0) + p;
  dyi = h - r;
}

// MAP 0: lane = pair_in_warp * 4 + strip (the kernel today)   MAP 1: lane = strip * 8 + pair_in_warp (a quarter warp = one
// strip, 8 pairs sorted by f2 row)   VEC: 4 = LDS.128, 2 = LDS.64, 1 = LDS.32
template <int MAP, int VEC>
__global__ void __launch_bounds__(512, 1) k72(float* out, long long* clk, int iters, int nwarps) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < F1_ELEMS + F2_ELEMS; i += blockDim.x) sm[i] = (float)(i & 1023) * 1e-3f;
  __syncthreads();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp >= nwarps) return;
  const int s8 = MAP == 0 ? (lane & 3) : (lane >> 3);
  int p = (warp * 8 + (MAP == 0 ? (lane >> 2) : (lane & 7))) % 72, r, dyi;
  pair_of(p, r, dyi);
  float acc[9][8];
#pragma unroll
  for (int d = 0; d < 9; ++d)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[d][q] = 0.f;
  const float* f1s = sm;
  const float* f2s = sm + F1_ELEMS;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll 2
    for (int cc = 0; cc < CC; ++cc) {
      const float* ap = f1s + (cc * TH + r) * F1_P + s8 * 8;
      const float* bp = f2s + (cc * F2_H + r + dyi) * F2_P + s8 * 8;
      float a[8], bv[16];
      if (VEC == 4) {
#pragma unroll
        for (int q = 0; q < 2; ++q) { float4 t = reinterpret_cast<const float4*>(ap)[q]; a[4 * q] = t.x; a[4 * q + 1] = t.y; a[4 * q + 2] = t.z; a[4 * q + 3] = t.w; }
#pragma unroll
        for (int q = 0; q < 4; ++q) { float4 t = reinterpret_cast<const float4*>(bp)[q]; bv[4 * q] = t.x; bv[4 * q + 1] = t.y; bv[4 * q + 2] = t.z; bv[4 * q + 3] = t.w; }
      } else if (VEC == 2) {
#pragma unroll
        for (int q = 0; q < 4; ++q) { float2 t = reinterpret_cast<const float2*>(ap)[q]; a[2 * q] = t.x; a[2 * q + 1] = t.y; }
#pragma unroll
        for (int q = 0; q < 8; ++q) { float2 t = reinterpret_cast<const float2*>(bp)[q]; bv[2 * q] = t.x; bv[2 * q + 1] = t.y; }
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) a[q] = ap[q];
#pragma unroll
        for (int q = 0; q < 16; ++q) bv[q] = bp[q];
      }
#pragma unroll
      for (int d = 0; d < 9; ++d)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[d][q] = fmaf(a[q], bv[q + d], acc[d][q]);
    }
    asm volatile("" ::: "memory");
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int d = 0; d < 9; ++d)
#pragma unroll
    for (int q = 0; q < 8; ++q) s += acc[d][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

// 144-accumulator tile: a thread owns TWO tile rows (r, r+1) at displacement rows (dy, dy-1): both read f2 row r+dy.
// Per channel 4 + 4 LDS.128 for 144 FFMA (0.222 operand words per FMA instead of 0.333).  nwarps warps, all lanes busy.
template <int MAP>
__global__ void __launch_bounds__(256, 1) k144(float* out, long long* clk, int iters, int nwarps) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < F1_ELEMS + F2_ELEMS; i += blockDim.x) sm[i] = (float)(i & 1023) * 1e-3f;
  __syncthreads();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp >= nwarps) return;
  const int s8 = MAP == 0 ? (lane & 3) : (lane >> 3);
  const int slot = (warp * 8 + (MAP == 0 ? (lane >> 2) : (lane & 7))) % 40;   // 40 thread rows: 7 x ... (r, h) combos
  const int r = slot % 7, h = r + (slot / 7) + 1;                             // representative addresses only
  float acc0[9][8], acc1[9][8];
#pragma unroll
  for (int d = 0; d < 9; ++d)
#pragma unroll
    for (int q = 0; q < 8; ++q) { acc0[d][q] = 0.f; acc1[d][q] = 0.f; }
  const float* f1s = sm;
  const float* f2s = sm + F1_ELEMS;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int cc = 0; cc < CC; ++cc) {
      const float* ap = f1s + (cc * TH + r) * F1_P + s8 * 8;
      const float* bp = f2s + (cc * F2_H + h) * F2_P + s8 * 8;
      float a0[8], a1[8], bv[16];
#pragma unroll
      for (int q = 0; q < 2; ++q) { float4 t = reinterpret_cast<const float4*>(ap)[q]; a0[4 * q] = t.x; a0[4 * q + 1] = t.y; a0[4 * q + 2] = t.z; a0[4 * q + 3] = t.w; }
#pragma unroll
      for (int q = 0; q < 2; ++q) { float4 t = reinterpret_cast<const float4*>(ap + F1_P)[q]; a1[4 * q] = t.x; a1[4 * q + 1] = t.y; a1[4 * q + 2] = t.z; a1[4 * q + 3] = t.w; }
#pragma unroll
      for (int q = 0; q < 4; ++q) { float4 t = reinterpret_cast<const float4*>(bp)[q]; bv[4 * q] = t.x; bv[4 * q + 1] = t.y; bv[4 * q + 2] = t.z; bv[4 * q + 3] = t.w; }
#pragma unroll
      for (int d = 0; d < 9; ++d)
#pragma unroll
        for (int q = 0; q < 8; ++q) { acc0[d][q] = fmaf(a0[q], bv[q + d], acc0[d][q]); acc1[d][q] = fmaf(a1[q], bv[q + d], acc1[d][q]); }
    }
    asm volatile("" ::: "memory");
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int d = 0; d < 9; ++d)
#pragma unroll
    for (int q = 0; q < 8; ++q) s += acc0[d][q] + acc1[d][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}


// FFMA2 (fma.rn.f32x2) variant of the 72-accumulator tile: accumulators paired along the pixel axis; the f2 pairs of odd
// displacements are re-packed from the loaded words (7 pairs per channel step).
__device__ __forceinline__ void fma2(uint64_t& d, uint64_t a, uint64_t b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ uint64_t pack2(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
template <int MAP>
__global__ void __launch_bounds__(512, 1) k72f2(float* out, long long* clk, int iters, int nwarps) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < F1_ELEMS + F2_ELEMS; i += blockDim.x) sm[i] = (float)(i & 1023) * 1e-3f;
  __syncthreads();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp >= nwarps) return;
  const int s8 = MAP == 0 ? (lane & 3) : (lane >> 3);
  int p = (warp * 8 + (MAP == 0 ? (lane >> 2) : (lane & 7))) % 72, r, dyi;
  pair_of(p, r, dyi);
  uint64_t acc[9][4];
#pragma unroll
  for (int d = 0; d < 9; ++d)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[d][q] = 0;
  const float* f1s = sm;
  const float* f2s = sm + F1_ELEMS;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll 2
    for (int cc = 0; cc < CC; ++cc) {
      const float* ap = f1s + (cc * TH + r) * F1_P + s8 * 8;
      const float* bp = f2s + (cc * F2_H + r + dyi) * F2_P + s8 * 8;
      uint64_t a2[4], be[8], bo[7];
#pragma unroll
      for (int q = 0; q < 2; ++q) { ulonglong2 t = reinterpret_cast<const ulonglong2*>(ap)[q]; a2[2 * q] = t.x; a2[2 * q + 1] = t.y; }
#pragma unroll
      for (int q = 0; q < 4; ++q) { ulonglong2 t = reinterpret_cast<const ulonglong2*>(bp)[q]; be[2 * q] = t.x; be[2 * q + 1] = t.y; }
#pragma unroll
      for (int k = 0; k < 7; ++k) bo[k] = pack2((uint32_t)(be[k] >> 32), (uint32_t)be[k + 1]);
#pragma unroll
      for (int d = 0; d < 9; ++d)
#pragma unroll
        for (int q = 0; q < 4; ++q) fma2(acc[d][q], a2[q], (d & 1) ? bo[q + (d >> 1)] : be[q + (d >> 1)]);
    }
    asm volatile("" ::: "memory");
  }
  long long t1 = clock64();
  uint64_t s = 0;
#pragma unroll
  for (int d = 0; d < 9; ++d)
#pragma unroll
    for (int q = 0; q < 4; ++q) s ^= acc[d][q];
  reinterpret_cast<uint64_t*>(out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

// pure issue-rate loops: 72 FFMA vs 36 FFMA2 per iteration on register operands only
template <int F2>
__global__ void __launch_bounds__(512, 1) kpure(float* out, long long* clk, int iters, int nwarps) {
  const int warp = threadIdx.x >> 5;
  if (warp >= nwarps) return;
  uint64_t acc[9][4], a2[4], be[8], bo[7];
  const uint64_t* in = reinterpret_cast<const uint64_t*>(out);
  for (int q = 0; q < 4; ++q) a2[q] = in[threadIdx.x + q];
  for (int q = 0; q < 8; ++q) be[q] = in[threadIdx.x + 8 + q];
  for (int q = 0; q < 7; ++q) bo[q] = in[threadIdx.x + 16 + q];
#pragma unroll
  for (int d = 0; d < 9; ++d)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[d][q] = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters * 8; ++it) {
#pragma unroll
    for (int d = 0; d < 9; ++d)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint64_t b = (d & 1) ? bo[q + (d >> 1)] : be[q + (d >> 1)];
        if (F2) {
          fma2(acc[d][q], a2[q], b);
        } else {
          float lo = __uint_as_float((uint32_t)acc[d][q]), hi = __uint_as_float((uint32_t)(acc[d][q] >> 32));
          lo = fmaf(__uint_as_float((uint32_t)a2[q]), __uint_as_float((uint32_t)b), lo);
          hi = fmaf(__uint_as_float((uint32_t)(a2[q] >> 32)), __uint_as_float((uint32_t)(b >> 32)), hi);
          acc[d][q] = pack2(__float_as_uint(lo), __float_as_uint(hi));
        }
      }
    a2[it & 3] += 1;
  }
  long long t1 = clock64();
  uint64_t s = 0;
#pragma unroll
  for (int d = 0; d < 9; ++d)
#pragma unroll
    for (int q = 0; q < 4; ++q) s ^= acc[d][q];
  reinterpret_cast<uint64_t*>(out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

template <class K>
static void run(const char* name, K kern, int threads, int nwarps, float* out, long long* clk, double fma_per_step) {
  const int iters = 400;
  const int smem = (F1_ELEMS + F2_ELEMS) * 4;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long c = 0;
  for (int rep = 0; rep < 2; ++rep) { kern<<<148, threads, smem>>>(out, clk, iters, nwarps); cudaDeviceSynchronize(); }
  cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
  const double per = (double)c / (iters * 8.0);
  printf("%-44s %2d warps: %7.1f clk per channel step  (%.1f FMA lanes/clk/SM; %s)\n", name, nwarps, per, fma_per_step / per,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float* out; long long* clk;
  cudaMalloc(&out, 64 << 20); cudaMalloc(&clk, 8);
  for (int nw : {9, 8, 12}) {
    const double f = nw * 32 * 72.0;
    run("72 acc, lane=pair*4+strip (today), LDS.128", k72<0, 4>, 512, nw, out, clk, f);
    run("72 acc, lane=strip*8+pair,         LDS.128", k72<1, 4>, 512, nw, out, clk, f);
    run("72 acc, lane=pair*4+strip,         LDS.64 ", k72<0, 2>, 512, nw, out, clk, f);
    run("72 acc, lane=strip*8+pair,         LDS.64 ", k72<1, 2>, 512, nw, out, clk, f);
    run("72 acc, lane=strip*8+pair,         LDS.32 ", k72<1, 1>, 512, nw, out, clk, f);
  }
  for (int nw : {9, 8, 12, 16}) {
    const double f = nw * 32 * 72.0;
    run("72 acc FFMA2 (pairs along px), map 0, LDS.128", k72f2<0>, 512, nw, out, clk, f);
    run("72 acc FFMA2 (pairs along px), map 1, LDS.128", k72f2<1>, 512, nw, out, clk, f);
    run("registers only: 72 FFMA per step              ", kpure<0>, 512, nw, out, clk, f);
    run("registers only: 36 FFMA2 per step             ", kpure<1>, 512, nw, out, clk, f);
  }
  for (int nw : {4, 5, 8}) {
    const double f = nw * 32 * 144.0;
    run("144 acc (2 rows share the f2 row), map 0   ", k144<0>, 256, nw, out, clk, f);
    run("144 acc (2 rows share the f2 row), map 1   ", k144<1>, 256, nw, out, clk, f);
  }
  return 0;
}
