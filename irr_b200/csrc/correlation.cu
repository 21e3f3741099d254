// Cost volume (A1) fused with LeakyReLU (A3) and, optionally, with the bilinear warp + hard mask that feeds it (A2).
//
// Reference semantics: models/pwc_modules.py:42-62 (compute_cost_volume) == Correlation(pad 4, k 1, md 4, s 1/1) of
// models/correlation_package/correlation_cuda_kernel.cu:41-114:  out[b, (dy+4)*9+(dx+4), y, x] =
// (1/C) sum_c f1[b,c,y,x] * f2w[b,c,y+dy,x+dx], zero outside the image.
//
// B200 design (HBM-bound op sitting at the fp32-FMA ridge):
//   * one CTA = one 8 x 32 output tile of one image, 9 warps; warp w owns displacement row dy = w-4, lane = (row r of
//     the tile, 8-pixel strip s).  Each thread keeps 8 px x 9 dx = 72 fp32 accumulators, so one channel step is
//     2 + 4 LDS.128 for 72 FFMA — the register tile that keeps the FMA pipe, not the LSU, the limiter.
//   * operands are staged per 8-channel chunk in shared memory: f1 tile [8][8][36], f2 halo tile [8][16][44].  The row
//     pitches 36 / 44 (== 4, 12 mod 32 words) make every quarter-warp LDS.128 (8 rows, same strip) hit 32 distinct
//     banks.  Global reads are coalesced along W; f2's 2.5x halo re-read is served by the 126 MB L2 (the largest
//     level-4 map of cfg 3 is 28.6 MB), so DRAM traffic stays at the algorithmic B*H*W*(8C+324) bytes.
//   * in the fused-warp variant the f2 halo tile is produced by a 4-tap gather straight into shared memory; the
//     sample coordinates / weights / mask of the 16 x 40 halo pixels are computed once per tile (bit-exact recipe in
//     common.cuh) and reused for every channel.  The warped tensor never exists in HBM.
//   * epilogue: * 1/C, LeakyReLU, 16-byte stores into the caller's channel slice of the estimator input buffer.
#include "common.cuh"

namespace irr {

constexpr int TH = 8, TW = 32, MD = 4, ND = 9, PX = 8;
constexpr int F1_P = 36;                    // f1 row pitch (floats)
constexpr int F2_H = TH + 2 * MD;           // 16
constexpr int F2_WV = TW + 2 * MD;          // 40 valid halo columns
constexpr int F2_P = 44;                    // f2 row pitch (floats)
constexpr int CC = 8;                       // channels per chunk
constexpr int CORR_THREADS = 32 * ND;       // 288
constexpr int F1_ELEMS = CC * TH * F1_P;    // 2304
constexpr int F2_ELEMS = CC * F2_H * F2_P;  // 5632
constexpr int NHALO = F2_H * F2_WV;         // 640

struct HaloTap {  // 24 bytes per halo pixel (fused-warp variant)
  int off;        // clamped (y0*W + x0)
  int dxy;        // bit0: x1 = x0+1 is a distinct in-range column; bit1: same for y
  float w00, w01, w10, w11;  // mask already folded in
};

template <bool FUSED>
__global__ void __launch_bounds__(CORR_THREADS, 2)
    corr_kernel(const float* __restrict__ f1, long long f1_bs, const float* __restrict__ f2, long long f2_bs,
                const float* __restrict__ flow, long long flow_bs, float* __restrict__ out, long long out_bs,
                GridArgs g, int B, int C, int H, int W, int shift, float slope, int vec_ok) {
  extern __shared__ __align__(16) float smem[];
  float* f1s = smem;
  float* f2s = smem + F1_ELEMS;
  HaloTap* taps = reinterpret_cast<HaloTap*>(smem + F1_ELEMS + F2_ELEMS);

  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
  const int HW = H * W;
  int b2 = b + shift;
  if (b2 >= B) b2 -= B;
  const float* f1b = f1 + (size_t)b * f1_bs;
  const float* f2b = f2 + (size_t)b2 * f2_bs;

  if (FUSED) {
    const float* fl = flow + (size_t)b * flow_bs;
    for (int i = tid; i < NHALO; i += CORR_THREADS) {
      int hr = i / F2_WV, hx = i - hr * F2_WV;
      int gy = y0 - MD + hr, gx = x0 - MD + hx;
      HaloTap t;
      t.off = 0; t.dxy = 0; t.w00 = t.w01 = t.w10 = t.w11 = 0.f;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
        float u = __ldg(fl + (size_t)gy * W + gx), v = __ldg(fl + HW + (size_t)gy * W + gx);
        float ix, iy;
        sample_coords(g, u, v, gx, gy, W, H, ix, iy);
        Taps tp = make_taps(ix, iy, W, H);
        if (tp.mask != 0.f) {
          int xa = min(max(tp.x0, 0), W - 1), xb = min(max(tp.x0 + 1, 0), W - 1);
          int ya = min(max(tp.y0, 0), H - 1), yb = min(max(tp.y0 + 1, 0), H - 1);
          t.off = ya * W + xa;
          t.dxy = (xb != xa ? 1 : 0) | (yb != ya ? 2 : 0);
          // a clamped (out-of-range) tap has zero weight (make_taps), so aliasing it onto its in-range
          // neighbour's address is harmless: the four reads below are always in bounds.
          t.w00 = tp.w00; t.w01 = tp.w01; t.w10 = tp.w10; t.w11 = tp.w11;
        }
      }
      taps[i] = t;
    }
    __syncthreads();
  }

  const int dyi = tid >> 5;  // 0..8  -> dy = dyi - 4
  const int lane = tid & 31;
  const int r = lane & 7, s = lane >> 3;

  float acc[ND][PX];
#pragma unroll
  for (int d = 0; d < ND; ++d)
#pragma unroll
    for (int p = 0; p < PX; ++p) acc[d][p] = 0.f;

  for (int c0 = 0; c0 < C; c0 += CC) {
    // ---- stage f1 chunk: [CC][TH][TW], coalesced along x
    for (int i = tid; i < CC * TH * TW; i += CORR_THREADS) {
      int xx = i & (TW - 1);
      int rr = (i >> 5) & (TH - 1);
      int cc = i >> 8;
      int gy = y0 + rr, gx = x0 + xx, c = c0 + cc;
      float v = 0.f;
      if (c < C && gy < H && gx < W) v = __ldg(f1b + (size_t)c * HW + (size_t)gy * W + gx);
      f1s[(cc * TH + rr) * F1_P + xx] = v;
    }
    // ---- stage f2 halo chunk: [CC][16][40]
    for (int i = tid; i < CC * NHALO; i += CORR_THREADS) {
      int cc = i / NHALO;
      int h = i - cc * NHALO;
      int hr = h / F2_WV, hx = h - hr * F2_WV;
      int c = c0 + cc;
      float v = 0.f;
      if (FUSED) {
        if (c < C) {
          HaloTap t = taps[h];
          const float* p = f2b + (size_t)c * HW + t.off;
          int dx = t.dxy & 1, dy = (t.dxy & 2) ? W : 0;
          float a = __fmul_rn(__ldg(p), t.w00);
          a = fmaf(__ldg(p + dx), t.w01, a);
          a = fmaf(__ldg(p + dy), t.w10, a);
          a = fmaf(__ldg(p + dy + dx), t.w11, a);
          v = a;
        }
      } else {
        int gy = y0 - MD + hr, gx = x0 - MD + hx;
        if (c < C && gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(f2b + (size_t)c * HW + (size_t)gy * W + gx);
      }
      f2s[(cc * F2_H + hr) * F2_P + hx] = v;
    }
    __syncthreads();
    // ---- 8 channels x (8 px x 9 dx) FFMA per thread
#pragma unroll 2
    for (int cc = 0; cc < CC; ++cc) {
      const float4* ap = reinterpret_cast<const float4*>(f1s + (cc * TH + r) * F1_P + s * PX);
      const float4* bp = reinterpret_cast<const float4*>(f2s + (cc * F2_H + r + dyi) * F2_P + s * PX);
      float a[PX], bv[PX + 2 * MD];
      float4 t0 = ap[0], t1 = ap[1];
      a[0] = t0.x; a[1] = t0.y; a[2] = t0.z; a[3] = t0.w; a[4] = t1.x; a[5] = t1.y; a[6] = t1.z; a[7] = t1.w;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 t = bp[q];
        bv[4 * q] = t.x; bv[4 * q + 1] = t.y; bv[4 * q + 2] = t.z; bv[4 * q + 3] = t.w;
      }
#pragma unroll
      for (int d = 0; d < ND; ++d)
#pragma unroll
        for (int p = 0; p < PX; ++p) acc[d][p] = fmaf(a[p], bv[p + d], acc[d][p]);
    }
    __syncthreads();
  }

  // ---- epilogue: mean over channels (pwc_modules.py:59 / .cu:107 divide by nelems), LeakyReLU (IRR_PWC.py:94-95)
  const int gy = y0 + r;
  const int gx = x0 + s * PX;
  if (gy < H && gx < W) {
    const float fc = (float)C;
    float* op = out + (size_t)b * out_bs + (size_t)(dyi * ND) * HW + (size_t)gy * W + gx;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      float v[PX];
#pragma unroll
      for (int p = 0; p < PX; ++p) v[p] = leaky(__fdiv_rn(acc[d][p], fc), slope);
      float* q = op + (size_t)d * HW;
      if (vec_ok && gx + PX <= W) {
        reinterpret_cast<float4*>(q)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(q)[1] = make_float4(v[4], v[5], v[6], v[7]);
      } else {
#pragma unroll
        for (int p = 0; p < PX; ++p)
          if (gx + p < W) q[p] = v[p];
      }
    }
  }
}

// Generic correlation (any odd kernel_size, strides): one thread per output element.  Restates
// correlation_cuda_kernel.cu:41-114 directly on NCHW with on-the-fly zero padding; API completeness only.
__global__ void corr_generic_kernel(const float* __restrict__ in1, const float* __restrict__ in2,
                                    float* __restrict__ out, int C, int H, int W, int pad, int krad, int md, int s1,
                                    int s2, int drad, int oc, int oh, int ow, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int bx = (int)(i % ow);
  long long rr = i / ow;
  int by = (int)(rr % oh);
  rr /= oh;
  int tc = (int)(rr % oc);
  int n = (int)(rr / oc);
  int dsize = 2 * drad + 1;
  int tj = tc / dsize - drad, ti = tc % dsize - drad;
  int y1 = by * s1 + md - pad, x1 = bx * s1 + md - pad;  // un-padded coordinates
  int y2 = y1 + tj * s2, x2 = x1 + ti * s2;
  const float* a = in1 + (size_t)n * C * H * W;
  const float* bq = in2 + (size_t)n * C * H * W;
  float acc = 0.f;
  for (int j = -krad; j <= krad; ++j)
    for (int ii = -krad; ii <= krad; ++ii) {
      int ya = y1 + j, xa = x1 + ii, yb = y2 + j, xb = x2 + ii;
      if (ya < 0 || ya >= H || xa < 0 || xa >= W || yb < 0 || yb >= H || xb < 0 || xb >= W) continue;
      for (int c = 0; c < C; ++c)
        acc = fmaf(__ldg(a + ((size_t)c * H + ya) * W + xa), __ldg(bq + ((size_t)c * H + yb) * W + xb), acc);
    }
  int ks = 2 * krad + 1;
  out[i] = acc / (float)(ks * ks * C);
}

template <bool FUSED>
static int launch_corr(const char* fn, const float* f1, long long f1_bs, const float* f2, long long f2_bs,
                       const float* flow, long long flow_bs, float* out, long long out_bs, const GridArgs& g, int B,
                       int C, int H, int W, int shift, float slope, cudaStream_t st) {
  size_t smem = (size_t)(F1_ELEMS + F2_ELEMS) * sizeof(float) + (FUSED ? NHALO * sizeof(HaloTap) : 0);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(corr_kernel<FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute: %s", fn, cudaGetErrorString(e));
      return (int)e;
    }
    attr_done = true;
  }
  int vec_ok = (W % 4 == 0) && (out_bs % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, B);
  corr_kernel<FUSED><<<grid, CORR_THREADS, smem, st>>>(f1, f1_bs, f2, f2_bs, flow, flow_bs, out, out_bs, g, B, C, H, W,
                                                       shift, slope, vec_ok);
  return check_launch(fn);
}

}  // namespace irr

using namespace irr;

extern "C" {

int irr_correlation_fwd(const float* f1, long long f1_bs, const float* f2, long long f2_bs, float* out,
                        long long out_bs, int B, int C, int H, int W, int max_disp, int f2_batch_shift,
                        float leaky_slope, irr_stream_t stream) {
  const char* fn = "irr_correlation_fwd";
  IRR_REQUIRE(f1 && f2 && out, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, fn, "non-positive size");
  IRR_REQUIRE(B <= 65535 && (H + TH - 1) / TH <= 65535, fn, "size exceeds grid limits");
  IRR_REQUIRE(max_disp == MD, fn, "only max_disp == 4 is compiled (use irr_correlation_generic_fwd)");
  IRR_REQUIRE(f2_batch_shift >= 0 && f2_batch_shift < B, fn, "f2_batch_shift out of range");
  GridArgs g = make_grid_args(nullptr, nullptr, H, W, H, W, 1.f, 0);
  return launch_corr<false>(fn, f1, f1_bs, f2, f2_bs, nullptr, 0, out, out_bs, g, B, C, H, W, f2_batch_shift,
                            leaky_slope, as_stream(stream));
}

int irr_warp_correlation_fwd(const float* f1, long long f1_bs, const float* f2, long long f2_bs, const float* flow,
                             long long flow_bs, const float* lin_x, const float* lin_y, float* out, long long out_bs,
                             int B, int C, int H, int W, int H_im, int W_im, float div_flow, int max_disp,
                             int f2_batch_shift, float leaky_slope, int grid_flags, irr_stream_t stream) {
  const char* fn = "irr_warp_correlation_fwd";
  IRR_REQUIRE(f1 && f2 && flow && out, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && H_im > 0 && W_im > 0, fn, "non-positive size");
  IRR_REQUIRE(B <= 65535 && (H + TH - 1) / TH <= 65535, fn, "size exceeds grid limits");
  IRR_REQUIRE(max_disp == MD, fn, "only max_disp == 4 is compiled");
  IRR_REQUIRE(f2_batch_shift >= 0 && f2_batch_shift < B, fn, "f2_batch_shift out of range");
  GridArgs g = make_grid_args(lin_x, lin_y, H, W, H_im, W_im, div_flow, grid_flags);
  return launch_corr<true>(fn, f1, f1_bs, f2, f2_bs, flow, flow_bs, out, out_bs, g, B, C, H, W, f2_batch_shift,
                           leaky_slope, as_stream(stream));
}

int irr_correlation_generic_out_shape(int H, int W, int pad_size, int kernel_size, int max_displacement, int stride1,
                                      int stride2, int* out_c, int* out_h, int* out_w) {
  const char* fn = "irr_correlation_generic_out_shape";
  IRR_REQUIRE(H > 0 && W > 0 && pad_size >= 0 && max_displacement >= 0 && stride1 >= 1 && stride2 >= 1, fn, "sizes");
  IRR_REQUIRE(kernel_size >= 1 && (kernel_size & 1), fn, "kernel_size must be odd");
  int krad = (kernel_size - 1) / 2, border = krad + max_displacement;  // correlation_cuda.cc:23-32
  int d = (max_displacement / stride2) * 2 + 1;
  int pH = H + 2 * pad_size, pW = W + 2 * pad_size;
  int oh = (pH - 2 * border + stride1 - 1) / stride1, ow = (pW - 2 * border + stride1 - 1) / stride1;
  IRR_REQUIRE(oh > 0 && ow > 0, fn, "empty output");
  if (out_c) *out_c = d * d;
  if (out_h) *out_h = oh;
  if (out_w) *out_w = ow;
  return 0;
}

int irr_correlation_generic_fwd(const float* in1, const float* in2, float* out, int B, int C, int H, int W,
                                int pad_size, int kernel_size, int max_displacement, int stride1, int stride2,
                                irr_stream_t stream) {
  const char* fn = "irr_correlation_generic_fwd";
  IRR_REQUIRE(in1 && in2 && out, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0, fn, "non-positive size");
  int oc, oh, ow;
  int rc = irr_correlation_generic_out_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2, &oc, &oh,
                                             &ow);
  if (rc) return rc;
  long long total = (long long)B * oc * oh * ow;
  int krad = (kernel_size - 1) / 2;
  corr_generic_kernel<<<(unsigned)((total + 127) / 128), 128, 0, as_stream(stream)>>>(
      in1, in2, out, C, H, W, pad_size, krad, max_displacement, stride1, stride2, max_displacement / stride2, oc, oh, ow,
      total);
  return check_launch(fn);
}

}  // extern "C"
