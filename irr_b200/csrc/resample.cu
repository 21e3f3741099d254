// Small bandwidth-bound kernels of the IRR-PWC level loop: bilinear resize (align_corners), channel scaling,
// nearest-x2 (+ bilinear fix-up), spatial-mean subtraction, channel L2 norm and the softmax(-f^2) 3x3 gather.
// All are one-pass, coalesced along W, and write into channel slices (ptr + batch stride) so no torch.cat exists.
#include "common.cuh"

namespace irr {

// ---------------------------------------------------------------------------------------------------------
// upsample2d_as (models/pwc_modules.py:65-67) — bilinear, align_corners=True; arithmetic order follows
// aten UpSampleBilinear2d.cu (rheight = (H-1)/(OH-1); h1r = rheight*h2; lambda = h1r - (int)h1r).
__global__ void resize_bilinear_ac_kernel(const float* __restrict__ x, long long x_bs, float* __restrict__ y,
                                          long long y_bs, int C, int H, int W, int OH, int OW, int PI, int PO,
                                          float rh, float rw, float s_even, float s_odd, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int ox = (int)(i % OW);
  long long r = i / OW;
  int oy = (int)(r % OH);
  r /= OH;
  int c = (int)(r % C);
  int b = (int)(r / C);
  float h1r = __fmul_rn(rh, (float)oy);
  int h1 = (int)h1r;
  int h1p = (h1 < H - 1) ? 1 : 0;
  float h1l = __fsub_rn(h1r, (float)h1), h0l = __fsub_rn(1.0f, h1l);
  float w1r = __fmul_rn(rw, (float)ox);
  int w1 = (int)w1r;
  int w1p = (w1 < W - 1) ? 1 : 0;
  float w1l = __fsub_rn(w1r, (float)w1), w0l = __fsub_rn(1.0f, w1l);
  const float* p = x + (size_t)b * x_bs + (size_t)c * H * PI;   // PI / PO: row pitch of the input / output (>= W / OW)
  float a = __ldg(p + (size_t)h1 * PI + w1), bb = __ldg(p + (size_t)h1 * PI + w1 + w1p);
  float cc = __ldg(p + (size_t)(h1 + h1p) * PI + w1), d = __ldg(p + (size_t)(h1 + h1p) * PI + w1 + w1p);
  float top = __fadd_rn(__fmul_rn(w0l, a), __fmul_rn(w1l, bb));
  float bot = __fadd_rn(__fmul_rn(w0l, cc), __fmul_rn(w1l, d));
  float v = __fadd_rn(__fmul_rn(h0l, top), __fmul_rn(h1l, bot));
  float s = (c & 1) ? s_odd : s_even;
  y[(size_t)b * y_bs + ((size_t)c * OH + oy) * PO + ox] = (s == 1.0f) ? v : __fmul_rn(v, s);
}

__global__ void scale_channels_kernel(const float* __restrict__ x, long long x_bs, float* __restrict__ y,
                                      long long y_bs, int C, long long HW, float s_even, float s_odd,
                                      long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long pix = i % HW;
  long long r = i / HW;
  int c = (int)(r % C);
  int b = (int)(r / C);
  float v = __ldg(x + (size_t)b * x_bs + (size_t)c * HW + pix);
  y[(size_t)b * y_bs + (size_t)c * HW + pix] = __fmul_rn(v, (c & 1) ? s_odd : s_even);
}

__global__ void scale_channels_v4_kernel(const float* __restrict__ x, long long x_bs, float* __restrict__ y,
                                         long long y_bs, int C, long long HW4, float s_even, float s_odd, long long total4) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  long long q = i % HW4;
  long long r = i / HW4;
  int c = (int)(r % C);
  int b = (int)(r / C);
  float4 v = __ldg(reinterpret_cast<const float4*>(x + (size_t)b * x_bs + (size_t)c * HW4 * 4) + q);
  const float s = (c & 1) ? s_odd : s_even;
  v.x = __fmul_rn(v.x, s); v.y = __fmul_rn(v.y, s); v.z = __fmul_rn(v.z, s); v.w = __fmul_rn(v.w, s);
  reinterpret_cast<float4*>(y + (size_t)b * y_bs + (size_t)c * HW4 * 4)[q] = v;
}

// BASELINE config 5 ("mixed bf16 features"): y = float(bf16_rn(x)) — the feature pyramid carries bf16 VALUES (what a
// `.bfloat16()` cast before the warp / correlation produces) in the fp32 NCHW layout every consumer already reads.
__global__ void round_bf16_kernel(const float* __restrict__ x, long long x_bs, float* __restrict__ y, long long y_bs,
                                  long long CHW, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long e = i % CHW;
  long long b = i / CHW;
  float v = __ldg(x + (size_t)b * x_bs + e);
  uint32_t u = __float_as_uint(v);
  // round to nearest even on the upper 16 bits (NaN stays NaN: the quiet bit is forced, as torch's cast does)
  uint32_t r = ((u & 0x7fffffffu) > 0x7f800000u) ? ((u >> 16) | 0x0040u) : ((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
  y[(size_t)b * y_bs + e] = __uint_as_float(r << 16);
}

// The same rounding with a second destination: packed bf16 STORAGE (2 bytes per element, its own batch stride / row pitch)
// for the consumers that can read it (the correlation kernels, dtype_in = IRR_DTYPE_BF16).  y (fp32 layout) is optional.
__global__ void round_bf16_store_kernel(const float* __restrict__ x, long long x_bs, int PX, float* __restrict__ y,
                                        long long y_bs, unsigned short* __restrict__ y16, long long y16_bs, int P16, int C,
                                        int H, int W, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int w = (int)(i % W);
  long long r = i / W;
  const int h = (int)(r % H);
  r /= H;
  const int c = (int)(r % C);
  const long long b = r / C;
  const size_t e = ((size_t)c * H + h) * PX + w;
  const uint32_t u = __float_as_uint(__ldg(x + (size_t)b * x_bs + e));
  const uint32_t q = ((u & 0x7fffffffu) > 0x7f800000u) ? ((u >> 16) | 0x0040u) : ((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
  if (y != nullptr) y[(size_t)b * y_bs + e] = __uint_as_float(q << 16);
  y16[(size_t)b * y16_bs + ((size_t)c * H + h) * P16 + w] = (unsigned short)q;
}

// upsample_factor2 (models/irr_modules.py:21-27).
__global__ void upsample_nearest2x_kernel(const float* __restrict__ x, long long x_bs, float* __restrict__ y,
                                          long long y_bs, int C, int H, int W, int OH, int OW, int PI, int PO, int exact,
                                          float sh, float sw, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int ox = (int)(i % OW);
  long long r = i / OW;
  int oy = (int)(r % OH);
  r /= OH;
  int c = (int)(r % C);
  int b = (int)(r / C);
  const float* p = x + (size_t)b * x_bs + (size_t)c * H * PI;
  float v;
  if (exact) {
    v = __ldg(p + (size_t)(oy >> 1) * PI + (ox >> 1));
  } else {
    // bilinear align_corners=False over the virtual (2H x 2W) nearest-upsampled image (aten UpSample.cuh:96-130)
    int IH = 2 * H, IW = 2 * W;
    float h1r = __fsub_rn(__fmul_rn(sh, __fadd_rn((float)oy, 0.5f)), 0.5f);
    if (h1r < 0.f) h1r = 0.f;
    float w1r = __fsub_rn(__fmul_rn(sw, __fadd_rn((float)ox, 0.5f)), 0.5f);
    if (w1r < 0.f) w1r = 0.f;
    int h1 = (int)h1r, w1 = (int)w1r;
    int h1p = (h1 < IH - 1) ? 1 : 0, w1p = (w1 < IW - 1) ? 1 : 0;
    float h1l = __fsub_rn(h1r, (float)h1), h0l = __fsub_rn(1.0f, h1l);
    float w1l = __fsub_rn(w1r, (float)w1), w0l = __fsub_rn(1.0f, w1l);
    float a = __ldg(p + (size_t)(h1 >> 1) * PI + (w1 >> 1));
    float bb = __ldg(p + (size_t)(h1 >> 1) * PI + ((w1 + w1p) >> 1));
    float cc = __ldg(p + (size_t)((h1 + h1p) >> 1) * PI + (w1 >> 1));
    float d = __ldg(p + (size_t)((h1 + h1p) >> 1) * PI + ((w1 + w1p) >> 1));
    float top = __fadd_rn(__fmul_rn(w0l, a), __fmul_rn(w1l, bb));
    float bot = __fadd_rn(__fmul_rn(w0l, cc), __fmul_rn(w1l, d));
    v = __fadd_rn(__fmul_rn(h0l, top), __fmul_rn(h1l, bot));
  }
  y[(size_t)b * y_bs + ((size_t)c * OH + oy) * PO + ox] = v;
}

// subtract_mean (models/irr_modules.py:59-60): one CTA per (b, c) plane.
__global__ void __launch_bounds__(512) sub_spatial_mean_kernel(const float* __restrict__ x, long long x_bs,
                                                               float* __restrict__ y, long long y_bs, int C, int H, int W,
                                                               int P) {
  int c = blockIdx.x % C, b = blockIdx.x / C;
  const int HW = H * W;
  const float* p = x + (size_t)b * x_bs + (size_t)c * H * P;   // rows stored with pitch P >= W; the mean is over H x W
  float* q = y + (size_t)b * y_bs + (size_t)c * H * P;
  float s = 0.f;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) s += __ldg(p + (i / W) * P + (i % W));
  __shared__ float red[16];
  __shared__ float mean_s;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) mean_s = t / (float)HW;
  }
  __syncthreads();
  float m = mean_s;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const int o = (i / W) * P + (i % W);
    q[o] = __ldg(p + o) - m;
  }
}

// Evaluation metrics of the reference's eval-mode losses, one CTA per image, deterministic:
//   sums[b] = { S epe*valid, S valid, S outlier, S pred*true, S pred, S true, 0, 0 }   (float64)
// epe = ||target - flow||_2 per pixel (losses.py:8-10); valid == nullptr -> all ones (losses.py:635);
// outlier = (epe*valid > 3) * (epe*valid / (||target||_2 + 1e-8) > 0.05) * valid (losses.py:693-697);
// pred = round(sigmoid(occ logits)) (losses.py:636), true = target_occ; F1 is finished on the host (losses.py:27-37).
__global__ void __launch_bounds__(1024) eval_metrics_kernel(const float* __restrict__ flow, long long flow_bs,
                                                            const float* __restrict__ target, long long target_bs,
                                                            const float* __restrict__ valid, long long valid_bs,
                                                            const float* __restrict__ occ, long long occ_bs,
                                                            const float* __restrict__ tocc, long long tocc_bs,
                                                            double* __restrict__ sums, int HW) {
  const int b = blockIdx.x;
  const float* fu = flow + (size_t)b * flow_bs;
  const float* tu = target + (size_t)b * target_bs;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const float t0 = __ldg(tu + i), t1 = __ldg(tu + HW + i);
    const float d0 = __fsub_rn(t0, __ldg(fu + i)), d1 = __fsub_rn(t1, __ldg(fu + HW + i));
    const float epe = sqrtf(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)));
    const float v = valid ? __ldg(valid + (size_t)b * valid_bs + i) : 1.0f;
    const float ev = __fmul_rn(epe, v);
    const float mag = __fadd_rn(sqrtf(__fadd_rn(__fmul_rn(t0, t0), __fmul_rn(t1, t1))), 1e-8f);
    acc[0] += (double)ev;
    acc[1] += (double)v;
    acc[2] += (double)(((ev > 3.0f) && (__fdiv_rn(ev, mag) > 0.05f)) ? v : 0.0f);
    if (occ) {
      const float x = __ldg(occ + (size_t)b * occ_bs + i);
      const float sgm = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
      const float pr = sgm > 0.5f ? 1.0f : 0.0f;  // torch.round is half-to-even: 0.5 -> 0
      const float tr = tocc ? __ldg(tocc + (size_t)b * tocc_bs + i) : 0.0f;
      acc[3] += (double)(pr * tr);
      acc[4] += (double)pr;
      acc[5] += (double)tr;
    }
  }
  __shared__ double red[6][32];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double t = acc[k];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = t;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double t = (threadIdx.x < (blockDim.x >> 5)) ? red[k][threadIdx.x] : 0.0;
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (threadIdx.x == 0) sums[(size_t)b * 8 + k] = t;
    }
    if (threadIdx.x == 0) { sums[(size_t)b * 8 + 6] = 0.0; sums[(size_t)b * 8 + 7] = 0.0; }
  }
}

__global__ void channel_l2norm_kernel(const float* __restrict__ x, long long x_bs, float* __restrict__ y,
                                      long long y_bs, int C, long long HW, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long pix = i % HW;
  int b = (int)(i / HW);
  const float* p = x + (size_t)b * x_bs + pix;
  float s = 0.f;
  for (int c = 0; c < C; ++c) {
    float v = __ldg(p + (size_t)c * HW);
    s = fmaf(v, v, s);
  }
  y[(size_t)b * y_bs + pix] = sqrtf(s);
}

// RefineFlow / RefineOcc tail (models/irr_modules.py:89-104,130-138).
__global__ void refine_gather_kernel(const float* __restrict__ logits, long long l_bs, const float* __restrict__ src,
                                     long long s_bs, float* __restrict__ out, long long o_bs, int C, int H, int W, int P,
                                     long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int x = (int)(i % W);
  long long r = i / W;
  int yy = (int)(r % H);
  int b = (int)(r / H);
  size_t HW = (size_t)H * P;   // channel stride (rows stored with pitch P >= W)
  const float* lp = logits + (size_t)b * l_bs + (size_t)yy * P + x;
  float k[9];
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    float f = __ldg(lp + t * HW);
    k[t] = -__fmul_rn(f, f);
    mx = fmaxf(mx, k[t]);
  }
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    k[t] = expf(k[t] - mx);
    sum += k[t];
  }
  float inv = 1.0f / sum;
  int ys[3] = {max(yy - 1, 0), yy, min(yy + 1, H - 1)};
  int xs[3] = {max(x - 1, 0), x, min(x + 1, W - 1)};
  for (int c = 0; c < C; ++c) {
    const float* sp = src + (size_t)b * s_bs + (size_t)c * HW;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc += __ldg(sp + (size_t)ys[t / 3] * P + xs[t % 3]) * (k[t] * inv);
    out[(size_t)b * o_bs + (size_t)c * HW + (size_t)yy * P + x] = acc;
  }
}

static inline unsigned blocks_for(long long total, int threads) { return (unsigned)((total + threads - 1) / threads); }

}  // namespace irr

using namespace irr;

extern "C" {

int irr_resize_bilinear_ac_fwd(const float* x, long long x_bs, float* y, long long y_bs, int B, int C, int H, int W,
                               int OH, int OW, float scale_even, float scale_odd, int x_pitch, int y_pitch,
                               irr_stream_t stream) {
  const char* fn = "irr_resize_bilinear_ac_fwd";
  IRR_REQUIRE(x && y, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && OH > 0 && OW > 0, fn, "non-positive size");
  const int PI = x_pitch > 0 ? x_pitch : W, PO = y_pitch > 0 ? y_pitch : OW;
  IRR_REQUIRE(PI >= W && PO >= OW, fn, "row pitch smaller than the width");
  float rh = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.f;
  float rw = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.f;
  long long total = (long long)B * C * OH * OW;
  resize_bilinear_ac_kernel<<<blocks_for(total, 256), 256, 0, as_stream(stream)>>>(x, x_bs, y, y_bs, C, H, W, OH, OW, PI, PO,
                                                                                    rh, rw, scale_even, scale_odd, total);
  return check_launch(fn);
}

int irr_scale_channels_fwd(const float* x, long long x_bs, float* y, long long y_bs, int B, int C, long long HW,
                           float scale_even, float scale_odd, irr_stream_t stream) {
  const char* fn = "irr_scale_channels_fwd";
  IRR_REQUIRE(x && y, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && HW > 0, fn, "non-positive size");
  long long total = (long long)B * C * HW;
  // 16-byte path: the concat-slice copies of the level loop (113 channels at 109 x 256 ...) are pure streams
  if ((HW % 4) == 0 && (x_bs % 4) == 0 && (y_bs % 4) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    scale_channels_v4_kernel<<<blocks_for(total / 4, 256), 256, 0, as_stream(stream)>>>(x, x_bs, y, y_bs, C, HW / 4, scale_even,
                                                                                         scale_odd, total / 4);
    return check_launch(fn);
  }
  scale_channels_kernel<<<blocks_for(total, 256), 256, 0, as_stream(stream)>>>(x, x_bs, y, y_bs, C, HW, scale_even,
                                                                                scale_odd, total);
  return check_launch(fn);
}

int irr_round_bf16_fwd(const float* x, long long x_bs, float* y, long long y_bs, int B, int C, long long HW,
                       irr_stream_t stream) {
  const char* fn = "irr_round_bf16_fwd";
  IRR_REQUIRE(x && y, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && HW > 0, fn, "non-positive size");
  long long total = (long long)B * C * HW;
  round_bf16_kernel<<<blocks_for(total, 256), 256, 0, as_stream(stream)>>>(x, x_bs, y, y_bs, (long long)C * HW, total);
  return check_launch(fn);
}

int irr_round_bf16_store_fwd(const float* x, long long x_bs, int x_pitch, float* y, long long y_bs, void* y16,
                             long long y16_bs, int y16_pitch, int B, int C, int H, int W, irr_stream_t stream) {
  const char* fn = "irr_round_bf16_store_fwd";
  IRR_REQUIRE(x && y16, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, fn, "non-positive size");
  const int PX = x_pitch > 0 ? x_pitch : W, P16 = y16_pitch > 0 ? y16_pitch : W;
  IRR_REQUIRE(PX >= W && P16 >= W, fn, "row pitch smaller than the width");
  long long total = (long long)B * C * H * W;
  round_bf16_store_kernel<<<blocks_for(total, 256), 256, 0, as_stream(stream)>>>(
      x, x_bs, PX, y, y_bs, static_cast<unsigned short*>(y16), y16_bs, P16, C, H, W, total);
  return check_launch(fn);
}

int irr_upsample_nearest2x_fwd(const float* x, long long x_bs, float* y, long long y_bs, int B, int C, int H, int W,
                               int OH, int OW, int x_pitch, int y_pitch, irr_stream_t stream) {
  const char* fn = "irr_upsample_nearest2x_fwd";
  IRR_REQUIRE(x && y, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && OH > 0 && OW > 0, fn, "non-positive size");
  const int PI = x_pitch > 0 ? x_pitch : W, PO = y_pitch > 0 ? y_pitch : OW;
  IRR_REQUIRE(PI >= W && PO >= OW, fn, "row pitch smaller than the width");
  int exact = (OH == 2 * H && OW == 2 * W) ? 1 : 0;
  float sh = (float)(2 * H) / (float)OH, sw = (float)(2 * W) / (float)OW;
  long long total = (long long)B * C * OH * OW;
  upsample_nearest2x_kernel<<<blocks_for(total, 256), 256, 0, as_stream(stream)>>>(x, x_bs, y, y_bs, C, H, W, OH, OW, PI, PO,
                                                                                    exact, sh, sw, total);
  return check_launch(fn);
}

int irr_sub_spatial_mean_fwd(const float* x, long long x_bs, float* y, long long y_bs, int B, int C, int H, int W,
                             int pitch, irr_stream_t stream) {
  const char* fn = "irr_sub_spatial_mean_fwd";
  IRR_REQUIRE(x && y, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, fn, "non-positive size");
  const int P = pitch > 0 ? pitch : W;
  IRR_REQUIRE(P >= W, fn, "row pitch smaller than the width");
  sub_spatial_mean_kernel<<<B * C, 512, 0, as_stream(stream)>>>(x, x_bs, y, y_bs, C, H, W, P);
  return check_launch(fn);
}

int irr_channel_l2norm_fwd(const float* x, long long x_bs, float* y, long long y_bs, int B, int C, long long HW,
                           irr_stream_t stream) {
  const char* fn = "irr_channel_l2norm_fwd";
  IRR_REQUIRE(x && y, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && HW > 0, fn, "non-positive size");
  long long total = (long long)B * HW;
  channel_l2norm_kernel<<<blocks_for(total, 256), 256, 0, as_stream(stream)>>>(x, x_bs, y, y_bs, C, HW, total);
  return check_launch(fn);
}

int irr_eval_metrics_fwd(const float* flow, long long flow_bs, const float* target, long long target_bs,
                         const float* valid, long long valid_bs, const float* occ_logits, long long occ_bs,
                         const float* target_occ, long long tocc_bs, double* sums, int B, int H, int W,
                         irr_stream_t stream) {
  const char* fn = "irr_eval_metrics_fwd";
  IRR_REQUIRE(flow && target && sums, fn, "null pointer");
  IRR_REQUIRE(B > 0 && H > 0 && W > 0 && (long long)H * W < 0x7fffffffLL, fn, "bad size");
  IRR_REQUIRE(!(target_occ && !occ_logits), fn, "target_occ without occ_logits");
  eval_metrics_kernel<<<B, 1024, 0, as_stream(stream)>>>(flow, flow_bs, target, target_bs, valid, valid_bs, occ_logits,
                                                         occ_bs, target_occ, tocc_bs, sums, H * W);
  return check_launch(fn);
}

int irr_refine_gather_fwd(const float* logits, long long logits_bs, const float* src, long long src_bs, float* out,
                          long long out_bs, int B, int C, int H, int W, int pitch, irr_stream_t stream) {
  const char* fn = "irr_refine_gather_fwd";
  IRR_REQUIRE(logits && src && out, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, fn, "non-positive size");
  const int P = pitch > 0 ? pitch : W;
  IRR_REQUIRE(P >= W, fn, "row pitch smaller than the width");
  long long total = (long long)B * H * W;
  refine_gather_kernel<<<blocks_for(total, 128), 128, 0, as_stream(stream)>>>(logits, logits_bs, src, src_bs, out,
                                                                               out_bs, C, H, W, P, total);
  return check_launch(fn);
}

}  // extern "C"
