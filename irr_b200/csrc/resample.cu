// Small bandwidth-bound kernels of the IRR-PWC level loop: bilinear resize (align_corners), channel scaling,
// nearest-x2 (+ bilinear fix-up), spatial-mean subtraction, channel L2 norm and the softmax(-f^2) 3x3 gather.
// All are one-pass, coalesced along W, and write into channel slices (ptr + batch stride) so no torch.cat exists.
#include "common.cuh"

namespace irr {

// ---------------------------------------------------------------------------------------------------------
// upsample2d_as (models/pwc_modules.py:65-67) — bilinear, align_corners=True; arithmetic order follows
// aten UpSampleBilinear2d.cu (rheight = (H-1)/(OH-1); h1r = rheight*h2; lambda = h1r - (int)h1r).
__global__ void resize_bilinear_ac_kernel(const float* __restrict__ x, long long x_bs, float* __restrict__ y,
                                          long long y_bs, int C, int H, int W, int OH, int OW, float rh, float rw,
                                          float s_even, float s_odd, long long total) {
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
  const float* p = x + (size_t)b * x_bs + (size_t)c * H * W;
  float a = __ldg(p + (size_t)h1 * W + w1), bb = __ldg(p + (size_t)h1 * W + w1 + w1p);
  float cc = __ldg(p + (size_t)(h1 + h1p) * W + w1), d = __ldg(p + (size_t)(h1 + h1p) * W + w1 + w1p);
  float top = __fadd_rn(__fmul_rn(w0l, a), __fmul_rn(w1l, bb));
  float bot = __fadd_rn(__fmul_rn(w0l, cc), __fmul_rn(w1l, d));
  float v = __fadd_rn(__fmul_rn(h0l, top), __fmul_rn(h1l, bot));
  float s = (c & 1) ? s_odd : s_even;
  y[(size_t)b * y_bs + ((size_t)c * OH + oy) * OW + ox] = (s == 1.0f) ? v : __fmul_rn(v, s);
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

// upsample_factor2 (models/irr_modules.py:21-27).
__global__ void upsample_nearest2x_kernel(const float* __restrict__ x, long long x_bs, float* __restrict__ y,
                                          long long y_bs, int C, int H, int W, int OH, int OW, int exact,
                                          float sh, float sw, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int ox = (int)(i % OW);
  long long r = i / OW;
  int oy = (int)(r % OH);
  r /= OH;
  int c = (int)(r % C);
  int b = (int)(r / C);
  const float* p = x + (size_t)b * x_bs + (size_t)c * H * W;
  float v;
  if (exact) {
    v = __ldg(p + (size_t)(oy >> 1) * W + (ox >> 1));
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
    float a = __ldg(p + (size_t)(h1 >> 1) * W + (w1 >> 1));
    float bb = __ldg(p + (size_t)(h1 >> 1) * W + ((w1 + w1p) >> 1));
    float cc = __ldg(p + (size_t)((h1 + h1p) >> 1) * W + (w1 >> 1));
    float d = __ldg(p + (size_t)((h1 + h1p) >> 1) * W + ((w1 + w1p) >> 1));
    float top = __fadd_rn(__fmul_rn(w0l, a), __fmul_rn(w1l, bb));
    float bot = __fadd_rn(__fmul_rn(w0l, cc), __fmul_rn(w1l, d));
    v = __fadd_rn(__fmul_rn(h0l, top), __fmul_rn(h1l, bot));
  }
  y[(size_t)b * y_bs + ((size_t)c * OH + oy) * OW + ox] = v;
}

// subtract_mean (models/irr_modules.py:59-60): one CTA per (b, c) plane.
__global__ void __launch_bounds__(512) sub_spatial_mean_kernel(const float* __restrict__ x, long long x_bs,
                                                               float* __restrict__ y, long long y_bs, int C, int HW) {
  int c = blockIdx.x % C, b = blockIdx.x / C;
  const float* p = x + (size_t)b * x_bs + (size_t)c * HW;
  float* q = y + (size_t)b * y_bs + (size_t)c * HW;
  float s = 0.f;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) s += __ldg(p + i);
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
  for (int i = threadIdx.x; i < HW; i += blockDim.x) q[i] = __ldg(p + i) - m;
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
                                     long long s_bs, float* __restrict__ out, long long o_bs, int C, int H, int W,
                                     long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int x = (int)(i % W);
  long long r = i / W;
  int yy = (int)(r % H);
  int b = (int)(r / H);
  size_t HW = (size_t)H * W;
  const float* lp = logits + (size_t)b * l_bs + (size_t)yy * W + x;
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
    for (int t = 0; t < 9; ++t) acc += __ldg(sp + (size_t)ys[t / 3] * W + xs[t % 3]) * (k[t] * inv);
    out[(size_t)b * o_bs + (size_t)c * HW + (size_t)yy * W + x] = acc;
  }
}

static inline unsigned blocks_for(long long total, int threads) { return (unsigned)((total + threads - 1) / threads); }

}  // namespace irr

using namespace irr;

extern "C" {

int irr_resize_bilinear_ac_fwd(const float* x, long long x_bs, float* y, long long y_bs, int B, int C, int H, int W,
                               int OH, int OW, float scale_even, float scale_odd, irr_stream_t stream) {
  const char* fn = "irr_resize_bilinear_ac_fwd";
  IRR_REQUIRE(x && y, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && OH > 0 && OW > 0, fn, "non-positive size");
  float rh = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.f;
  float rw = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.f;
  long long total = (long long)B * C * OH * OW;
  resize_bilinear_ac_kernel<<<blocks_for(total, 256), 256, 0, as_stream(stream)>>>(x, x_bs, y, y_bs, C, H, W, OH, OW,
                                                                                    rh, rw, scale_even, scale_odd, total);
  return check_launch(fn);
}

int irr_scale_channels_fwd(const float* x, long long x_bs, float* y, long long y_bs, int B, int C, long long HW,
                           float scale_even, float scale_odd, irr_stream_t stream) {
  const char* fn = "irr_scale_channels_fwd";
  IRR_REQUIRE(x && y, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && HW > 0, fn, "non-positive size");
  long long total = (long long)B * C * HW;
  scale_channels_kernel<<<blocks_for(total, 256), 256, 0, as_stream(stream)>>>(x, x_bs, y, y_bs, C, HW, scale_even,
                                                                                scale_odd, total);
  return check_launch(fn);
}

int irr_upsample_nearest2x_fwd(const float* x, long long x_bs, float* y, long long y_bs, int B, int C, int H, int W,
                               int OH, int OW, irr_stream_t stream) {
  const char* fn = "irr_upsample_nearest2x_fwd";
  IRR_REQUIRE(x && y, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && OH > 0 && OW > 0, fn, "non-positive size");
  int exact = (OH == 2 * H && OW == 2 * W) ? 1 : 0;
  float sh = (float)(2 * H) / (float)OH, sw = (float)(2 * W) / (float)OW;
  long long total = (long long)B * C * OH * OW;
  upsample_nearest2x_kernel<<<blocks_for(total, 256), 256, 0, as_stream(stream)>>>(x, x_bs, y, y_bs, C, H, W, OH, OW,
                                                                                    exact, sh, sw, total);
  return check_launch(fn);
}

int irr_sub_spatial_mean_fwd(const float* x, long long x_bs, float* y, long long y_bs, int B, int C, int H, int W,
                             irr_stream_t stream) {
  const char* fn = "irr_sub_spatial_mean_fwd";
  IRR_REQUIRE(x && y, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, fn, "non-positive size");
  sub_spatial_mean_kernel<<<B * C, 512, 0, as_stream(stream)>>>(x, x_bs, y, y_bs, C, H * W);
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

int irr_refine_gather_fwd(const float* logits, long long logits_bs, const float* src, long long src_bs, float* out,
                          long long out_bs, int B, int C, int H, int W, irr_stream_t stream) {
  const char* fn = "irr_refine_gather_fwd";
  IRR_REQUIRE(logits && src && out, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, fn, "non-positive size");
  long long total = (long long)B * H * W;
  refine_gather_kernel<<<blocks_for(total, 128), 128, 0, as_stream(stream)>>>(logits, logits_bs, src, src_bs, out,
                                                                               out_bs, C, H, W, total);
  return check_launch(fn);
}

}  // extern "C"
