// C-ABI dispatch for conv() (include/irr_b200.h): routes to the CUDA-core path (conv_simt.cu) or the tcgen05 path
// (conv_tc.cu) according to `math`.  There is no CPU or library fallback: an unsupported request is an error.
#include "common.cuh"

namespace irr {
size_t simt_packed_bytes(int Cout, int Cin, int ks);
int simt_pack(const float* w, void* out, int Cout, int Cin, int ks, cudaStream_t st);
bool simt_direct_supported(int Cout, int Cin, int ks);
int simt_conv(const float* x, long long x_bs, const void* w, const float* bias, const float* addend, long long a_bs,
              float* y, long long y_bs, int B, int Cin, int H, int W, int Cout, int ks, int stride, int dil,
              float slope, float alpha, cudaStream_t st, int pitch_in, int pitch_out);
size_t tc_packed_bytes(int Cout, int Cin, int ks, int math);
int tc_pack(const float* w, void* out, int Cout, int Cin, int ks, int math, cudaStream_t st);
int tc_conv(const float* x, long long x_bs, const void* w, const float* bias, const float* addend, long long a_bs,
            float* y, long long y_bs, int B, int Cin, int H, int W, int Cout, int ks, int stride, int dil, float slope,
            float alpha, int math, cudaStream_t st);
bool tc_supported(int Cout, int Cin, int ks, int stride, int dil);
size_t h16_packed_bytes(int Cout, int Cin, int ks);
int h16_pack(const float* w, void* out, int Cout, int Cin, int ks, cudaStream_t st);
struct HSeg {  // must match conv_tc16.cu
  int n_begin, pre;
  float slope, alpha;
  const float* addend; long long a_bs;
  float* y; long long y_bs;
};
int h16_conv(const float* x, long long x_bs, const void* w, const float* bias, const HSeg* segs, int nseg, int B, int Cin,
             int H, int W, int Cout, int ks, int stride, int dil, void* ws, size_t ws_bytes, cudaStream_t st, int pitch_in,
             int pitch_out);
size_t h16_workspace_bytes(int B, int Cin, int H, int W, int Cout, int ks, int stride, int dil);
}  // namespace irr

using namespace irr;

extern "C" {

size_t irr_conv2d_packed_bytes(int Cout, int Cin, int ksize, int math) {
  if (Cout <= 0 || Cin <= 0 || (ksize != 1 && ksize != 3)) return 0;
  if (math == IRR_MATH_FP32_SIMT) return simt_packed_bytes(Cout, Cin, ksize);
  if (math == IRR_MATH_TC_3XTF32 || math == IRR_MATH_TC_TF32) return tc_packed_bytes(Cout, Cin, ksize, math);
  if (math == IRR_MATH_TC_3XF16) return Cout <= 256 ? h16_packed_bytes(Cout, Cin, ksize) : 0;
  return 0;
}

int irr_conv2d_math_supported(int Cout, int Cin, int ksize, int stride, int dilation, int math) {
  if (Cout <= 0 || Cin <= 0 || (ksize != 1 && ksize != 3) || stride < 1 || dilation < 1) return 0;
  if (math == IRR_MATH_FP32_SIMT) return 1;
  if (math == IRR_MATH_TC_3XTF32 || math == IRR_MATH_TC_TF32 || math == IRR_MATH_TC_3XF16)
    return tc_supported(Cout, Cin, ksize, stride, dilation) ? 1 : 0;
  return 0;
}

int irr_conv2d_direct_supported(int Cout, int Cin, int ksize) { return simt_direct_supported(Cout, Cin, ksize) ? 1 : 0; }

int irr_conv2d_pack_weights(const float* w_oihw, void* w_packed, int Cout, int Cin, int ksize, int math,
                            irr_stream_t stream) {
  const char* fn = "irr_conv2d_pack_weights";
  IRR_REQUIRE(w_oihw && w_packed, fn, "null pointer");
  IRR_REQUIRE(Cout > 0 && Cin > 0, fn, "non-positive size");
  IRR_REQUIRE(ksize == 1 || ksize == 3, fn, "kernel_size must be 1 or 3");
  IRR_REQUIRE((reinterpret_cast<uintptr_t>(w_packed) & 15) == 0, fn, "w_packed must be 16-byte aligned");
  if (math == IRR_MATH_FP32_SIMT) return simt_pack(w_oihw, w_packed, Cout, Cin, ksize, as_stream(stream));
  if (math == IRR_MATH_TC_3XTF32 || math == IRR_MATH_TC_TF32)
    return tc_pack(w_oihw, w_packed, Cout, Cin, ksize, math, as_stream(stream));
  if (math == IRR_MATH_TC_3XF16) {
    IRR_REQUIRE(Cout <= 256, fn, "Cout > 256 not supported by the tcgen05 path");
    return h16_pack(w_oihw, w_packed, Cout, Cin, ksize, as_stream(stream));
  }
  return fail_arg(fn, "unknown math mode");
}

size_t irr_conv2d_workspace_bytes(int B, int Cin, int H, int W, int Cout, int ksize, int stride, int dilation, int math) {
  if (math != IRR_MATH_TC_3XF16 || B <= 0 || Cin <= 0 || H <= 0 || W <= 0 || Cout <= 0 || Cout > 256 ||
      (ksize != 1 && ksize != 3) || stride < 1 || dilation < 1)
    return 0;
  return h16_workspace_bytes(B, Cin, H, W, Cout, ksize, stride, dilation);
}

int irr_conv2d_fwd(const float* x, long long x_bs, const void* w_packed, const float* bias, const float* addend,
                   long long addend_bs, float* y, long long y_bs, int B, int Cin, int H, int W, int Cout, int ksize,
                   int stride, int dilation, float leaky_slope, float alpha, int math, int x_pitch, int y_pitch,
                   irr_stream_t stream) {
  return irr_conv2d_fwd_ws(x, x_bs, w_packed, bias, addend, addend_bs, y, y_bs, B, Cin, H, W, Cout, ksize, stride,
                           dilation, leaky_slope, alpha, math, nullptr, 0, x_pitch, y_pitch, stream);
}

int irr_conv2d_fwd_ws(const float* x, long long x_bs, const void* w_packed, const float* bias, const float* addend,
                      long long addend_bs, float* y, long long y_bs, int B, int Cin, int H, int W, int Cout, int ksize,
                      int stride, int dilation, float leaky_slope, float alpha, int math, void* workspace,
                      size_t workspace_bytes, int x_pitch, int y_pitch, irr_stream_t stream) {
  const char* fn = "irr_conv2d_fwd";
  const bool pitched = (x_pitch > 0 && x_pitch != W) ||
                       (y_pitch > 0 && y_pitch != (W + 2 * (((ksize - 1) * dilation) / 2) - dilation * (ksize - 1) - 1) / stride + 1);
  IRR_REQUIRE(!pitched || math == IRR_MATH_TC_3XF16 || (math == IRR_MATH_FP32_SIMT && simt_direct_supported(Cout, Cin, ksize)), fn,
              "row pitches are implemented by the IRR_MATH_TC_3XF16 path and the direct thin-layer kernel only");
  IRR_REQUIRE(x && w_packed && bias && y, fn, "null pointer");
  IRR_REQUIRE(B > 0 && Cin > 0 && H > 0 && W > 0 && Cout > 0, fn, "non-positive size");
  IRR_REQUIRE(ksize == 1 || ksize == 3, fn, "kernel_size must be 1 or 3");
  IRR_REQUIRE(stride >= 1 && dilation >= 1, fn, "stride/dilation must be >= 1");
  IRR_REQUIRE((reinterpret_cast<uintptr_t>(w_packed) & 15) == 0, fn, "w_packed must be 16-byte aligned");
  if (math == IRR_MATH_FP32_SIMT)
    return simt_conv(x, x_bs, w_packed, bias, addend, addend_bs, y, y_bs, B, Cin, H, W, Cout, ksize, stride, dilation,
                     leaky_slope, alpha, as_stream(stream), x_pitch, y_pitch);
  if (math == IRR_MATH_TC_3XTF32 || math == IRR_MATH_TC_TF32) {
    if (!tc_supported(Cout, Cin, ksize, stride, dilation)) {
      set_error("%s: layer shape not supported by the tcgen05 path (Cout=%d Cin=%d k=%d)", fn, Cout, Cin, ksize);
      return IRR_E_UNSUPPORTED;
    }
    return tc_conv(x, x_bs, w_packed, bias, addend, addend_bs, y, y_bs, B, Cin, H, W, Cout, ksize, stride, dilation,
                   leaky_slope, alpha, math, as_stream(stream));
  }
  if (math == IRR_MATH_TC_3XF16) {
    if (!tc_supported(Cout, Cin, ksize, stride, dilation)) {
      set_error("%s: layer shape not supported by the tcgen05 path (Cout=%d Cin=%d k=%d)", fn, Cout, Cin, ksize);
      return IRR_E_UNSUPPORTED;
    }
    HSeg sg = {0, 0, leaky_slope, alpha, addend, addend_bs, y, y_bs};
    return h16_conv(x, x_bs, w_packed, bias, &sg, 1, B, Cin, H, W, Cout, ksize, stride, dilation, workspace, workspace_bytes,
                    as_stream(stream), x_pitch, y_pitch);
  }
  return fail_arg(fn, "unknown math mode");
}

int irr_conv2d_fwd_dual(const float* x, long long x_bs, const void* w_packed, const float* bias, const float* addend,
                        long long addend_bs, float* y, long long y_bs, int B, int Cin, int H, int W, int Cout, int ksize,
                        int stride, int dilation, float leaky_slope, float alpha, int n_split, const float* addend2,
                        long long addend2_bs, float* y2, long long y2_bs, float leaky_slope2, float alpha2, int math,
                        void* workspace, size_t workspace_bytes, int x_pitch, int y_pitch, irr_stream_t stream) {
  const char* fn = "irr_conv2d_fwd_dual";
  IRR_REQUIRE(x && w_packed && bias && y && y2, fn, "null pointer");
  IRR_REQUIRE(B > 0 && Cin > 0 && H > 0 && W > 0 && Cout > 0, fn, "non-positive size");
  IRR_REQUIRE(ksize == 1 || ksize == 3, fn, "kernel_size must be 1 or 3");
  IRR_REQUIRE(stride >= 1 && dilation >= 1, fn, "stride/dilation must be >= 1");
  IRR_REQUIRE(math == IRR_MATH_TC_3XF16, fn, "dual output is implemented by the IRR_MATH_TC_3XF16 path only");
  IRR_REQUIRE(n_split > 0 && n_split < Cout && (n_split % 16) == 0, fn, "n_split must be a multiple of 16 inside (0, Cout)");
  IRR_REQUIRE((reinterpret_cast<uintptr_t>(w_packed) & 15) == 0, fn, "w_packed must be 16-byte aligned");
  if (!tc_supported(Cout, Cin, ksize, stride, dilation)) {
    set_error("%s: layer shape not supported by the tcgen05 path (Cout=%d Cin=%d k=%d)", fn, Cout, Cin, ksize);
    return IRR_E_UNSUPPORTED;
  }
  HSeg sg[2] = {{0, 0, leaky_slope, alpha, addend, addend_bs, y, y_bs},
                {n_split, 0, leaky_slope2, alpha2, addend2, addend2_bs, y2, y2_bs}};
  return h16_conv(x, x_bs, w_packed, bias, sg, 2, B, Cin, H, W, Cout, ksize, stride, dilation, workspace, workspace_bytes,
                  as_stream(stream), x_pitch, y_pitch);
}

int irr_conv2d_fwd_multi(const float* x, long long x_bs, const void* w_packed, const float* bias, int B, int Cin, int H,
                         int W, int Cout, int ksize, int stride, int dilation, const irr_conv_seg* segs, int n_segs, int math,
                         void* workspace, size_t workspace_bytes, int x_pitch, int y_pitch, irr_stream_t stream) {
  const char* fn = "irr_conv2d_fwd_multi";
  IRR_REQUIRE(x && w_packed && bias && segs, fn, "null pointer");
  IRR_REQUIRE(B > 0 && Cin > 0 && H > 0 && W > 0 && Cout > 0, fn, "non-positive size");
  IRR_REQUIRE(ksize == 1 || ksize == 3, fn, "kernel_size must be 1 or 3");
  IRR_REQUIRE(stride >= 1 && dilation >= 1, fn, "stride/dilation must be >= 1");
  IRR_REQUIRE(math == IRR_MATH_TC_3XF16, fn, "output segments are implemented by the IRR_MATH_TC_3XF16 path only");
  IRR_REQUIRE(n_segs >= 1 && n_segs <= IRR_CONV_MAX_SEGS, fn, "n_segs must be in [1, IRR_CONV_MAX_SEGS]");
  IRR_REQUIRE((reinterpret_cast<uintptr_t>(w_packed) & 15) == 0, fn, "w_packed must be 16-byte aligned");
  if (!tc_supported(Cout, Cin, ksize, stride, dilation)) {
    set_error("%s: layer shape not supported by the tcgen05 path (Cout=%d Cin=%d k=%d)", fn, Cout, Cin, ksize);
    return IRR_E_UNSUPPORTED;
  }
  HSeg sg[IRR_CONV_MAX_SEGS];
  for (int i = 0; i < n_segs; ++i) {
    IRR_REQUIRE(segs[i].y != nullptr, fn, "segment without a destination");
    sg[i].n_begin = segs[i].n_begin; sg[i].pre = segs[i].addend_pre ? 1 : 0; sg[i].slope = segs[i].leaky_slope;
    sg[i].alpha = segs[i].alpha; sg[i].addend = segs[i].addend; sg[i].a_bs = segs[i].addend_bs; sg[i].y = segs[i].y;
    sg[i].y_bs = segs[i].y_bs;
  }
  return h16_conv(x, x_bs, w_packed, bias, sg, n_segs, B, Cin, H, W, Cout, ksize, stride, dilation, workspace,
                  workspace_bytes, as_stream(stream), x_pitch, y_pitch);
}

}  // extern "C"
