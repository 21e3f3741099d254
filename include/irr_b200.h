/* irr_b200 — C ABI of the B200-native IRR-PWC inference hot path (sm_100a).
 *
 * This is the drop-in boundary.  It replaces
 *   (1) the reference's pybind11 extension `correlation_cuda`
 *         int forward(Tensor& input1, Tensor& input2, Tensor& rInput1, Tensor& rInput2, Tensor& output,
 *                     int pad_size, int kernel_size, int max_displacement, int stride1, int stride2,
 *                     int corr_type_multiply)            — models/correlation_package/correlation_cuda.cc:8-14,165-168
 *   (2) the ATen library ops the PWC models actually execute on the path (conv2d+leaky_relu, grid_sampler_2d,
 *       upsample_bilinear2d, the 81-iteration slice/mul/mean loop of compute_cost_volume, softmax/unfold of
 *       RefineFlow/RefineOcc) — call sites cited per function below.
 *
 * Conventions (all entry points):
 *   - extern "C", plain pointers and sizes, no C++/torch types.  `irr_stream_t` is a cudaStream_t/CUstream.
 *   - every tensor is fp32, NCHW, W-contiguous; a tensor argument is (pointer, batch_stride) where the pointer
 *     already points at the first channel of a channel-slice of a possibly larger buffer and `*_bs` is the
 *     distance in ELEMENTS between consecutive batch items (= C_total*H*W of the owning buffer).  This is how the
 *     reference's torch.cat() concatenations are eliminated: producers write straight into channel slices.
 *   - device = the caller's current CUDA device; pointers are device pointers owned by the caller (PyTorch's
 *     caching allocator).  The library never allocates, frees or synchronises; launches are asynchronous on
 *     `stream`, so calls are capturable into CUDA graphs.
 *   - return 0 on success; <0 = argument error (IRR_E_*), >0 = cudaError_t from the launch.  Never throws.
 *     `irr_last_error()` returns a thread-local message.  (The reference printf()s and returns 0/1, turned into
 *     AT_ERROR -> RuntimeError by correlation_cuda.cc:78-80; the Python host in irr_b200/_lib.py raises
 *     RuntimeError on any non-zero return to keep that behaviour.)
 *   - re-entrant per stream; no global mutable state except the error string and one-time function attributes.
 *   - ROW PITCH (ABI 2).  The spatial entry points take `*_pitch` arguments: the distance in elements between two rows
 *     of a plane; channel stride = H * pitch; 0 means "dense" (= W).  A pitch that is a multiple of 4 makes every row
 *     16-byte aligned whatever the width, which is what TMA needs: the host side stores KITTI's 621 / 311 / 78 / 39-wide
 *     levels with pitch 624 / 312 / 80 / 40 (the reference handles odd sizes natively, models/irr_modules.py:21-27), the
 *     tensor maps zero-fill columns >= W, and nobody reads or relies on the pad columns.  The flat kernels
 *     (irr_scale_channels_fwd, irr_round_bf16_fwd, irr_channel_l2norm_fwd) take HW = H * pitch for such tensors.
 */
#ifndef IRR_B200_H_
#define IRR_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* irr_stream_t;

#define IRR_ABI_VERSION 2   /* 2: row pitches (below), workspaces for the correlation, output segments for the conv */

#define IRR_E_ARG (-1)       /* null pointer / non-positive size / unsupported parameter */
#define IRR_E_ALIGN (-2)     /* pointer alignment requirement violated */
#define IRR_E_UNSUPPORTED (-3)

/* grid flags for the warp family */
#define IRR_GRID_TRUE_DIV 0   /* flow*2/(dim-1)/div_flow with IEEE divisions: what the reference computes on the CPU */
#define IRR_GRID_RECIP_MUL 1  /* a*(1/b): what torch's CUDA `tensor / python_scalar` computes */

/* conv math modes */
#define IRR_MATH_FP32_SIMT 0  /* CUDA-core FFMA implicit GEMM, fp32 accumulate */
#define IRR_MATH_TC_3XTF32 1  /* tcgen05 kind::tf32, hi/lo split (3 MMAs) — fp32-grade accuracy */
#define IRR_MATH_TC_TF32 2    /* tcgen05 kind::tf32 single pass — ~1e-3 relative, opt-in */
#define IRR_MATH_TC_3XF16 3   /* tcgen05 kind::f16, hi/lo split (3 MMAs at twice the tf32 rate), activations staged
                                 through shared memory by TMA — fp32-grade accuracy for |activation| < 65504 */

int irr_abi_version(void);
const char* irr_last_error(void);
/* sm count / compute capability of the current device. */
int irr_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* A1 + A3 — cost volume of models/pwc_modules.py:42-62 (== Correlation(pad=md, k=1, md, 1, 1) of
 * correlation_package/correlation.py:47-61 -> correlation_cuda_kernel.cu:41-114):
 *   out[b, (dy+md)*(2md+1)+(dx+md), y, x] = act( (1/C) * sum_c f1[b,c,y,x] * f2[b2,c,y+dy,x+dx] ),  zero outside,
 *   b2 = (b + f2_batch_shift) mod B,  act(v) = v>0 ? v : leaky_slope*v  (IRR_PWC.py:94-95; pass 1.0f for none).
 * Only max_disp == 4 is compiled (the only value any PWC model uses, IRR_PWC.py:19). */
int irr_correlation_fwd(const float* f1, long long f1_bs, const float* f2, long long f2_bs, float* out,
                        long long out_bs, int B, int C, int H, int W, int max_disp, int f2_batch_shift,
                        float leaky_slope, irr_stream_t stream);

/* A2 fused into A1 — out = act(cost_volume(f1, mask*warp(f2, flow))) without materialising the warped tensor
 * (IRR_PWC.py:86-95).  warp = WarpingLayer.forward of models/pwc_modules.py:119-133: bilinear grid_sample
 * (align_corners=True, zeros padding) at  lin_x[x] + flow_u*2/max(W_im-1,1)/div_flow  (likewise y), times the hard
 * mask (sum of in-bounds bilinear weights >= 1.0f), with PyTorch's grid_sampler_2d weight arithmetic reproduced
 * op for op so the mask is bit-identical.  lin_x (W floats) / lin_y (H floats) are the host torch.linspace(-1,1,n)
 * vectors the reference uploads (pwc_modules.py:108-111); NULL => computed in-kernel as -1 + i*(2/(n-1)).
 * flow is B x 2 x H x W (channel 0 = u). */
int irr_warp_correlation_fwd(const float* f1, long long f1_bs, const float* f2, long long f2_bs, const float* flow,
                             long long flow_bs, const float* lin_x, const float* lin_y, float* out, long long out_bs,
                             int B, int C, int H, int W, int H_im, int W_im, float div_flow, int max_disp,
                             int f2_batch_shift, float leaky_slope, int grid_flags, irr_stream_t stream);

/* The two above with an optional scratch buffer (flow == NULL: no warp, the plain cost volume), used for two things:
 *  - a launch with far fewer 8x32 tiles than the GPU has SMs (the 7x16 ... 14x32 pyramid levels: 16-32 tiles, up to 25
 *    serial 8-channel chunks each) deals the channel chunks of a tile to several CTAs; partial sums go to the workspace,
 *    a second launch adds them in a fixed order, scales by 1/C and applies the activation (deterministic);
 *  - fused launches compute the bilinear taps / hard mask of every pixel ONCE in a small pre-pass (20 bytes per pixel in
 *    the workspace) instead of per tile-halo position on the correlation kernel's compute warps.
 * workspace may be NULL (neither happens; results are the same up to the association of the channel sum);
 * irr_correlation_workspace_bytes(fused = 0/1) returns the size that allows both (0 = nothing to gain); the buffer is
 * caller-owned device memory, 256-byte aligned, private to the call until it completes on `stream`. */
size_t irr_correlation_workspace_bytes(int B, int C, int H, int W, int fused);
int irr_warp_correlation_fwd_ws(const float* f1, long long f1_bs, const float* f2, long long f2_bs, const float* flow,
                                long long flow_bs, const float* lin_x, const float* lin_y, float* out, long long out_bs,
                                int B, int C, int H, int W, int H_im, int W_im, float div_flow, int max_disp,
                                int f2_batch_shift, float leaky_slope, int grid_flags, void* workspace,
                                size_t workspace_bytes, int pitch, irr_stream_t stream);

/* The same with a choice of INPUT STORAGE for f1 / f2 (SURVEY.md §8(b) sketch `dtype_in`; BASELINE configs[4] "mixed bf16
 * features"): dtype_in = IRR_DTYPE_BF16 reads packed bf16 (2 bytes per element: half the f1 / f2 bytes of the launch);
 * f1_bs / f2_bs / in_pitch are then in bf16 ELEMENTS and must be multiples of 8 with 16-byte aligned bases (TMA rows of
 * whole 16-byte units; IRR_E_ALIGN otherwise — there is no unaligned bf16 path).  flow / out stay fp32 with row pitch
 * `pitch`.  The arithmetic is the fp32 kernel's on the converted values: results are bit-identical to
 * irr_warp_correlation_fwd_ws on the same values held in fp32.  The reference has no such mode
 * (correlation_cuda_kernel.cu:352 dispatches float / double / half as the tensor comes). */
#define IRR_DTYPE_F32 0
#define IRR_DTYPE_BF16 1
int irr_warp_correlation_fwd_dt(const void* f1, long long f1_bs, const void* f2, long long f2_bs, int dtype_in, int in_pitch,
                                const float* flow, long long flow_bs, const float* lin_x, const float* lin_y, float* out,
                                long long out_bs, int B, int C, int H, int W, int H_im, int W_im, float div_flow, int max_disp,
                                int f2_batch_shift, float leaky_slope, int grid_flags, void* workspace,
                                size_t workspace_bytes, int pitch, irr_stream_t stream);

/* A2 standalone — out[b] = mask*warp(x[(b+x_batch_shift) mod B], flow[b]);  if minuend != NULL:
 * out = minuend - mask*warp(...)   (IRR_PWC.py:132-133,144-145 feed `a - warp(b)` to the refinement nets).
 * mask_out (optional, B x H x W floats 0/1) receives the validity mask. */
int irr_warp_fwd(const float* x, long long x_bs, const float* flow, long long flow_bs, const float* lin_x,
                 const float* lin_y, const float* minuend, long long minuend_bs, float* out, long long out_bs,
                 float* mask_out, int B, int C, int H, int W, int H_im, int W_im, float div_flow, int x_batch_shift,
                 int grid_flags, int pitch, irr_stream_t stream);

/* Generic Correlation (kernel_size odd >= 1, stride1, stride2, pad_size) for API completeness of
 * correlation_package/correlation.py:47-61; output shape per correlation_cuda.cc:23-32.  corr_type_multiply is
 * ignored exactly like the reference.  Simple kernel, not tuned (no PWC model uses k>1 or strides>1). */
int irr_correlation_generic_fwd(const float* in1, const float* in2, float* out, int B, int C, int H, int W,
                                int pad_size, int kernel_size, int max_displacement, int stride1, int stride2,
                                irr_stream_t stream);
int irr_correlation_generic_out_shape(int H, int W, int pad_size, int kernel_size, int max_displacement, int stride1,
                                      int stride2, int* out_c, int* out_h, int* out_w);

/* A4-A7, A9, A10 — conv() of models/pwc_modules.py:8-19 (Conv2d k in {1,3}, padding=((k-1)*dilation)//2, bias,
 * optional LeakyReLU) as an implicit GEMM, with the reference's torch.cat / residual adds folded in:
 *   y = addend + alpha * act(conv(x, w) + bias)      (addend may be NULL; alpha = 1 for the plain case)
 * x: B x Cin x H x W slice, y: B x Cout x Ho x Wo slice, Ho = floor((H + 2*pad - dil*(k-1) - 1)/stride) + 1.
 * w_packed: produced by irr_conv2d_pack_weights for the same `math`. */
size_t irr_conv2d_packed_bytes(int Cout, int Cin, int ksize, int math);
/* 1 if `math` can run this layer shape, 0 otherwise (the CUDA-core mode runs every k in {1,3} shape). */
int irr_conv2d_math_supported(int Cout, int Cin, int ksize, int stride, int dilation, int math);
/* 1 when IRR_MATH_FP32_SIMT runs this layer on the DIRECT thin-layer kernel (K = Cin*k*k <= 32, Cout <= 16: the first
 * pyramid layer 3 -> 16 and the 16 -> 3 1x1 convs — pure HBM streams, one thread per output pixel, fp32 FMAs); that
 * kernel honours row pitches, and the host side routes these layers to it whatever the default math mode is. */
int irr_conv2d_direct_supported(int Cout, int Cin, int ksize);
int irr_conv2d_pack_weights(const float* w_oihw, void* w_packed, int Cout, int Cin, int ksize, int math,
                            irr_stream_t stream);
int irr_conv2d_fwd(const float* x, long long x_bs, const void* w_packed, const float* bias, const float* addend,
                   long long addend_bs, float* y, long long y_bs, int B, int Cin, int H, int W, int Cout, int ksize,
                   int stride, int dilation, float leaky_slope, float alpha, int math, int x_pitch, int y_pitch,
                   irr_stream_t stream);

/* Same, with an optional scratch buffer that lets the tensor-core path split the K loop of a layer over several CTAs
 * when the layer has far fewer output tiles than the GPU has SMs (the 7x16 ... 14x32 pyramid levels): partial sums go
 * to `workspace`, a second launch adds them in a fixed order and applies the epilogue (deterministic).  workspace may
 * be NULL (never split).  irr_conv2d_workspace_bytes returns the size that allows every split this shape may use
 * (0 = never split); the buffer is caller-owned device memory, 16-byte aligned, private to the call until it
 * completes on `stream`. */
size_t irr_conv2d_workspace_bytes(int B, int Cin, int H, int W, int Cout, int ksize, int stride, int dilation, int math);
int irr_conv2d_fwd_ws(const float* x, long long x_bs, const void* w_packed, const float* bias, const float* addend,
                      long long addend_bs, float* y, long long y_bs, int B, int Cin, int H, int W, int Cout, int ksize,
                      int stride, int dilation, float leaky_slope, float alpha, int math, void* workspace,
                      size_t workspace_bytes, int x_pitch, int y_pitch, irr_stream_t stream);

/* Two layers that read the same input in one pass (IRR_MATH_TC_3XF16 only): output channels [0, n_split) get the
 * first epilogue and go to y, channels [n_split, Cout) get (addend2, alpha2, leaky_slope2) and go to y2 (a
 * B x (Cout - n_split) x Ho x Wo slice).  n_split must be a multiple of 16.  Used by the dense estimators
 * (models/pwc_modules.py:163-170): conv_last(cat[conv5(x4), x4]) = W_last[:, :32] * conv5(x4) + W_last[:, 32:] * x4,
 * so the 531/530-channel part of conv_last rides along with conv5 as extra output columns and only a 32-channel
 * conv remains. */
int irr_conv2d_fwd_dual(const float* x, long long x_bs, const void* w_packed, const float* bias, const float* addend,
                        long long addend_bs, float* y, long long y_bs, int B, int Cin, int H, int W, int Cout, int ksize,
                        int stride, int dilation, float leaky_slope, float alpha, int n_split, const float* addend2,
                        long long addend2_bs, float* y2, long long y2_bs, float leaky_slope2, float alpha2, int math,
                        void* workspace, size_t workspace_bytes, int x_pitch, int y_pitch, irr_stream_t stream);

/* The general form of the two above (IRR_MATH_TC_3XF16 only): the output channels are cut into up to IRR_CONV_MAX_SEGS
 * consecutive segments, segment i = channels [n_begin_i, n_begin_{i+1}) (n_begin_0 = 0, every n_begin a multiple of 16),
 * each with its own destination slice and epilogue
 *     y_i = addend_i + alpha_i * act_i(conv + bias)                  (addend_pre == 0)
 *     y_i = alpha_i * act_i(conv + bias + addend_i)                  (addend_pre != 0: a partial sum of the SAME layer
 *                                                                     computed by an earlier pass over other input channels)
 * This is how the dense estimators (models/pwc_modules.py:153-170) run conv4, the part of conv5 and the part of
 * conv_last that read conv4's input as ONE pass: the thin layers' partial sums ride along as extra output columns of a
 * layer whose activations are being staged anyway.  `segs` is a host array read during the call. */
#define IRR_CONV_MAX_SEGS 4
typedef struct irr_conv_seg {
  int n_begin;          /* first output channel of the segment */
  int addend_pre;       /* 0: addend is added after the activation; 1: before it */
  float leaky_slope;    /* 1.0f = no activation */
  float alpha;
  const float* addend;  /* B x n_i x Ho x Wo slice or NULL */
  long long addend_bs;
  float* y;             /* destination slice: channel n of the layer goes to y[:, n - n_begin] */
  long long y_bs;
} irr_conv_seg;
int irr_conv2d_fwd_multi(const float* x, long long x_bs, const void* w_packed, const float* bias, int B, int Cin, int H,
                         int W, int Cout, int ksize, int stride, int dilation, const irr_conv_seg* segs, int n_segs, int math,
                         void* workspace, size_t workspace_bytes, int x_pitch, int y_pitch, irr_stream_t stream);

/* A8 — upsample2d_as (models/pwc_modules.py:65-67): bilinear, align_corners=True, any in/out size, fused with an
 * optional per-channel-parity scale (even channels * scale_even, odd * scale_odd): rescale_flow of
 * pwc_modules.py:70-82 applied to the resized flow, or the final *(1/div_flow) of IRR_PWC.py:176. */
int irr_resize_bilinear_ac_fwd(const float* x, long long x_bs, float* y, long long y_bs, int B, int C, int H, int W,
                               int OH, int OW, float scale_even, float scale_odd, int x_pitch, int y_pitch,
                               irr_stream_t stream);

/* A8 — rescale_flow value semantics / channel-slice copy: y[b,c] = x[b,c] * (c even ? scale_even : scale_odd). */
int irr_scale_channels_fwd(const float* x, long long x_bs, float* y, long long y_bs, int B, int C, long long HW,
                           float scale_even, float scale_odd, irr_stream_t stream);

/* BASELINE config 5 (mixed bf16 features; no reference counterpart — the reference is fp32 only):
 * y = float(bfloat16(x)), round-to-nearest-even, on a channel-slice view.  In place (y == x) is allowed. */
int irr_round_bf16_fwd(const float* x, long long x_bs, float* y, long long y_bs, int B, int C, long long HW,
                       irr_stream_t stream);

/* The same rounding with a second, PACKED destination: y16[b,c,h,w] = bf16_rn(x[b,c,h,w]) as 2-byte elements (batch stride
 * y16_bs and row pitch y16_pitch in bf16 elements) for irr_warp_correlation_fwd_dt; y (fp32 layout, pitch x_pitch, may be x
 * itself or NULL) receives float(bf16_rn(x)) as irr_round_bf16_fwd does. */
int irr_round_bf16_store_fwd(const float* x, long long x_bs, int x_pitch, float* y, long long y_bs, void* y16,
                             long long y16_bs, int y16_pitch, int B, int C, int H, int W, irr_stream_t stream);

/* A10 — upsample_factor2 (models/irr_modules.py:21-27): nearest x2, then (only if (OH,OW) != (2H,2W)) bilinear
 * align_corners=False resize to OH x OW. */
int irr_upsample_nearest2x_fwd(const float* x, long long x_bs, float* y, long long y_bs, int B, int C, int H, int W,
                               int OH, int OW, int x_pitch, int y_pitch, irr_stream_t stream);

/* A9 head — RefineFlow input preparation (models/irr_modules.py:59-60,85-88):
 *   y[b,c] = x[b,c] - mean_w(mean_h(x[b,c]))  for C channels.  One CTA per (b,c). */
int irr_sub_spatial_mean_fwd(const float* x, long long x_bs, float* y, long long y_bs, int B, int C, int H, int W,
                             int pitch, irr_stream_t stream);
/*   y[b,0] = sqrt(sum_c x[b,c]^2)   (torch.norm(p=2, dim=1), irr_modules.py:86). */
int irr_channel_l2norm_fwd(const float* x, long long x_bs, float* y, long long y_bs, int B, int C, long long HW,
                           irr_stream_t stream);

/* A9 tail — softmax(-logits^2) over the 9 taps, applied to the replicate-padded 3x3 neighbourhood of every
 * channel of src (models/irr_modules.py:89-104, 130-138).  logits: B x 9 x H x W, src/out: B x C x H x W. */
int irr_refine_gather_fwd(const float* logits, long long logits_bs, const float* src, long long src_bs, float* out,
                          long long out_bs, int B, int C, int H, int W, int pitch, irr_stream_t stream);

/* §8(f).4 — backward of the cost volume for the PWC parameters (pad 4, k 1, md 4, stride 1/1): replaces
 * correlation_cuda.backward -> correlation_backward_input1/_input2 (correlation_cuda.cc:86-163,
 * correlation_cuda_kernel.cu:116-300).  grad_out: B x 81 x H x W; grad_f1 / grad_f2: B x C x H x W, either may be NULL.
 *   grad_f1[b,c,y,x] = (1/C) sum_d grad_out[b,d,y,x] * f2[b,c,y+dy,x+dx]
 *   grad_f2[b,c,y,x] = (1/C) sum_d grad_out[b,d,y-dy,x-dx] * f1[b,c,y-dy,x-dx]      (zero outside the image) */
int irr_correlation_bwd(const float* f1, long long f1_bs, const float* f2, long long f2_bs, const float* grad_out,
                        long long go_bs, float* grad_f1, long long g1_bs, float* grad_f2, long long g2_bs, int B, int C,
                        int H, int W, int max_disp, irr_stream_t stream);

/* §8(f).4 — backward of WarpingLayer.forward (models/pwc_modules.py:119-133; autograd through grid_sample with the hard
 * mask as a constant): grad_x[b,c,tap] += mask * w_tap * grad_out[b,c,y,x] (atomic adds — zero-fill grad_x first),
 * grad_flow[b,0/1,y,x] = mask * sum_c grad_out * d(bilinear)/d(ix,iy) * (dim-1)/max(dim_im-1,1)/div_flow (zero-fill
 * grad_flow first).  Either gradient may be NULL.  No batch rotation (autograd use only). */
int irr_warp_bwd(const float* x, long long x_bs, const float* flow, long long flow_bs, const float* lin_x,
                 const float* lin_y, const float* grad_out, long long go_bs, float* grad_x, long long gx_bs,
                 float* grad_flow, long long gf_bs, int B, int C, int H, int W, int H_im, int W_im, float div_flow,
                 int grid_flags, irr_stream_t stream);

/* §8(f).1 — evaluation metrics of the reference's eval-mode losses (losses.py:8-10, 24-37, 634-636, 688-697), one
 * launch per batch, deterministic.  flow / target: B x 2 x H x W; valid, occ_logits, target_occ: B x 1 x H x W or NULL.
 * sums: B x 8 float64 = { S epe*valid, S valid, S outlier, S pred*true, S pred, S true, 0, 0 } with
 * epe = ||target - flow||, outlier = (epe*valid > 3) & (epe*valid / (||target|| + 1e-8) > 0.05), pred = round(sigmoid(occ)). */
int irr_eval_metrics_fwd(const float* flow, long long flow_bs, const float* target, long long target_bs,
                         const float* valid, long long valid_bs, const float* occ_logits, long long occ_bs,
                         const float* target_occ, long long tocc_bs, double* sums, int B, int H, int W,
                         irr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* IRR_B200_H_ */
