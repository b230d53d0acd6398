/* ppst_b200 -- C ABI of the B200-native (sm_100a) PPST swap / inference hot path.
 *
 * Plain pointers and sizes only: no torch / ATen / pybind types cross this boundary.
 * Ownership: the caller allocates every buffer (inputs, outputs, scratch); kernels never allocate
 * or free device memory.  Every entry point enqueues on `stream` (a cudaStream_t passed as void*,
 * NULL = legacy default stream) and returns without synchronising.
 * Return value: 0 on success, a negative PPST_ERR_* code otherwise; ppst_last_error() returns the
 * text of the last error raised on the calling thread.  No C++ exception crosses the ABI.
 *
 * Each declaration cites the reference interface it replaces (paths relative to the reference
 * repository root, wangxb29/PPST).
 */
#ifndef PPST_B200_H_
#define PPST_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PPST_API __attribute__((visibility("default")))
#else
#define PPST_API
#endif

/* element types of activation buffers */
enum { PPST_F32 = 0, PPST_F16 = 1, PPST_BF16 = 2 };

/* error codes */
enum {
  PPST_OK = 0,
  PPST_ERR_ARG = -1,        /* bad shape / null pointer / unsupported combination */
  PPST_ERR_DTYPE = -2,      /* unsupported element type */
  PPST_ERR_CUDA = -3,       /* a CUDA runtime / driver call failed (text in ppst_last_error) */
  PPST_ERR_UNSUPPORTED = -4 /* configuration the kernels do not implement (the reference would return
                               uninitialised memory here, upfirdn2d_kernel.cu:172-223; we refuse) */
};

/* activation selectors (fused_bias_act `act` argument, fused_bias_act_kernel.cu:37-46) */
enum { PPST_ACT_LINEAR = 1, PPST_ACT_LRELU = 3 };

PPST_API const char* ppst_last_error(void);
PPST_API int ppst_version(void);
/* number of kernels launched by this library on the calling process since load (for bench `gpu_launches`) */
PPST_API long long ppst_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * stylegan2_op plugin operators
 * ------------------------------------------------------------------------------------------- */

/* Replaces the pybind entry `fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)`
 * (models/networks/stylegan2_op/fused_bias_act.cpp:11-20, kernel fused_bias_act_kernel.cu:19-49,
 * host wrapper :52-99).
 *   y[i] = f(x[i] + b[(i / step_b) % size_b]) * scale
 *   act=3: grad=0 f(v)= v>0 ? v : v*alpha ; grad=1 f(v)= ref[i]>0 ? v : v*alpha ; grad=2 f=0
 *   act=1: grad=0,1 f(v)=v ; grad=2 f=0
 * `b` may be NULL (size_b = 0: no bias), `ref` may be NULL unless grad==1 with act==3.
 * n = number of elements; step_b = prod(x.shape[2:]).  In-place (y == x) is allowed. */
PPST_API int ppst_fused_bias_act(void* y, const void* x, const void* b, const void* ref, int64_t n, int64_t step_b,
                        int64_t size_b, int act, int grad, float alpha, float scale, int dtype, void* stream);

/* grad_bias[c] = sum over (batch, spatial) of g[n, c, :]  -- the reduction the reference does with
 * `grad_input.sum(dim)` (stylegan2_op/fused_act.py:37-42).  g is [outer, size_b, step_b]; out is f32[size_b]. */
PPST_API int ppst_bias_grad_reduce(float* out, const void* g, int64_t outer, int64_t size_b, int64_t step_b, int dtype,
                          void* stream);

/* Backward of bias + leaky-ReLU on NHWC activations in one pass (training path): gp = g * (y > 0 ? 1 : alpha) * scale
 * (FusedLeakyReLUFunctionBackward, fused_act.py:23-43) AND bias_grad[c] = sum over pixels of the stored gp (`grad_input.sum`,
 * :37-42).  act = 0: no activation, gp is not written (pass NULL), bias_grad = sum of g.  g, y, gp: [pixels, c] bf16 / f32;
 * scratch: ppst_act_grad_sum_scratch_elems(pixels, c) floats.  Deterministic. */
PPST_API int64_t ppst_act_grad_sum_scratch_elems(int64_t pixels, int c);
PPST_API int ppst_act_grad_sum(void* gp, float* bias_grad, const void* g, const void* y, float* scratch, int64_t pixels, int c,
                      float alpha, float scale, int act, int dtype, void* stream);

/* Replaces `upfirdn2d_op.upfirdn2d(input[major,H,W,minor], kernel, up_x, up_y, down_x, down_y,
 * pad_x0, pad_x1, pad_y0, pad_y1)` (stylegan2_op/upfirdn2d.cpp:12-23, upfirdn2d_kernel.cu:53-137,140-272).
 * Zero-insert upsample, pad (negative pads crop), FIR with the flipped `kernel` (f32 [kh,kw], device
 * memory), decimate.  y is [major, out_h, out_w, minor] with
 *   out = (in*up + pad0 + pad1 - k + down) / down      (upfirdn2d_kernel.cu:167-168).
 * Any up/down >= 1 and kernel <= 32x32 is computed (tiled shared-memory path for the six reference modes,
 * a generic path otherwise); nothing is ever left uninitialised. */
PPST_API int ppst_upfirdn2d(void* y, const void* x, const float* kernel, int64_t major, int in_h, int in_w, int minor,
                   int kh, int kw, int up_x, int up_y, int down_x, int down_y, int px0, int px1, int py0,
                   int py1, int dtype, void* stream);

/* Blur (upfirdn2d with up = down = 1, stylegan2_layers.py:142-164) on NHWC activations for the training path:
 *   out[n, oy, ox, :] = sum_{ky,kx} in[n, oy + ky - py0, ox + kx - px0, :] * f[ky][kx]   (zero outside),
 * out size = in + p0 + p1 - k + 1 per axis; flip = 1: f = the flipped kernel (what ppst_upfirdn2d computes); flip = 0: the
 * kernel as given -- with pads (k-1-p0, k-1-p1) that is the adjoint (UpFirDn2dBackward, upfirdn2d.py:24-90).
 * kernel: f32 [kh, kw] device memory, square, 1..4 taps per axis; bf16 (c % 8 == 0) or f32 (c % 4 == 0). */
PPST_API int ppst_blur_nhwc(void* out, const void* in, const float* kernel, int64_t n, int in_h, int in_w, int c, int kh, int kw,
                   int py0, int py1, int px0, int px1, int flip, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * segmented (per-sample, per-channel, optionally per-region) reductions
 * Serves nn.InstanceNorm2d statistics (stylegan2_layers.py:421,527; generator.py:17,21),
 * GAP / GMP (encoder_col.py:39-40,160-161), the masked GAP/GMP of the region branch
 * (encoder_col.py:173-190,217-245) keyed by the integer label map (data/CelebAMask_dataset.py:54-60),
 * and models/networks/function.py:3-35.
 * ------------------------------------------------------------------------------------------- */

/* x: activations, NCHW (channels_last = 0) or NHWC (channels_last = 1), n x c x hw.
 * labels: NULL, or uint8 [n, hw]; n_regions >= 2 with labels, 1 without.
 *   label_mode 0: the byte is the region index (membership: label == r)
 *   label_mode 1: the byte is a membership bitmask (membership: (label >> r) & 1) -- what max-pooled
 *                 one-hot masks become below full resolution (encoder_col.py:217, self.pool(mask))
 * shift: NULL or f32 [n, c]: values are accumulated as (x - shift) (second pass of the exact variance).
 * Outputs (f32, [n, n_regions, c]; any may be NULL):
 *   sum   = sum over pixels of region r of (x - shift)
 *   sumsq = sum over pixels of region r of (x - shift)^2
 *   vmax  = max over ALL pixels of (x if pixel in region r else 0)   == amax(x * onehot_r)
 *           (with n_regions == 1 and labels == NULL: plain max)
 * scratch: f32 workspace of ppst_plane_reduce_scratch_elems(n, c, hw, n_regions) elements.
 * border > 0 (channels_last only; w_row = row width, hw % w_row == 0): x is [n, hw/w_row + 2*border, w_row + 2*border, c] and only its
 * interior is reduced (labels stay [n, hw]).
 * Deterministic (no floating-point atomics). */
PPST_API int64_t ppst_plane_reduce_scratch_elems(int64_t n, int64_t c, int64_t hw, int n_regions);
PPST_API int ppst_plane_reduce(float* sum, float* sumsq, float* vmax, const void* x, const uint8_t* labels,
                      const float* shift, float* scratch, int64_t n, int64_t c, int64_t hw, int n_regions,
                      int label_mode, int channels_last, int dtype, int w_row, int border, void* stream);

/* GAP / GMP of a bilinear (align_corners=False) x`scale` upsampling of an f32 NHWC map without materialising it:
 * the correlation branch of the colour encoder interpolates the warped 64x64 features back to the level's resolution
 * and then only pools them (encoder_col.py:119-130 followed by :160-166).  x: f32 [n, lh, lw, c_stride], channel slice
 * [c_off, c_off + c).  sum / vmax: f32 [n, c] over the (lh*scale) x (lw*scale) virtual map (either may be NULL).
 * scale in {1, 2, 4, 8}; scratch: f32 workspace of ppst_upsample_reduce_scratch_elems(n, lh, c) elements.
 * Deterministic; the reduction split does not depend on n. */
PPST_API int64_t ppst_upsample_reduce_scratch_elems(int64_t n, int lh, int c);
PPST_API int ppst_upsample_reduce(float* sum, float* vmax, const float* x, int64_t n, int lh, int lw, int c_stride,
                         int c_off, int c, int scale, float* scratch, void* stream);

/* Backward reductions of InstanceNorm (+ StyleMod) for the training path: per (sample, channel) sum_g = sum_p g and
 * sum_gx = sum_p g*x over the h*w pixels of the NHWC tensors g (upstream gradient) and x (the layer input), f32 [n, c]
 * each.  With the forward mean / rstd they give the whole gradient of y = (x - mean) * rstd * (1 + s) + b
 * (nn.InstanceNorm2d + StyleMod, stylegan2_layers.py:361-374,414-437): dx = alpha*g + beta*x + gamma per (n, c) -- one
 * ppst_affine_apply pass -- and the style gradients.  c: a multiple of 8 (bf16) / 4 (f32), at most 2048 / 1024.  Deterministic.
 * scratch: f32 workspace of ppst_plane_dot_scratch_elems(n, c, hw) elements. */
PPST_API int64_t ppst_plane_dot_scratch_elems(int64_t n, int64_t c, int64_t hw);
PPST_API int ppst_plane_dot(float* sum_g, float* sum_gx, const void* g, const void* x, float* scratch, int64_t n, int64_t c,
                   int64_t hw, int dtype, void* stream);

/* Backward of the GAP || GMP pooling of the colour encoder's heads (encoder_col.py:39-40,160-161) for the training path;
 * the forward is ppst_plane_reduce (sum, vmax).  ppst_plane_count_eq: cnt[n,c] = number of pixels with x == mx[n,c] (ties
 * share the maximum's gradient, as torch.amax does); scratch: ppst_plane_dot_scratch_elems(n, c, hw) floats.
 * ppst_pool_backward: dx[n,p,c] = ga[n,c] + (x[n,p,c] == mx[n,c] ? gm[n,c] : 0) with ga = grad_gap / hw, gm = grad_gmp / cnt. */
PPST_API int ppst_plane_count_eq(float* cnt, const void* x, const float* mx, float* scratch, int64_t n, int64_t c, int64_t hw,
                        int dtype, void* stream);
PPST_API int ppst_pool_backward(void* dx, const void* x, const float* ga, const float* gm, const float* mx, int64_t n, int64_t c,
                       int64_t hw, int dtype, void* stream);

/* uint8 label map [n, hw] -> one-hot f32 [n, n_labels, hw]  (CelebAMaskDataset._mask_labels,
 * data/CelebAMask_dataset.py:54-60; bit-exact membership test label == i). */
PPST_API int ppst_labels_to_onehot(float* onehot, const uint8_t* labels, int64_t n, int64_t hw, int n_labels, void* stream);

/* Region membership bitmasks at reduced resolution: out[n, h/f, w/f] = OR over the f x f window of
 * (in_mode 0: 1 << label, in_mode 1: the bitmask itself).  Equals MaxPool2d(f) of the one-hot mask
 * (encoder_col.py:38,217).  f = 1 converts an index map into a bitmask map. */
PPST_API int ppst_labels_pool_or(uint8_t* out, const uint8_t* in, int64_t n, int h, int w, int factor, int in_mode,
                        void* stream);

/* ---------------------------------------------------------------------------------------------
 * layout
 * ------------------------------------------------------------------------------------------- */
/* NCHW (src_dtype) -> NHWC (dst_dtype), and back.  dst has channel stride dst_c >= c and the block is
 * written at channel offset c_off (used to build concatenated feature maps, generator.py:272-278). */
PPST_API int ppst_nchw_to_nhwc(void* dst, const void* src, int64_t n, int64_t c, int64_t hw, int64_t dst_c, int64_t c_off,
                      int src_dtype, int dst_dtype, void* stream);
/* nhwc_to_nchw: a, b = NULL, or f32 [n, c] (b may be NULL): dst = src * a + b per (sample, channel) -- the InstanceNorm +
 * StyleMod of ToRGB (stylegan2_layers.py:477-495) folded into the conversion to the reference's NCHW image layout. */
PPST_API int ppst_nhwc_to_nchw(void* dst, const void* src, int64_t n, int64_t c, int64_t hw, int64_t src_c, int64_t c_off,
                      int src_dtype, int dst_dtype, const float* a, const float* b, void* stream);

/* ---------------------------------------------------------------------------------------------
 * convolution (NHWC activations, weights packed [phase][cout][kh*kw*cin])
 * Replaces the F.conv2d / F.conv_transpose2d calls of EqualConv2d (stylegan2_layers.py:188),
 * EqualizedConv2d (:320,331,339-341), nn.Conv2d in generator.py:14-19,180-238, and the 1x1 heads.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  /* geometry */
  int32_t n, in_h, in_w, cin;     /* input  [n, in_h, in_w, cin]  (already padded if the layer pre-pads) */
  int32_t out_h, out_w, cout;     /* logical output grid of ONE phase */
  int32_t kh, kw;                 /* taps of one phase */
  int32_t stride;                 /* input stride (1 or 2) */
  int32_t off_y, off_x;           /* input row/col of tap (0,0) for output (0,0); negative = zero padding */
  /* output placement: out pixel (oy,ox) of phase p is stored at
       (oy*out_sy + p_y, ox*out_sx + p_x) of a [n, full_h, full_w, cout_stride] tensor, channel offset c_off */
  int32_t phases_y, phases_x;     /* 1,1 for a plain conv; 2,2 for the stride-2 transposed conv */
  int32_t phase_off_dy, phase_off_dx; /* off_y/off_x increment per phase index (transposed conv: +1) */
  int32_t full_h, full_w, cout_stride, c_off;
  /* fused epilogue: v = acc + bias[co] + noise_w * noise[n, y, x];  v = v > 0 ? v : v*slope;  v *= gain */
  int32_t has_act;                /* 0: linear (slope ignored, gain still applied) */
  float slope, gain, noise_w;
  int32_t in_dtype, out_dtype;    /* PPST_F32 / PPST_BF16 */
  /* input addressing in ELEMENTS: pixel (y,x) of image i starts at i*in_img_pitch + y*in_row_pitch + x*cin.
     0 = dense (in_w*cin, in_h*in_w*cin).  Lets a conv read the interior of a padded buffer. */
  int64_t in_row_pitch, in_img_pitch;
  /* GEMM use of the tensor-core kernel (ppst_conv_tc only; corrm / warp, ppst_model.py:363, encoder_col.py:124):
     w_img_rows  > 0: image i multiplies rows [i*w_img_rows, ...) of `weight` (a per-sample B matrix) instead of a
                      shared weight.
     split_terms = 3: fp32-accurate product from bf16 operands.  Both operands hold [hi | lo] halves along the
                      channel dim (cin = 2 x segment; hi = bf16(x), lo = bf16(x - hi)); the kernel accumulates
                      hi*hi + hi*lo + lo*hi.  Requires kh = kw = 1. */
  int32_t w_img_rows, split_terms;
  /* ppst_conv_tc only.  out_pad > 0: `out` is a [n, full_h + 2*out_pad, full_w + 2*out_pad, cout_stride] buffer and
     the result is stored into its interior (the border belongs to the caller: see ppst_fill_border) -- the next
     convolution then reads the padded buffer as is.  bias_img_stride > 0: image i adds bias[i*bias_img_stride + c]
     (a per-sample bias, see ppst_modulate_weights). */
  int32_t out_pad, bias_img_stride;
  /* out_pad_trim (0 .. out_pad): the bottom / right border is out_pad - out_pad_trim pixels wide, i.e. the buffer is
     [n, full_h + 2*out_pad - out_pad_trim, full_w + 2*out_pad - out_pad_trim, cout_stride] -- the asymmetric
     (2, 1, 2, 1) reflection padding in front of the blurred stride-2 convolutions (stylegan2_layers.py:517-523);
     the border itself is written by ppst_pad_border. */
  int32_t out_pad_trim;
} ppst_conv_desc;

/* CUDA-core implicit GEMM (fp32 accumulate; any channel counts).  fp32 mode and the thin layers
 * (FromRGB 3->32, ToRGB 128->3).  weight: f32 [phases_y*phases_x][cout][kh*kw*cin]; bias f32[cout] or NULL;
 * noise f32 [n, full_h, full_w] or NULL. */
PPST_API int ppst_conv_simt(void* out, const void* in, const float* weight, const float* bias, const float* noise,
                   const ppst_conv_desc* d, void* stream);

/* tcgen05 / TMEM / TMA implicit GEMM (bf16 in, fp32 accumulate in tensor memory, bf16 or f32 out).
 * Requires cin % 32 == 0, cout % 32 == 0, stride 1 or 2.  weight: bf16, same packing as above.
 * stats (optional, f32 [n_tiles][2][cout_stride]... see ppst_conv_tc_stats_elems): per-tile partial sum / sum of
 * squares of the stored (rounded) outputs, reduced later by ppst_in_coeffs -- the instance-norm statistics
 * of LayerEpilogue (stylegan2_layers.py:414-437) come for free with the convolution. */
PPST_API int64_t ppst_conv_tc_stats_elems(const ppst_conv_desc* d);
PPST_API int ppst_conv_tc(void* out, const void* in, const void* weight, const float* bias, const float* noise,
                 float* stats, const ppst_conv_desc* d, void* stream);
/* Thin 1x1 convolutions that are HBM streams rather than GEMMs.
 * ppst_conv1x1_in_nchw : FromRGB = ConvLayer(3, 32, 1) + FusedLeakyReLU (stylegan2_layers.py:497-555;
 *   encoder_con.py:26, encoder_col.py:27).  x is the NCHW f32 image [n, cin<=4, hw] (the layout conversion is fused),
 *   out NHWC [n, hw, cout_stride] slice [c_off, c_off+cout); w f32 [cout][cin]; v = lrelu(x.w + bias) * gain.
 *   out_border > 0 (needs the row width w_row, hw % w_row == 0): out is [n, hw/w_row + 2*out_border, w_row + 2*out_border, cout_stride]
 *   and the result goes to its interior (border: ppst_pad_border).
 * ppst_conv1x1_thin_out: ToRGB's EqualConv2d(128, 3, 1) (stylegan2_layers.py:484): x NHWC [pixels, cin],
 *   out f32 [pixels, cout], w f32 [cout][cin]. */
PPST_API int ppst_conv1x1_in_nchw(void* out, const float* x, const float* w, const float* bias, int64_t n, int cin,
                         int64_t hw, int cout, int cout_stride, int c_off, int has_act, float slope, float gain,
                         int out_dtype, int w_row, int out_border, void* stream);
PPST_API int ppst_conv1x1_thin_out(float* out, const void* x, const float* w, const float* bias, int64_t pixels, int cin,
                          int cout, int in_dtype, void* stream);

/* Weight gradient of the thin input convolution above (training; torch.autograd's share of
 * ConvLayer(3, C, 1) in the encoders and the discriminator, stylegan2_layers.py:582-649):
 *   dw f32 [cout][cin] = sum_{n,pix} g[n,pix,co] * x[n,ci,pix],  g NHWC (bf16 / f32) [n, hw, cout], x NCHW f32 [n, cin<=4, hw].
 * scratch: f32 workspace of ppst_conv1x1_thin_wgrad_scratch_elems(n, hw, cout) elements.  Deterministic. */
PPST_API int64_t ppst_conv1x1_thin_wgrad_scratch_elems(int64_t n, int64_t hw, int cout);
PPST_API int ppst_conv1x1_thin_wgrad(float* dw, const void* g, const float* x, int64_t n, int cin, int64_t hw, int cout,
                            int g_dtype, float* scratch, void* stream);

/* Per-sample input-channel scaling of packed weights (StyleGAN2-style weight modulation) for ppst_conv_tc with
 * w_img_rows = rows:  out bf16 [n][rows][taps*cin] = w f32 [rows][taps*cin] * a f32 [n][cin].  Folds the InstanceNorm
 * that PRECEDES the first convolution of the generator's feature heads (generator.py:180-231) into that convolution. */
PPST_API int ppst_scale_weights(void* out, const float* w, const float* a, int64_t n, int64_t rows, int taps, int cin,
                       void* stream);

/* Lazy InstanceNorm + StyleMod in front of a zero-padded convolution (StyledConv conv1 -> conv2 inside the generator's
 * residual blocks, stylegan2_layers.py:439-475, generator.py:47-77): with x = a*y + b per (sample, channel),
 *     conv(zero_pad(x), W)[co] = conv(y', W*a)[co] + sum_{tap,c} W[co][tap][c] * b[c],   y' = y with border -b/a.
 * ppst_modulate_weights: w f32 [rows, taps*cin] (packed), a, b f32 [n, cin], base_bias f32 [rows] or NULL
 *     -> out_w bf16 [n, rows, taps*cin] = W*a,  out_bias f32 [n, rows] = base_bias + sum W*b
 *     (feed to ppst_conv_tc with w_img_rows = rows, bias_img_stride = rows).
 * ppst_fill_border: the `pad`-pixel border of the bf16 buffer [n, h+2*pad, w+2*pad, c] <- -b/a (|a| clamped to 1e-20).
 * The producing convolution writes the interior (ppst_conv_desc.out_pad). */
PPST_API int ppst_modulate_weights(void* out_w, float* out_bias, const float* w, const float* a, const float* b,
                          const float* base_bias, int64_t n, int64_t rows, int taps, int cin, void* stream);
PPST_API int ppst_fill_border(void* buf, const float* a, const float* b, int64_t n, int h, int w, int c, int pad, int dtype,
                     void* stream);

/* Border of a padded NHWC buffer [n, pt + h + pb, pl + w + pr, c] from its own interior: mode 1 reflect
 * (nn.ReflectionPad2d, stylegan2_layers.py:517), 2 replicate (nn.ReplicationPad2d, generator.py:12-26,180-231), 0 zero.
 * Producers write the interior (ppst_conv_desc.out_pad, ppst_apply_desc.out_border, ppst_resample out_border, ...), so
 * padding costs one pass over the BORDER instead of a read + write of the whole tensor.  bf16 (c % 8 == 0) or f32
 * (c % 4 == 0), dense channels. */
PPST_API int ppst_pad_border(void* buf, int64_t n, int h, int w, int c, int pt, int pb, int pl, int pr, int mode, int dtype,
                    void* stream);

/* Weight gradient of a convolution on the tcgen05 path -- the training step's share of the convolution work
 * (SURVEY.md 8(f)4).  Replaces what torch.autograd / cuDNN compute for the `weight` of F.conv2d in EqualConv2d
 * (stylegan2_layers.py:188), EqualizedConv2d (:320,331,339-341) and nn.Conv2d (generator.py:14-19); with x and dy
 * exchanged it is also the weight gradient of F.conv_transpose2d (:320).
 *   dw[co][ty][tx][ci] = sum_{n,oy,ox} dy[n,oy,ox,co] * x[n, oy*stride + ty + off_y, ox*stride + tx + off_x, ci]
 * (x = 0 outside the image: off < 0 is zero padding).  x: bf16 NHWC [n,in_h,in_w,cin]; dy: bf16 NHWC [n,out_h,out_w,cout];
 * dw: f32 [cout][kh*kw*cin], the packing ppst_conv_tc consumes.  fp32 accumulation in tensor memory, deterministic
 * (fixed-order reduction over samples and pixel chunks).  The input gradient needs no entry point of its own: it is
 * ppst_conv_tc over dy with the flipped / transposed weights (stride 2: four sub-pixel phases).
 * workspace: device scratch of ppst_conv_wgrad_workspace_bytes(d) bytes, owned by the caller. */
typedef struct {
  int32_t n, in_h, in_w, cin;
  int32_t out_h, out_w, cout;
  int32_t kh, kw, stride;         /* stride 1 or 2 */
  int32_t off_y, off_x;           /* input row / col of tap (0,0) for output (0,0), as in ppst_conv_desc */
} ppst_wgrad_desc;
PPST_API int64_t ppst_conv_wgrad_workspace_bytes(const ppst_wgrad_desc* d);
PPST_API int ppst_conv_wgrad(float* dw, const void* x, const void* dy, const ppst_wgrad_desc* d, void* workspace,
                    int64_t workspace_bytes, void* stream);

/* Per-step weight preparation of the training path (the parameters stay f32 [co, ci, kh, kw], the reference's
 * state_dict layout).
 * ppst_fold_fir: Blur + EqualConv2d of a downsampling ConvLayer (stylegan2_layers.py:142-164,497-555) as ONE
 *   convolution: out [m][kh+fh-1][kw+fw-1] = stencil product of w [m][kh][kw] (m = co*ci) and the FIR kernel k [fh][fw]
 *   (device memory); adjoint = 1: gradient of the folded weight (w, folded size) -> gradient of the original (out [m][kh][kw]).
 * ppst_pack_conv_weights: out bf16 [groups][rows][slots*cols] = scale * w[..][tap slot_src[g*slots + s]] (-1: zero);
 *   transpose = 0: rows = co, cols = ci (forward packing); 1: rows = ci, cols = co (input-gradient packings).
 *   slot_src is HOST memory (at most 256 entries). */
PPST_API int ppst_fold_fir(float* out, const float* w, const float* k, int64_t m, int kh, int kw, int fh, int fw, int adjoint,
                  void* stream);
PPST_API int ppst_pack_conv_weights(void* out, const float* w, int co, int ci, int kh, int kw, int groups, int slots,
                           const int* slot_src, int transpose, float scale, void* stream);

/* 1 if this build/device can run ppst_conv_tc (sm_100 family + driver entry point for tensor maps) */
PPST_API int ppst_conv_tc_available(void);

/* ---------------------------------------------------------------------------------------------
 * normalisation / modulation epilogues (NHWC)
 * ------------------------------------------------------------------------------------------- */
/* Per-(n,c) affine coefficients from reduced statistics:
 *   mean = sum/count (+ shift), var = sumsq/count - (sum/count)^2  (biased, nn.InstanceNorm2d), rstd = rsqrt(var+eps)
 *   a = rstd * (1 + style[n, c]) ; b = -mean * a + style[n, C + c] + bias[c]
 * i.e. InstanceNorm followed by StyleMod `x * (s + 1) + b` (stylegan2_layers.py:361-374) and/or a
 * per-channel bias (FusedLeakyReLU bias after IN in ConvLayer, :527-531).  style (f32, row n = [scale(c) | shift(c)],
 * row stride style_ld elements, 0 = 2c: all style vectors of a generator pass come out of ONE batched linear) / bias /
 * shift may be NULL.  If use_norm == 0: a = 1 + style, b = style_shift + bias.
 * partials: if n_partials > 0, sum/sumsq are [n][n_partials][c] tile partials (from ppst_conv_tc) and are
 * reduced here in a fixed order. */
PPST_API int ppst_in_coeffs(float* a, float* b, const float* sum, const float* sumsq, const float* shift,
                   const float* style, int64_t style_ld, const float* bias, int64_t n, int64_t c, int64_t n_partials,
                   double count, float eps, int use_norm, void* stream);

typedef struct {
  int32_t n, h, w, c;             /* logical input grid (x is [n, h, w, c]) */
  int32_t up;                     /* 1, or 2 = nearest x2 (Upscale2d, stylegan2_layers.py:88-105) */
  int32_t pad_t, pad_b, pad_l, pad_r;
  int32_t pad_mode;               /* 0 zero, 1 reflect (nn.ReflectionPad2d), 2 replicate (nn.ReplicationPad2d) */
  /* y = post( pre(x*a + b) * pre_gain + R ) * post_gain
     pre/post(v) = v >= 0 ? v : v*slope  (slope = 1: identity)
     R = 0 | res*ra + rb | bilinear_x2(res)*ra + rb  (F.interpolate(scale_factor=2, align_corners=False),
     generator.py:75) */
  float pre_slope, pre_gain, post_slope, post_gain;
  int32_t res_mode;               /* 0 none, 1 same size, 2 bilinear x2 from [n, h/2, w/2, c] */
  int32_t x_dtype, res_dtype, out_dtype;
  int32_t out_c_stride, out_c_off; /* destination channel stride / offset (concat) */
  /* out_border > 0: `out` is [n, OH + 2*out_border, OW + 2*out_border, out_c_stride] and the (OH x OW) result goes to
     its interior (border: ppst_pad_border).  res_border > 0: `res` carries such a border and is read at its interior. */
  int32_t out_border, res_border;
} ppst_apply_desc;

/* out: [n, h*up + pad_t + pad_b, w*up + pad_l + pad_r, out_c_stride]; a,b,ra,rb: f32 [n, c] or NULL (1 / 0). */
PPST_API int ppst_affine_apply(void* out, const void* x, const float* a, const float* b, const void* res,
                      const float* ra, const float* rb, const ppst_apply_desc* d, void* stream);

/* ppst_affine_apply that ALSO emits the instance-norm statistics of what it writes: stats f32 [2][n][OH][c] = per output
 * row (of the padded output, OH = h + pad_t + pad_b) and channel the sum / the sum of squares of the stored (rounded)
 * values -- feed to ppst_in_coeffs with n_partials = OH, count = OH * OW.  The feature heads of the generator
 * (generator.py:180-231: ReplicationPad2d -> InstanceNorm2d -> conv) normalise the tensor this pass produces, so neither a
 * padding pass nor a reduction pass over it is needed.  bf16, c % 8 == 0, up == 1; res_mode 0 / 1 with any padding mode;
 * res_mode 2 (bilinear-x2 skip) with no padding or a replicated 1-pixel border (written by the kernel itself). */
PPST_API int ppst_affine_apply_stats(void* out, const void* x, const float* a, const float* b, const void* res,
                            const float* ra, const float* rb, const ppst_apply_desc* d, float* stats, void* stream);

/* ppst_affine_apply with a bilinear-x2 skip (the end of UpsamplingResBlock64, generator.py:63-77) fused with the ToRGB
 * 1x1 convolution that follows it (stylegan2_layers.py:477-495, conv only: 128 -> 3 channels + bias):
 *   rgb_out f32 [n, h, w, 3] = conv1x1(bf16(apply(x, a, b, res)), rgb_w [3, 128]) + rgb_bias [3]
 * out: the bf16 activation itself [n, h, w, out_c_stride], or NULL when nothing else reads it (plain decode).
 * d: c == 128, res_mode == 2, up == 1, no padding, bf16 in / out, h even, w % 64 == 0. */
PPST_API int ppst_affine_apply_torgb(void* out, float* rgb_out, const void* x, const float* a, const float* b,
                            const void* res, const float* rgb_w, const float* rgb_bias, const ppst_apply_desc* d,
                            void* stream);

/* Resampling into a channel slice of an NHWC tensor.
 * mode 0: bilinear, align_corners=False (F.interpolate, generator.py:274,277; encoder_col.py:130)
 * mode 1: adaptive average pool with integer window (nn.AdaptiveAvgPool2d, generator.py:257,276;
 *         F.adaptive_avg_pool2d encoder_col.py:119-122) -- requires in % out == 0
 * a, b: NULL, or f32 [n, c] (b may be NULL = 0): every SOURCE sample is first mapped to lrelu(x*a + b, slope) -- the
 *       InstanceNorm + PReLU that ends a feature head (generator.py:180-231) applied on load, so the normalised map
 *       that generator.py:272-278 pools and interpolates is never written to memory.
 * in_border / out_border: x / out carry a border of that many pixels (buffers [n, H + 2*border, W + 2*border, .])
 *       and are addressed at their interiors. */
PPST_API int ppst_resample(void* out, const void* x, int64_t n, int in_h, int in_w, int out_h, int out_w, int c,
                  int out_c_stride, int out_c_off, int mode, int in_dtype, int out_dtype, const float* a,
                  const float* b, float slope, int in_border, int out_border, void* stream);

/* Adjoints of the x2 upsamplings and of the reflection pad for the training path, all as gathers with a fixed summation
 * order (deterministic; ATen's upsample_bilinear2d_backward scatters with atomics).  bf16 (c % 8 == 0) or f32 (c % 4 == 0).
 *   ppst_upsample2x_backward: g NHWC [n, 2h, 2w, c] -> gin [n, h, w, c]; nearest = 0: adjoint of ppst_resample mode 0 at scale 2
 *     (the bilinear skip of UpsamplingResnetBlock, generator.py:75); nearest = 1: adjoint of Upscale2d (stylegan2_layers.py:
 *     88-105) = the 2x2 box sum.
 *   ppst_reflect_pad_backward: g NHWC [n, h+pt+pb, w+pl+pr, c] -> gin [n, h, w, c], adjoint of nn.ReflectionPad2d
 *     (ppst_affine_apply with pad_mode 1). */
PPST_API int ppst_upsample2x_backward(void* gin, const void* g, int64_t n, int h, int w, int c, int nearest, int dtype,
                             void* stream);
PPST_API int ppst_reflect_pad_backward(void* gin, const void* g, int64_t n, int h, int w, int c, int pt, int pb, int pl, int pr,
                              int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * small dense layers and vector normalisation (f32)
 * ------------------------------------------------------------------------------------------- */
/* y[m, n] = act_in(x)[m, :] . w[n, :] * w_scale + b[n] * b_scale ; relu_in applies ReLU to x first
 * (projector MLPs, encoder_col.py:50-86; EqualLinear / EqualizedLinear, stylegan2_layers.py:205-273).
 * x: [m, k] row stride ldx; w: [n, k] f32; b NULL or [n]. */
PPST_API int ppst_linear(float* y, const float* x, const float* w, const float* b, int64_t m, int64_t n, int64_t k,
                int64_t ldx, float w_scale, float b_scale, int relu_in, void* stream);

/* Backward of ppst_linear for the training path (EqualLinear / EqualizedLinear / the projector MLPs under torch.autograd in
 * the reference): y[m, n] = x[m, k] . w[n, k]^T with a tiny batch m.
 *   ppst_linear_dgrad: dx[m, k] = gy[m, n] . w[n, k]    (scratch: ppst_linear_dgrad_scratch_elems(m, n, k) floats)
 *   ppst_linear_wgrad: dw[n, k] = gy[m, n]^T . x[m, k]
 * All f32, contiguous, k a multiple of 4; fixed summation order. */
PPST_API int64_t ppst_linear_dgrad_scratch_elems(int64_t m, int64_t n, int64_t k);
PPST_API int ppst_linear_dgrad(float* dx, const float* gy, const float* w, float* scratch, int64_t m, int64_t n, int64_t k,
                      void* stream);
PPST_API int ppst_linear_wgrad(float* dw, const float* gy, const float* x, int64_t m, int64_t n, int64_t k, void* stream);

/* mode 0: y = v * rsqrt(sum v^2 + eps)        (util.normalize, util/util.py:18-22, eps = 1e-8)
 * mode 1: y = v / max(||v||, eps)              (F.normalize, encoder_col.py:166, eps = 1e-12) */
PPST_API int ppst_l2_normalize(float* y, const float* x, int64_t rows, int64_t cols, float eps, int mode, void* stream);

/* y = a*(1-r) + b*r  (util.lerp, util/util.py:32-35) */
PPST_API int ppst_lerp(float* y, const float* a, const float* b, float r, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * correlation swap (models/ppst_model.py:330-387)
 * ------------------------------------------------------------------------------------------- */
/* PPSTModel.Rselfcorr (:330-339): fea NHWC [n, 256, 256, 64] -> out NHWC slice [n, 64, 64, 256] written at
 * channel offset c_off of a tensor with channel stride out_c_stride. */
PPST_API int ppst_rselfcorr(void* out, const void* fea, int64_t n, int h, int w, int c, int out_c_stride, int c_off,
                   int in_dtype, int out_dtype, void* stream);

/* corrm preparation (:349-361): per pixel, centre the first `c_center` channels over the channel dim, then
 * L2-normalise all c channels (+ DBL_EPSILON).  in: NHWC [rows, c]; out f32 [rows, c]. */
PPST_API int ppst_corr_prepare(float* out, const void* in, int64_t rows, int c, int c_center, int in_dtype, void* stream);

/* Operands of the 3-term split tensor-core GEMM (ppst_conv_desc.split_terms = 3): every f32 value x becomes
 * hi = bf16(x), lo = bf16(x - hi), stored as [hi | lo] halves along the K dimension.
 *   ppst_corr_prepare_split : ppst_corr_prepare with split output, bf16 [rows, 2c]
 *   ppst_split_bf16         : f32 [batch, rows, cols] -> bf16 [batch, rows, 2*cols]; transpose=1: -> bf16
 *                             [batch, cols, 2*rows] (K-major B operand of corr @ feat from an NHWC feature map)
 *   ppst_softmax_rows_split : softmax(x / divisor) per row, bf16 [rows, 2*cols] (corr as the A operand of warp) */
PPST_API int ppst_corr_prepare_split(void* out, const void* in, int64_t rows, int c, int c_center, int in_dtype,
                            void* stream);
PPST_API int ppst_split_bf16(void* out, const float* in, int64_t batch, int64_t rows, int64_t cols, int transpose,
                    void* stream);
PPST_API int ppst_softmax_rows_split(void* out, const float* x, int64_t rows, int64_t cols, float divisor, void* stream);

/* Batched f32 GEMM  C[b] = alpha * A[b] (m x k, row-major, lda) * B[b]  with B either k x n row-major
 * (trans_b = 0) or n x k row-major (trans_b = 1); 3-term bf16 split on tcgen05 is a later round -- this is
 * the exact-f32 CUDA-core path (corrm logits, :363; warp corr @ feat, encoder_col.py:124,133). */
PPST_API int ppst_gemm_f32(float* c, const float* a, const float* b, int64_t batch, int64_t m, int64_t n, int64_t k,
                  int64_t lda, int64_t ldb, int64_t ldc, int64_t stride_a, int64_t stride_b, int64_t stride_c,
                  int trans_b, float alpha, void* stream);

/* in-place row softmax of (x / divisor), f32 [rows, cols]  (F.softmax(logits / 0.01, dim=-1), ppst_model.py:363) */
PPST_API int ppst_softmax_rows(float* x, int64_t rows, int64_t cols, float divisor, void* stream);

/* ---------------------------------------------------------------------------------------------
 * decode(target=...) post-process
 * ------------------------------------------------------------------------------------------- */
/* Guided-filter smoothing of a decoded image against the content image, entirely on the device: replaces
 * PPSTModel.decode's target branch (models/ppst_model.py:288-306) = util.tensor2im (util/util.py:98-131) ->
 * photo_gif.GIFSmoothing(r, eps).process (photo_gif.py:13-46: cv2.ximgproc.guidedFilter(guide=target, src=out)) ->
 * ToTensor -> (x - 0.5) * 2.  src / guide / out: NCHW f32 [n, 3, h, w] with values in [-1, 1]; radius < min(h, w).
 * workspace: device scratch of at least ppst_guided_smooth_workspace_bytes(n, h, w) bytes, owned by the caller. */
PPST_API int64_t ppst_guided_smooth_workspace_bytes(int64_t n, int h, int w);
PPST_API int ppst_guided_smooth(float* out, const float* src, const float* guide, int64_t n, int h, int w, int radius,
                       float eps, void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * image pre- / post-processing around the path, integer / byte work: bit-exact
 * ------------------------------------------------------------------------------------------- */
/* One pass of Pillow's 8-bit resampler (Image.resize(..., Image.BICUBIC) in data/base_dataset.py:150-172; Pillow
 * src/libImaging/Resample.c ImagingResampleHorizontal_8bpc / Vertical_8bpc):
 *   out = clip8((2^21 + sum_k in[first + k] * kk[o][k]) >> 22)
 * in: uint8 [n, in_h, in_w, c]; axis 0 resamples x (out [n, in_h, out_len, c]), axis 1 resamples y (out [n, out_len,
 * in_w, c]).  bounds int32 [out_len][2] = (first input index, tap count), kk int32 [out_len][ksize] fixed-point
 * (22 fractional bits) coefficients, both as Pillow's precompute_coeffs + normalize_coeffs_8bpc produce them (built on
 * the host: ppst_b200/imageio.py). */
PPST_API int ppst_resample_u8(uint8_t* out, const uint8_t* in, const int* bounds, const int* kk, int64_t n, int in_h, int in_w,
                     int out_len, int c, int ksize, int axis, void* stream);
/* transforms.ToTensor() + Normalize(mean, std) (data/base_dataset.py:143-148): uint8 [n, hw, c] -> f32 [n, c, hw],
 * (x / 255 - mean) / std with every step rounded to fp32. */
PPST_API int ppst_u8_to_nchw_norm(float* out, const uint8_t* in, int64_t n, int64_t hw, int c, float mean, float stdv,
                         void* stream);
/* util.tensor2im (util/util.py:98-131): f32 [n, c, hw] -> uint8 [n, hw, c];
 * normalize != 0: trunc(clip((x + 1) / 2 * 255, 0, 255)), else trunc(clip(x * 255, 0, 255)), in fp32. */
PPST_API int ppst_tensor2im(uint8_t* out, const float* in, int64_t n, int64_t hw, int c, int normalize, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PPST_B200_H_ */
