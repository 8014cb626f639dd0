/*
 * riders_b200 - C ABI of the B200-native RIDERS radar-to-dense-depth hot path.
 *
 * Drop-in boundary.  The reference (MMOCKING/RIDERS) is
 * pure Python: its "FFI" for this path is the set of Python call sites listed
 * beside each entry point below (paths relative to the reference repo).  A
 * maintainer binds this library with ctypes and keeps the
 * Python classes / signatures unchanged.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - all tensors are dense float32 unless stated; layouts are spelled per call;
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous and
 *     never synchronise; workspaces are caller-owned (size from *_workspace_bytes);
 *   - return 0 on success, a negative RIDERS_E* code otherwise;
 *     riders_last_error_string() describes the last failure on this thread;
 *   - no CPU fallback exists: without an sm_100 device every launch returns
 *     RIDERS_ENODEVICE / RIDERS_ECUDA.
 */
#ifndef RIDERS_B200_H_
#define RIDERS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RIDERS_OK          0
#define RIDERS_EINVAL     -1   /* bad argument (shape, null pointer, unsupported size)        */
#define RIDERS_ECUDA      -2   /* CUDA runtime error at launch                                */
#define RIDERS_ENODEVICE  -3   /* no sm_100 device / kernel image not loadable                */
#define RIDERS_EQHULL     -4   /* triangulation impossible (<3 knots or all collinear)        */
#define RIDERS_EINTERNAL  -5   /* internal capacity exceeded (reported, never silent)         */

#define RIDERS_LAYOUT_NCHW 0
#define RIDERS_LAYOUT_NHWC 1   /* torch channels_last: logical NCHW, memory N,H,W,C            */

#define RIDERS_MAX_LEVELS 8

const char* riders_last_error_string(void);
int         riders_abi_version(void);
/* compute capability major*10+minor of the current device, or <0 */
int         riders_device_arch(void);
/* number of CUDA kernels this library has launched in this process (all threads, monotonic) */
unsigned long long riders_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * A1  per-radar-point feature gather  ==  torchvision.ops.roi_pool x n_levels
 *     replaces RCNet/networks.py:418-433 (5 calls: latent @1/32 + skips @1/2..1/16).
 *     feat[l]   : (B, C[l], Hl[l], Wl[l]) in `layout`
 *     rois      : (R, 5) rows (batch_index, x1, y1, x2, y2), RoI order = output order
 *     out[l]    : (R, C[l], oh[l], ow[l]) in the SAME layout
 *     argmax[l] : optional int32, same shape/layout as out[l]; flat index h*Wl+w of the
 *                 selected input element, -1 for an empty bin (needed by the backward)
 *     max_roi_w/h : upper bound of (x2-x1)/(y2-y1) in image pixels (RC-Net: the patch size);
 *                 sizes the TMA staging box of the channels-last path.  0 = unknown (generic
 *                 kernel); RoIs that exceed the bound are still correct (in-kernel direct reads).
 *     Bit-exact with torchvision's RoIPool (quantised coords, float32 bin edges, first max).
 * ------------------------------------------------------------------------------------------ */
int riders_roi_gather_fwd(int n_levels,
                          const float* const* feat, const int* C, const int* Hl, const int* Wl,
                          const float* spatial_scale, const int* oh, const int* ow,
                          int layout, int B,
                          const float* rois, int R,
                          int max_roi_w, int max_roi_h,
                          float* const* out, int32_t* const* argmax,
                          void* stream);

/* The same with a per-level output pixel stride (channels-last only): out[l] points at channel c0 of a wider
 * (R, oh, ow, ld) map, ld = out_pixel_stride[l] floats >= C[l], so the gathered skip lands directly inside the
 * decoder's concat buffer [upconv | skip] (utils/net_utils.py:474-569 torch.cat) and is never copied again.
 * out_pixel_stride NULL or 0 entries = dense.  argmax[l] (if any) stays dense.                             */
int riders_roi_gather_fwd_placed(int n_levels,
                                 const float* const* feat, const int* C, const int* Hl, const int* Wl,
                                 const float* spatial_scale, const int* oh, const int* ow,
                                 int layout, int B,
                                 const float* rois, int R,
                                 int max_roi_w, int max_roi_h,
                                 float* const* out, const int* out_pixel_stride, int32_t* const* argmax,
                                 void* stream);

/* gradient of A1 w.r.t. feat: deterministic (no float atomics); grad_feat must be zero-filled
 * by the caller.  Used by rcnet_main.train (RCNet/rcnet_main.py:342-358).                    */
int riders_roi_gather_bwd(int n_levels,
                          const float* const* grad_out, const int32_t* const* argmax,
                          const int* C, const int* Hl, const int* Wl, const int* oh, const int* ow,
                          const float* spatial_scale,
                          int layout, int B, const float* rois, int R,
                          float* const* grad_feat, void* stream);

/* ------------------------------------------------------------------------------------------
 * A2  LinearAttention.forward           RCNet/linear_attention.py:18-45
 *     q (N,L,H,D), k,v (N,S,H,D), masks (N,L)/(N,S) float or NULL, out (N,L,H,D)
 * ------------------------------------------------------------------------------------------ */
int riders_linear_attention_fwd(const float* q, const float* k, const float* v,
                                const float* q_mask, const float* kv_mask,
                                int N, int L, int S, int H, int D, float eps,
                                float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * A3  LoFTREncoderLayer.forward         RCNet/linear_attention.py:112-135
 *     x (N,L,C), src (N,S,C) -> out (N,L,C); all six linears are bias-free, row-major
 *     (out_features, in_features) exactly as torch stores nn.Linear.weight:
 *       wq,wk,wv,wm (C,C); w1 (2C,2C); w2 (C,2C); LayerNorm gamma/beta (C).
 *     impl: RIDERS_LOFTR_SIMT    fp32 FFMA kernels (any C%16==0), bit-for-bit fp32 semantics
 *           RIDERS_LOFTR_TCGEN05 fused tensor-core layer (C=128, nhead=8, L=S=21 tokens):
 *                                tcgen05.mma kind::f16 with a 2-term fp16 hi/lo operand split, 3 MMAs (fp32-accurate)
 *                                (attention blocks on mma.sync with the same split).  Operands must stay inside the fp16 range
 *                                (|value| <= 65504): a violation sets word 0 of `workspace` (int32, reset by the call) to 1 and
 *                                the caller recomputes with RIDERS_LOFTR_SIMT - never silent Inf / NaN.
 *     workspace: riders_loftr_workspace_bytes(N, max(L,S), C, impl) bytes, 256-byte aligned; its first 256 bytes are a status block.
 * ------------------------------------------------------------------------------------------ */
#define RIDERS_LOFTR_SIMT     0
#define RIDERS_LOFTR_TCGEN05  1

typedef struct {
    const float* wq; const float* wk; const float* wv; const float* wm;
    const float* w1; const float* w2;
    const float* ln1_g; const float* ln1_b; const float* ln2_g; const float* ln2_b;
} riders_loftr_weights;

size_t riders_loftr_workspace_bytes(int N, int tokens, int d_model, int impl);
/* 1 if `impl` is compiled into this library and covers (d_model, nhead, tokens), else 0 */
int    riders_loftr_impl_available(int impl, int d_model, int nhead, int tokens);

int riders_loftr_layer_fwd(const float* x, const float* src, const riders_loftr_weights* w,
                           int N, int L, int S, int d_model, int nhead,
                           float attn_eps, float ln_eps, int impl,
                           float* out, void* workspace, void* stream);

/* A3, training path (BASELINE configs[3]; the reference differentiates this layer with torch autograd inside
 * rcnet_main.train, RCNet/rcnet_main.py:342-358).
 *   riders_loftr_layer_fwd_train: the same layer, keeping q, k, v, msg, merge output, m, pre-ReLU h and the mlp output in
 *     `saved` (riders_loftr_saved_bytes, float32) for the backward.
 *   riders_loftr_layer_bwd: given d_out, writes d_x (and d_src for a cross layer, x != src; for a self layer d_x carries both
 *     paths and d_src is ignored) and ACCUMULATES the weight gradients into d_w (same struct, float32 buffers the caller zeroes).
 *   dtype: RIDERS_DTYPE_F32 / RIDERS_DTYPE_BF16 for the token tensors crossing the ABI (x, src, out, d_out, d_x, d_src); weights,
 *     weight gradients and `saved` are float32 (fp32 master weights, fp32 accumulation).  Deterministic (no atomics).
 *   workspace: riders_loftr_train_workspace_bytes bytes, 256-byte aligned.                                                   */
#define RIDERS_DTYPE_F32   0
#define RIDERS_DTYPE_BF16  1
size_t riders_loftr_saved_bytes(int N, int L, int S, int d_model);
size_t riders_loftr_train_workspace_bytes(int N, int L, int S, int d_model);
int riders_loftr_layer_fwd_train(const void* x, const void* src, const riders_loftr_weights* w,
                                 int N, int L, int S, int d_model, int nhead, float attn_eps, float ln_eps, int dtype,
                                 void* out, void* saved, void* workspace, void* stream);
int riders_loftr_layer_bwd(const void* x, const void* src, const riders_loftr_weights* w, const void* saved,
                           const void* d_out, int N, int L, int S, int d_model, int nhead,
                           float attn_eps, float ln_eps, int dtype,
                           void* d_x, void* d_src, const riders_loftr_weights* d_w,
                           void* workspace, void* stream);

/* A4  LocalFeatureTransformer.forward   RCNet/linear_attention.py:159-184
 *     layer_kind[i]: 0 = 'self' (f0=layer(f0,f0); f1=layer(f1,f1)), 1 = 'cross'
 *     (f0=layer(f0,f1); f1=layer(f1,f0_new)).  feat0 (N,L,C), feat1 (N,S,C) are not modified;
 *     results land in out0/out1 (may alias nothing else).  `layers` is a HOST array.
 *     RIDERS_LOFTR_TCGEN05 runs all 2 * n_layers layer calls in ONE persistent launch (n_layers <= 8).  */
int riders_loftr_transformer_fwd(const float* feat0, const float* feat1,
                                 const riders_loftr_weights* layers, const int* layer_kind, int n_layers,
                                 int N, int L, int S, int d_model, int nhead,
                                 float attn_eps, float ln_eps, int impl,
                                 float* out0, float* out1, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * A5  thresholded confidence -> quasi-dense depth   RCNet/rcnet_main.py:460-487
 *     conf (R, ph, pw): per-point correspondence crops; if conf_is_logits != 0 the kernel
 *                       applies the sigmoid of RCNetModel.forward (rcnet_model.py:123) itself
 *     pts (R,3): pad-shifted (x, y, z) exactly as forward_output receives them
 *     frame_offsets (B+1) int32 DEVICE prefix of points per frame (ragged K)
 *     depth, response (B,H,W) float32; select (B,H,W) int32 = argmax point index within the
 *     frame (first maximum; 0 where nothing fired), may be NULL.
 * ------------------------------------------------------------------------------------------ */
int riders_confidence_reduce(const float* conf, int conf_is_logits,
                             const float* pts, const int* frame_offsets,
                             int B, int H, int W, int ph, int pw, float thr,
                             float* depth, float* response, int32_t* select,
                             void* stream);

/* ------------------------------------------------------------------------------------------
 * A6/A7  compute_scale_and_shift_ls + LeastSquaresEstimator.apply/clamp
 *        modules/estimator.py:5-29, 108-125
 *     pred, target (B,H,W) float32; mask (B,H,W) uint8 (numpy/torch bool)
 *     sums (B,5) double out: a00,a01,a11,b0,b1 (deterministic fp64 tree, may be NULL)
 *     scale_shift (B,2) double out (0,0 when det <= 0)
 *     out (B,H,W) float32 or NULL: float32(float32(pred*float32(scale)) + float32(shift)),
 *     then out[out > clamp_hi] = clamp_hi if (clamp_flags & 1), out[out < clamp_lo] = clamp_lo
 *     if (clamp_flags & 2); the caller passes clamp_hi = float32(1/clamp_min) and
 *     clamp_lo = float32(1/clamp_max) (estimator.py:115-125; clamp_min <= 0 clears bit 0).
 * ------------------------------------------------------------------------------------------ */
size_t riders_lsq_workspace_bytes(int B, int H, int W);
int riders_lsq_align(const float* pred, const float* target, const uint8_t* mask,
                     int B, int H, int W,
                     double* sums, double* scale_shift,
                     float* out, float clamp_hi, float clamp_lo, int clamp_flags,
                     void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * N1  Optimizer.optimize_scale + apply_scale + clamp_min_max   modules/estimator.py:129-185
 *     (the scripts' default alignment, val_zju.py:156-167: global_alignment='s')
 *     minimises f(x) = sum_valid |x*pred - target| over [lo, hi] with the bounded Brent iteration of
 *     scipy.optimize.minimize_scalar (xatol, maxiter as SciPy: 1e-5, 500); f is evaluated in fp64 over
 *     the compacted valid pixels (deterministic).  max_valid >= valid pixels per frame (extra ones are
 *     dropped - pass H*W to be safe).  scale (B) double out; fval (B) double, nfev (B) int32 optional.
 *     out (B,H,W) or NULL: float32(pred * float32(scale)) then the clamps as in riders_lsq_align.
 * ------------------------------------------------------------------------------------------ */
size_t riders_l1_scale_workspace_bytes(int B, int max_valid);
int riders_l1_scale_align(const float* pred, const float* target, const uint8_t* mask,
                          int B, int H, int W, int max_valid,
                          double lo, double hi, double xatol, int maxiter,
                          double* scale, double* fval, int32_t* nfev,
                          float* out, float clamp_hi, float clamp_lo, int clamp_flags,
                          void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * N4  RC-Net input prologue: run_rcnet_zju.py:214-240 (shift points by the pads, box = shifted point -/+ pads),
 *     RCNet/rcnet_transforms.py:100-103 (max > 1 -> Tensor.int(), decided per frame) and :258-261 (/ 255.0),
 *     RCNet/rcnet_main.py:441-444 (edge pad by (pw//2, ph//2)).
 *     image (B,C,H,W) NCHW float32; points (R,3) UNSHIFTED (x, y, z), frames concatenated;
 *     frame_offsets (B+1) int32 DEVICE; padded (B,C,H+2*pad_y,W+2*pad_x) in out_layout;
 *     points_shifted (R,3); rois (R,5) = (frame, x1, y1, x2, y2) for riders_roi_gather_fwd;
 *     frame_flags (B) int32 workspace (1 where the frame held a value > 1).
 * N2  SML input assembly: val_zju.py:181-196 + modules/midas/transforms.py (Resize with cv2.INTER_NEAREST on
 *     every item, NormalizeIntermediate, PrepareForNet): x (B,3,out_h,out_w) = [(int_depth-mean)/std |
 *     (int_scales-mean)/std | 0.299 r + 0.587 g + 0.114 b], d (B,1,out_h,out_w) = resized int_depth.
 *     image: (B,H,W,3) if image_hwc (the reference's sample layout) else (B,3,H,W).
 * ------------------------------------------------------------------------------------------ */
int riders_rcnet_prologue(const float* image, int B, int C, int H, int W, int pad_x, int pad_y, int out_layout,
                          const float* points, const int32_t* frame_offsets, int R,
                          float* padded, float* points_shifted, float* rois, int32_t* frame_flags, void* stream);
int riders_sml_input_assemble(const float* int_depth, const float* int_scales, const float* image, int image_hwc,
                              int B, int H, int W, int out_h, int out_w,
                              float depth_mean, float depth_std, float scales_mean, float scales_std,
                              float* x, float* d, void* stream);

/* ------------------------------------------------------------------------------------------
 * N2/N3  scale scaffold (val_zju.py:135-138,171-179) with the RC-Net -> SML PNG hand-off in memory
 *     int_depth (aligned prior), rcnet_depth (quasi-dense METRIC depth from A5, or NULL),
 *     sparse_inv / sparse_valid (radar inverse depth and mask), all (B,H,W).
 *     png_quantize != 0 applies the 16-bit PNG codec of data/data_utils.py:94-143 (uint32(z*256)/256).
 *     int_scales (B,H,W) out = normalize_unit_range(scales); minmax (B,2) float32 optional;
 *     status (B) optional: 1 where max-min <= eps (the reference raises ValueError; scales left raw).
 * N4  riders_depth_metrics: per-frame MAE, RMSE, iMAE, iRMSE, AbsRel, SqRel, delta1, n  (B,8) double
 *     utils/eval_utils.py:17-117 as called from val_zju.py:212-231 (mask: gt > 0, min_val < gt < max_val).
 *     workspace: riders_glue_workspace_bytes(B, H, W), 8-byte aligned.
 * ------------------------------------------------------------------------------------------ */
size_t riders_glue_workspace_bytes(int B, int H, int W);
int riders_scale_scaffold(const float* int_depth, const float* rcnet_depth,
                          const float* sparse_inv, const uint8_t* sparse_valid,
                          int B, int H, int W, float min_depth, float max_depth, int png_quantize,
                          float* int_scales, float* minmax, int32_t* status,
                          void* workspace, void* stream);
int riders_depth_metrics(const float* pred, const float* gt, int B, int H, int W,
                         float min_val, float max_val, double* metrics, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * A8  Interpolator2D / interpolate_knots(method='linear')   modules/interpolator.py:7-50
 *     riders_extract_knots: row-major scan of `valid` -> knot_xy (B,max_knots,2) int32 (x,y),
 *       knot_scale = sparse/pred, knot_shift = sparse-pred (float32), knot_count (B) int32 (the
 *       true count; knots beyond max_knots are dropped - the caller must check).
 *     riders_interp_linear: Delaunay triangulation of each frame's knots on the device
 *       (exact integer predicates) + per-pixel barycentric fill in fp64, `fill` outside the
 *       convex hull, float32 out (B,H,W).  Knots must be distinct pixels with 0 <= x,y < 8192.
 *       status (B) int32: 0 ok, RIDERS_EQHULL where SciPy would raise QhullError (the frame is
 *       then filled with `fill`), RIDERS_EINTERNAL if the edge pool overflowed or knot_count[b] > max_knots
 *       (a truncated knot set is never triangulated: the frame is filled with `fill`).  Knots of frame b are knot_xy[b*max_knots .. +knot_count[b]).
 *       If tri_out != NULL it receives the triangles (B, 2*max_knots, 3) int32 (-1 padded) and
 *       tri_count (B).
 * ------------------------------------------------------------------------------------------ */
int riders_extract_knots(const float* pred, const float* sparse, const uint8_t* valid,
                         int B, int H, int W, int max_knots,
                         int32_t* knot_xy, float* knot_scale, float* knot_shift, int32_t* knot_count,
                         void* stream);
size_t riders_interp_workspace_bytes(int B, int max_knots, int H, int W);
/* griddata(method='nearest') of the same call site (interpolate_knots passes any griddata method through): value of the nearest
 * knot, exact integer distances, lowest knot index among equidistant ones; `fill` only for a frame without knots.            */
int riders_interp_nearest(const int32_t* knot_xy, const float* knot_val, const int32_t* knot_count,
                          int B, int max_knots, int H, int W, float fill, float* out, void* stream);
int riders_interp_linear(const int32_t* knot_xy, const float* knot_val, const int32_t* knot_count,
                         int B, int max_knots, int H, int W, float fill,
                         float* out, int32_t* status,
                         int32_t* tri_out, int32_t* tri_count,
                         void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * Decoder glue next to the gather: nearest-neighbour resize of channels-last maps, bit-exact with
 * torch.nn.functional.interpolate(x, size=(H, W)) (mode 'nearest'); utils/net_utils.py:149-190.
 *     x (N, C, h, w) / out (N, C, H, W), both in memory order N,H,W,C; C % 4 == 0.
 * ------------------------------------------------------------------------------------------ */
int riders_upsample_nearest_nhwc(const float* x, int N, int C, int h, int w, int H, int W,
                                 float* out, void* stream);
/* out[p][0:Ca] = leaky_relu(a[p][:], slope) (or a copy if !act_on_a); out[p][Ca:Ca+Cb] = b[p][:] for
 * `pixels` channels-last pixels: DecoderBlock's activation + torch.cat([y, skip], 1) in one pass
 * (utils/net_utils.py:474-569).  Ca, Cb multiples of 4; Cb may be 0 (activation only, out of place). */
int riders_leaky_concat_nhwc(const float* a, int Ca, const float* b, int Cb, long long pixels,
                             float slope, int act_on_a, float* out, void* stream);
/* the same with the (BatchNorm-folded) convolution bias added first: out[p][0:Ca] = leaky(a[p][:] + bias[:]) */
int riders_bias_leaky_concat_nhwc(const float* a, const float* bias, int Ca, const float* b, int Cb,
                                  long long pixels, float slope, float* out, void* stream);
/* ... into a map of out_pixel_stride >= Ca + Cb floats per pixel: with Cb = 0 and the skip channels already
 * in place behind channel Ca (riders_roi_gather_fwd_placed) only the Ca produced channels are written.   */
int riders_bias_leaky_place_nhwc(const float* a, const float* bias, int Ca, const float* b, int Cb,
                                 long long pixels, float slope, float* out, int out_pixel_stride, void* stream);
/* Sub-pixel form of "nearest 2x resize -> 3x3 convolution": y4 (N, 4*Ca, h, w) is the low-resolution convolution
 * with phase-major output channels ((2a+b)*Ca + c, see DecoderBlock.fold_upsample_); this writes
 * out (N, Ca+Cb, 2h, 2w): out[n][2i+a][2j+b][0:Ca] = act(y4[n][i][j][(2a+b)*Ca + :] + bias), [Ca:] = skip.   */
int riders_d2s_bias_act_concat_nhwc(const float* y4, const float* bias, int Ca, const float* skip, int Cb,
                                    int N, int h, int w, float slope, float* out, void* stream);
/* ... into a map of out_pixel_stride >= Ca + Cb floats per pixel (skip already in place: Cb = 0, skip = NULL) */
int riders_d2s_bias_act_place_nhwc(const float* y4, const float* bias, int Ca, const float* skip, int Cb,
                                   int N, int h, int w, float slope, float* out, int out_pixel_stride, void* stream);
/* Sub-pixel form for blocks whose target is 2n+1 along ONE axis (torch nearest: source max(j-1,0)>>1; RC-Net 30x12 -> 60x25,
 * 7x3 -> 15x6): y4 (N,4*Ca,hp,wp) = phase convolution of x replicate-padded by one on the left (sx) / top (sy); yfirst =
 * 3x1 / 1x3 phase convolution of x's first column (N,2*Ca,h,1) / row (N,2*Ca,1,w); out (N, ld, 2hp-sy, 2wp-sx), channels
 * [0,Ca) = act(. + bias) with pixel stride ld = out_pixel_stride (the skip, if any, is already in place behind them).    */
int riders_d2s_shift_bias_act_place_nhwc(const float* y4, const float* yfirst, const float* bias, int Ca,
                                         int N, int hp, int wp, int sy, int sx, float slope,
                                         float* out, int out_pixel_stride, void* stream);
/* The two consumers above for either form of the phase convolution.  taps = 3: y4 from the 3x3 phase kernels, (N,4*Ca,h,w)
 * [shift: (N,4*Ca,hp,wp)].  taps = 2: y4 from the equivalent 2x2 phase kernels with padding 1 (4/9 of the multiply-adds:
 * rows [W0, W1+W2] / [W0+W1, W2]), (N,4*Ca,h+1,w+1) [shift: (N,4*Ca,hp+1,wp+1)], phase (a,b) of pixel (i,j) at (i+a, j+b). */
int riders_subpixel_d2s_place_nhwc(const float* y4, int taps, const float* bias, int Ca, const float* skip, int Cb,
                                   int N, int h, int w, float slope, float* out, int out_pixel_stride, void* stream);
int riders_subpixel_d2s_shift_place_nhwc(const float* y4, int taps, const float* yfirst, const float* bias, int Ca,
                                         int N, int hp, int wp, int sy, int sx, float slope,
                                         float* out, int out_pixel_stride, void* stream);
/* out (N, C, h+sy, w+sx) channels-last = act(x + bias) replicate-padded by one column on the left (sx = 1) / one row on top
 * (sy = 1): the operand of the 2n+1 sub-pixel block, written by the pass that applies bias + LeakyReLU anyway.  Out of place. */
int riders_bias_act_pad_nhwc(const float* x, const float* bias, int C, int N, int h, int w, int sy, int sx,
                             float slope, float* out, void* stream);
/* Decoder head in front of riders_confidence_reduce: logits (N,H,W) = conv3x3(act(x + bias)), zero padding,
 * x (N,H,W,C=16) channels-last RAW output of the last decoder convolution, weight_host (1,C,3,3) and
 * bias_host (C, or NULL) in HOST memory (they become kernel parameters).  MultiScaleDecoder.output0,
 * RCNet/networks.py:678-779.                                                                          */
int riders_conv3x3_head_nhwc(const float* x, int N, int C, int H, int W,
                             const float* bias_host, float slope, const float* weight_host,
                             float* out, void* stream);
/* the same with torch.sigmoid fused (apply_sigmoid != 0): writes the CONFIDENCES RCNetModel.forward returns with
 * return_logits=False (RCNet/rcnet_model.py:123), same SFU formula as riders_confidence_reduce's logits mode.  */
int riders_conv3x3_head_sigmoid_nhwc(const float* x, int N, int C, int H, int W,
                                     const float* bias_host, float slope, const float* weight_host,
                                     int apply_sigmoid, float* out, void* stream);
/* Fused tail of the decoder in front of riders_confidence_reduce (TF32 path: the host selects it only where the
 * reference flow runs the convolution in TF32 too, torch.backends.cudnn.allow_tf32):
 *   y = act(depth_to_space(y4) + bias_a); z = act(conv3x3(y; Wb) + bias_b); out = conv3x3(z; Wh) [sigmoid]
 * = last DecoderBlock (deconv0: up-convolution in sub-pixel form, then its 16 -> 16 convolution) + output0 of
 * MultiScaleDecoder (RCNet/networks.py:678-779, utils/net_utils.py:474-569).  y4 (N, 64, h(+1), w(+1)) channels-last
 * phase convolution output (taps as in riders_subpixel_d2s_place_nhwc); wb_frag: DEVICE, Wb (16,16,3,3) in
 * mma.m16n8k8 B-fragment order [K step = (ky*3+kx)*2 + ci/8][lane][N half][2], TF32-rounded (ops.decoder_tail builds it);
 * bias_a_host / bias_b_host (16, or NULL) and wh_host (1,16,3,3) in HOST memory; out (N, 2h, 2w).          */
int riders_decoder_tail_nhwc(const float* y4, int taps, int N, int h, int w,
                             const float* bias_a_host, const float* wb_frag, const float* bias_b_host,
                             const float* wh_host, float slope, int apply_sigmoid,
                             float* out, void* stream);
/* The same fused tail on the 5th-generation tensor cores (tcgen05.mma kind::f16, accumulators in TMEM), streaming down strips
 * of the map row by row (csrc/decoder_tail_tc.cu): operands are fp16 (11-bit significand like TF32) with fp32 accumulation.
 * w_pack: DEVICE, 6144 bytes of fp16 built by ops._tail_tc_weights: for every horizontal tap dx the K-major no-swizzle B image
 * [dx][row group n/8][K chunk c/8][n%8][c%8] of Wb stacked over its three kernel rows (n = 16 ky + co, 48 rows), then of Wh
 * (n = ky, 16 rows of which 3 are used).  status: DEVICE int32 or NULL; bit 0 is OR-ed in when a staged value left the fp16
 * range (|v| > 65504 or NaN) - the result is then invalid and the caller re-runs riders_decoder_tail_nhwc.              */
int riders_decoder_tail_tc_nhwc(const float* y4, int taps, int N, int h, int w,
                                const float* bias_a_host, const void* w_pack, const float* bias_b_host,
                                float slope, int apply_sigmoid, float* out, int* status, void* stream);
/* The two kernels back to back without the float y4 map between them: riders_subpixel_conv_tc_half_nhwc is the tcgen05 phase
 * convolution (riders_subpixel_conv_tc_nhwc) whose epilogue finishes the up-convolution - depth-to-space, + bias_host (16, or
 * NULL), LeakyReLU(slope), fp16 - and writes yh [N][2h][2][2w][8] fp16: per output row two planes of 8 channels, the image of
 * the tail's MMA operand; riders_decoder_tail_tc_half_nhwc is the tcgen05 tail reading that map (its rows are bulk-copied
 * straight into the operand ring).  Halves the bytes of the interface (32 instead of 64 B per pixel, written once and read
 * once) and takes the staging arithmetic out of the tail.  Same rounding points as riders_decoder_tail_tc_nhwc fed by a TF32
 * phase convolution.  status as above (both kernels raise the same flag).                                               */
int riders_subpixel_conv_tc_half_nhwc(const float* x, const float* w_pack, int N, int h, int w, const float* bias_host,
                                      float slope, void* yh, int* status, void* stream);
int riders_decoder_tail_tc_half_nhwc(const void* yh, int N, int h, int w, const void* w_pack, const float* bias_b_host,
                                     float slope, int apply_sigmoid, float* out, int* status, void* stream);
/* The whole up-convolution of the last decoder block from the RAW output of the convolution in front of it (csrc/upconv_tc.cu):
 *   yh = fp16(act(depth_to_space(conv2x2(act_p(x + bias_p), W2, padding 1)) + bias_a)),   x (N, 32, h, w <= 54) channels-last fp32,
 * act_p = LeakyReLU(slope_p) (slope_p = 1 and bias_p_host = NULL: x is already activated), yh in the layout of
 * riders_subpixel_conv_tc_half_nhwc.  tcgen05.mma kind::f16, operands staged once (global -> registers -> fp16 planes), the four
 * taps are shifted views of the staged rows.  w_pack: DEVICE, 16 KB of fp16: for tap t = 2u + v and K half kh the K-major
 * unswizzled image [t][kh][n / 8][k / 8][n % 8][k % 8] of W2[n][16 kh + k][u][v] (ops._upconv_tc_weights).  bias_p_host (32),
 * bias_a_host (16): HOST memory or NULL.  status as for riders_decoder_tail_tc_nhwc.                                          */
int riders_upconv_tc_half_nhwc(const float* x, const float* bias_p_host, float slope_p, const void* w_pack,
                               int N, int h, int w, const float* bias_a_host, float slope, void* yh,
                               int* status, void* stream);
/* The same kernel one level up (deconv1 of RC-Net's decoder: 64 -> 32 channels, 60 x 25 -> 120 x 50), float32 out:
 *   out[n][2h][2w][0..32) = act(depth_to_space(conv2x2(act_p(x + bias_p), W2 (128,64,2,2), padding 1)) + bias_a),
 * x (N, 64, h, w <= 26) channels-last fp32; out: channels-last float32 with a pixel stride of ctot >= 32 channels - the block's
 * concat buffer, whose channels [32, ctot) already hold the skip (placed there by the gather).  Replaces the library phase
 * convolution and the depth-to-space / bias / LeakyReLU / concat pass of that block.  w_pack: 64 KB, layout as above with
 * n < 128 and four K steps per tap.  status: raised when a staged input value leaves the fp16 range.                        */
int riders_upconv_tc_f32_nhwc(const float* x, const float* bias_p_host, float slope_p, const void* w_pack,
                              int N, int h, int w, const float* bias_a_host, float slope, float* out, int ctot,
                              int* status, void* stream);
/* The 3x3 convolution between those two blocks (deconv1.conv of RC-Net's decoder: 64 -> 32 channels at 120 x 50, stride 1,
 * padding 1) on the same kernel template: out (N, h, w, ctot >= 32) float32 = act(conv3x3(act_p(x + bias_p), W) + bias) with
 * slope = 1 and bias_host = NULL for the RAW output (what riders_upconv_tc_half_nhwc takes).  x (N, 64, h, w <= 54) channels-last
 * fp32; w_pack: 36 KB fp16, [tap 3 ky + kx][K step][n / 8][k / 8][n % 8][k % 8] of W[n][16 kh + k][ky][kx]
 * (ops._conv3x3_tc_weights).  status: raised when a staged input value leaves the fp16 range.                                  */
int riders_conv3x3_tc_f32_nhwc(const float* x, const float* bias_p_host, float slope_p, const void* w_pack,
                               int N, int h, int w, const float* bias_host, float slope, float* out, int ctot,
                               int* status, void* stream);
/* Sub-pixel phase convolution of the last decoder block in its 2x2-tap form (32 -> 4 x 16 channels), TF32, with the bias
 * + LeakyReLU of the convolution that produced x fused into the operand load (deconv1.conv -> deconv0.deconv of
 * MultiScaleDecoder, RCNet/networks.py:678-779): y4 (N, 64, h+1, w+1) = conv2x2(act(x + bias), W2, padding 1), x (N, 32, h, w)
 * raw convolution output, both channels-last; bias_host (32, or NULL) in HOST memory; w_frag: DEVICE, W2 (64,32,2,2) in
 * mma.m16n8k8 B-fragment order [K step = (u*2+v)*4 + c/8][N tile pair][lane][4], TF32-rounded (ops.subpixel_conv_act). */
int riders_subpixel_conv_act_nhwc(const float* x, const float* bias_host, float slope, const float* w_frag,
                                  int N, int h, int w, float* y4, void* stream);
size_t riders_subpixel_conv_smem_bytes(int w);
/* The phase convolution itself on the 5th-generation tensor cores (TMA -> tcgen05.mma kind::tf32 -> TMEM), for an ALREADY
 * ACTIVATED input: y4 (N, 64, h+1, w+1) = conv2x2(x, W2, padding 1), x (N, 32, h, w), both channels-last, w <= 55.
 * w_pack: DEVICE, W2 (64,32,2,2) as four K-major SWIZZLE_128B B slabs [tap = u*2+v][n = 64 rows][c = 32 tf32], TF32-rounded
 * (ops.subpixel_conv_tc builds it).  Replaces cuDNN's generic implicit GEMM for this K = 128, N = 64 shape.            */
int riders_subpixel_conv_tc_nhwc(const float* x, const float* w_pack, int N, int h, int w, float* y4, void* stream);
/* issue-rate probe of mma.sync.m16n8k8.tf32 (cycles for iters x 8 MMAs per warp, per CTA) */
int riders_debug_mma_sync_rate(int warps, int iters, int ctas, long long* device_cycles, float* device_sink, void* stream);
/* out = act(x + bias[c]) over `pixels` channels-last pixels of C channels (in place when out == x);
 * slope 0.2 = LeakyReLU of utils/net_utils.py:15, 0 = ReLU, 1 = identity; bias may be NULL.          */
int riders_bias_act_nhwc(const float* x, const float* bias, int C, long long pixels, float slope,
                         float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Diagnostics.  riders_debug_tc_gemm: D (128,128) = A (128,K) * B (128,K)^T on the tcgen05 path
 * (K = 64 or 128; terms = 1 plain fp16, 3 = hi/lo split) through the same operand staging,
 * descriptors and TMEM read-back as the fused LoFTR layer - a device self-test of that plumbing.
 * ------------------------------------------------------------------------------------------ */
/* tuning knobs for experiments (not part of the drop-in surface): "gather_smem_kb" = TMA staging budget per CTA */
int riders_debug_set(const char* key, int value);
/* phase timestamps (clock64) of the first tile of two CTAs of the next tcgen05 LoFTR launches; NULL disables */
int riders_debug_loftr_trace(long long* device_buffer_256);
int riders_debug_tc_gemm(const float* a, const float* b, float* d, int K, int terms, void* stream);
/* MMA issue-rate probe (tools/mma_rate.py): mode = N code | accumulators << 4 | tf32 << 8; cycles per CTA. */
int riders_debug_tc_mma_rate(int mode, int count, int ctas, long long* device_cycles, void* stream);

/* the two halves of riders_interp_linear, callable (and timeable) separately: the triangulation is
 * latency bound (pointer chasing), the fill is the HBM kernel.  Same workspace for both.            */
int riders_interp_triangulate(const int32_t* knot_xy, const float* knot_val, const int32_t* knot_count,
                              int B, int max_knots, int32_t* status,
                              int32_t* tri_out, int32_t* tri_count, void* workspace, void* stream);
int riders_interp_fill(int B, int max_knots, int H, int W, float fill, float* out, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RIDERS_B200_H_ */
