/* b200unet — C ABI of the B200-native (sm_100a) U-Net hot path.
 *
 * One shared object (libb200unet.so), built by nvcc without libtorch/pybind, loaded with ctypes.  Every entry
 * point replaces one operator call the reference makes through torch.nn on its hot path (file:line are
 * relative to minghanz/pytorch-unet):
 *
 *   conv_fwd / conv_dgrad / conv_wgrad   nn.Conv2d(k=3, padding=int(padding)) + nn.ReLU   unet.py:92-93, 97-98
 *                                        (+ center_crop + torch.cat folded in: two sources  unet.py:152-163)
 *                                        taps = 1: nn.Conv2d(k=1)                           unet.py:147, 177
 *   convt_fwd / convt_dgrad / convt_wgrad  nn.ConvTranspose2d(k=2, s=2)                    unet.py:143, 173
 *   maxpool2x2_fwd / bwd                 F.max_pool2d(x, 2)                                unet.py:79
 *   bilinear_up2x_fwd / bwd              nn.Upsample(mode='bilinear', scale_factor=2)      unet.py:146, 176
 *   bn_fwd_train / bn_fwd_eval / bn_bwd  nn.BatchNorm2d (after the ReLU)                   unet.py:95, 100
 *   head_fwd / head_bwd                  nn.Conv2d(k=1) head [+ReLU]                       unet.py:65-71, 84
 *   head_ce_fwd / head_ce_bwd            head fused with F.cross_entropy                   README.md:58
 *   nchw_f32_to_nhwc_bf16, nhwc_bf16_to_nchw_f32, pack_*   boundary layout / precision transforms
 *   adam_plan / adam_upload / adam_step               torch.optim.Adam(model.parameters()).step()       README.md:49, run.py:71
 *   cvo_*                                the CVO kernel-Gramian loss that consumes the features    geometry.py:13-136
 *
 * Conventions
 *   - Activations and activation gradients are bf16, NHWC ("pixels x channels"), described by b200_view.  A
 *     view may be a window of a larger tensor (pointer offset + strides): this is how the skip connection's
 *     center crop is expressed — no cropped or concatenated tensor is ever materialised.
 *   - Parameters cross the boundary in the reference's state_dict layouts (fp32 OIHW etc.); packed bf16
 *     copies for the tensor-core kernels are produced by the pack_* entry points.  Parameter gradients are
 *     returned fp32 in state_dict layout.
 *   - The caller owns every buffer (inputs, outputs, workspace).  The library never allocates device memory,
 *     never synchronises, and launches only on the stream passed in -> CUDA-graph capturable.
 *   - Return value: 0 = ok; < 0 = argument / shape / alignment error detected before launch; > 0 = cudaError_t.
 *     b200unet_last_error() returns a thread-local message for the last non-zero return.
 *   - `stream` is a cudaStream_t passed as void*.
 */
#ifndef B200UNET_H
#define B200UNET_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200UNET_ABI_VERSION 4

/* Strided NHWC view of a bf16 tensor.  Channel stride is 1.  Strides in elements.
 *
 * Precision tiers.  bf16 tier: lo == NULL, the tensor is the bf16 plane at ptr.  Split tier (BatchNorm graphs, whose
 * forward error in plain bf16 exceeds the parity tolerance — DESIGN.md section 4): the value is carried as TWO bf16
 * planes of identical geometry, value = *ptr + *lo (~16 mantissa bits).  The FORWARD entry points (conv_fwd,
 * convt_fwd, bn_fwd_*, maxpool2x2_fwd, bilinear_up2x_fwd, head_fwd, head_ce_fwd, nchw_f32_to_nhwc_bf16) read
 * hi + lo from input views that carry a lo plane and write both planes of output views that carry one; conv weights
 * for them are packed with the *_SPLIT pack modes (three operand passes: hi*hi + lo*hi + hi*lo, fp32 accumulate).
 * Every BACKWARD entry point reads the hi plane only and ignores lo: measured on the oracle, gradient accuracy is
 * set by the forward activations, not by the operand precision of the backward GEMMs.                          */
typedef struct {
  void* ptr;
  void* lo; /* NULL, or the low-order plane (same extents and strides as ptr) */
  int32_t n, h, w, c;
  int64_t stride_n, stride_h, stride_w;
} b200_view;

/* implementation selector for the convolution family.  AUTO = tcgen05 kernel whenever the shape allows it
 * (every source / destination channel count a multiple of 8 and 16-byte aligned rows), else the CUDA-core one. */
enum { B200_IMPL_AUTO = 0, B200_IMPL_DIRECT = 1, B200_IMPL_UMMA = 2 };

/* ---- conv (3x3 or 1x1), forward.  y = act(conv(cat(src[0], src[1]), W) + b)
 * src[i]: the (already center-cropped) input windows, all of extent n x (h_out + (k-1) - 2 pad) x (w_out + ...).
 * Reads outside a window are zero (that is the zero padding of nn.Conv2d applied to the cropped tensor).
 * w_packed: bf16 [cout][taps][kpad], kpad = sum_i roundup(src[i].c, 64)   (pack_conv_weight, mode 0);
 *           split tier (any src[i].lo != NULL; then all sources and dst must carry lo planes, tcgen05 path only):
 *           [cout][taps][3*kpad] = {hi(W) | hi(W) | lo(W)} multiplying {hi(x) | lo(x) | hi(x)}  (mode 2)        */
typedef struct {
  b200_view src[2];
  int32_t num_src;
  int32_t taps; /* 9 or 1 */
  int32_t pad;  /* 0 or 1 */
  const void* w_packed; /* tcgen05 path */
  const float* w_f32;   /* fp32 OIHW master weights: CUDA-core path */
  const float* bias;    /* may be NULL */
  int32_t relu;
  b200_view dst;        /* n, h, w = output extent; c = cout */
  int32_t impl;
} b200_conv_fwd_params;

/* ---- conv backward-data.  dx = conv_transpose(dz, W) [* (mask > 0)]
 * dz view: n, h, w = forward output extent; c = cout.  dx is written to up to two destinations by input
 * channel: channels [0, dst[0].c) -> dst[0], the rest -> dst[1] (a window of the skip tensor's gradient).
 * mask[i], if not NULL, is a bf16 tensor laid out exactly like dst[i]; outputs where mask <= 0 are zeroed
 * (ReLU backward of the producer, unet.py:93).
 * w_packed: pack_conv_weight mode 1 ([cin_total][taps flipped][roundup(cout, 64)]).                        */
typedef struct {
  b200_view dz;
  int32_t taps, pad;
  const void* w_packed;
  const float* w_f32;
  b200_view dst[2];
  int32_t num_dst;
  const void* mask[2];
  int32_t impl;
} b200_conv_dgrad_params;

/* ---- conv backward-weights.  dw[o][c][r][s] = sum dz[n,h,w,o] * x[n, h+r-pad, w+s-pad, c]; db[o] = sum dz
 * dw_f32 / db_f32: fp32, state_dict layout [cout][cin_total][kh][kw] / [cout].  db_f32 may be NULL.
 * workspace: b200unet_conv_wgrad_workspace_bytes().                                                       */
typedef struct {
  b200_view dz;
  b200_view src[2];
  int32_t num_src;
  int32_t taps, pad;
  float* dw_f32;
  float* db_f32;
  int32_t impl;
} b200_conv_wgrad_params;

/* ---- ConvTranspose2d(k=2, s=2).  y[n, 2i+a, 2j+b, o] = bias[o] + sum_c x[n,i,j,c] * W[c,o,a,b]
 * w_packed (fwd):  bf16 [(a*2+b)*cout + o][roundup(cin,64)]   (pack_convt_weight mode 0; split tier: mode 2,
 *                  [..][3*roundup(cin,64)] laid out like the conv's)
 * w_packed (dgrad): bf16 [cin][a*2+b][roundup(cout,64)]       (pack_convt_weight mode 1)                   */
typedef struct {
  b200_view x;  /* n,h,w,cin */
  b200_view y;  /* n,2h,2w,cout */
  const void* w_packed;
  const float* w_f32; /* [cin][cout][2][2] */
  const float* bias;
  int32_t impl;
} b200_convt_fwd_params;

typedef struct {
  b200_view dy; /* n,2h,2w,cout */
  b200_view dx; /* n,h,w,cin */
  const void* w_packed;
  const float* w_f32;
  const void* mask; /* optional, laid out like dx */
  int32_t impl;
} b200_convt_dgrad_params;

typedef struct {
  b200_view x;
  b200_view dy;
  float* dw_f32; /* [cin][cout][2][2] */
  float* db_f32; /* [cout] or NULL */
  int32_t impl;
} b200_convt_wgrad_params;

int b200unet_abi_version(void);
const char* b200unet_last_error(void);
/* 1 if the tcgen05 path can run on the current device (sm_100) */
int b200unet_device_ok(void);
int b200unet_num_sms(void);
/* number of kernel launches this library has made in this process (for bench.py's gpu_launches) */
unsigned long long b200unet_launch_count(void);
/* convolution calls that B200_IMPL_AUTO had to route to the CUDA-core fallback although they have more than 7 channels
 * (a 10-50x performance cliff; also reported once per entry point on stderr unless B200UNET_QUIET is set) */
unsigned long long b200unet_fallback_count(void);

int b200unet_conv_fwd(const b200_conv_fwd_params* p, void* stream);
int b200unet_conv_dgrad(const b200_conv_dgrad_params* p, void* stream);
size_t b200unet_conv_wgrad_workspace_bytes(const b200_conv_wgrad_params* p);
int b200unet_conv_wgrad(const b200_conv_wgrad_params* p, void* workspace, size_t workspace_bytes, void* stream);

int b200unet_convt_fwd(const b200_convt_fwd_params* p, void* stream);
int b200unet_convt_dgrad(const b200_convt_dgrad_params* p, void* stream);
size_t b200unet_convt_wgrad_workspace_bytes(const b200_convt_wgrad_params* p);
int b200unet_convt_wgrad(const b200_convt_wgrad_params* p, void* workspace, size_t workspace_bytes, void* stream);

/* Which implementation AUTO resolves to for these parameters (B200_IMPL_DIRECT or B200_IMPL_UMMA). */
int b200unet_conv_fwd_impl(const b200_conv_fwd_params* p);
int b200unet_conv_dgrad_impl(const b200_conv_dgrad_params* p);
int b200unet_conv_wgrad_impl(const b200_conv_wgrad_params* p);
int b200unet_convt_fwd_impl(const b200_convt_fwd_params* p);
int b200unet_convt_dgrad_impl(const b200_convt_dgrad_params* p);
int b200unet_convt_wgrad_impl(const b200_convt_wgrad_params* p);

/* ---- weight packing (fp32 state_dict layout -> bf16 GEMM operand layout)
 * conv: w [cout][cin_total][k][k]; src_c[i] = channels of source i (cin_total = sum).
 *   mode 0 (fprop): out [cout][taps][kpad],             kpad = sum roundup(src_c[i], 64)
 *   mode 1 (dgrad): out [cin_total][taps][roundup(cout,64)], taps spatially flipped
 *   mode 2 (fprop, split tier): out [cout][taps][3*kpad] = {hi(w) | hi(w) | lo(w)}, hi = bf16(w), lo = bf16(w - hi)
 * convt: w [cin][cout][2][2]
 *   mode 0 (fwd):   out [(a*2+b)*cout + o][roundup(cin,64)]
 *   mode 1 (dgrad): out [cin][a*2+b][roundup(cout,64)]
 *   mode 2 (fwd, split tier): out [(a*2+b)*cout + o][3*roundup(cin,64)] = {hi | hi | lo}                    */
size_t b200unet_pack_conv_weight_bytes(int cout, int num_src, const int* src_c, int taps, int mode);
int b200unet_pack_conv_weight(const float* w, int cout, int num_src, const int* src_c, int taps, int mode,
                              void* out, void* stream);
size_t b200unet_pack_convt_weight_bytes(int cin, int cout, int mode);
int b200unet_pack_convt_weight(const float* w, int cin, int cout, int mode, void* out, void* stream);

/* ---- F.max_pool2d(x, 2).  idx8: uint8 code a*2+b of the arg-max inside each window (first max in row-major
 * order wins, NaN propagates — ATen semantics), dense [n][h/2][w/2][c]; idx64 (optional, may be NULL): int64
 * flat h*W+w index in the input plane, exactly what F.max_pool2d(..., return_indices=True) yields, stored
 * NHWC like idx8.                                                                                          */
int b200unet_maxpool2x2_fwd(const b200_view* x, const b200_view* y, uint8_t* idx8, int64_t* idx64, void* stream);
/* dx = scatter(dy, idx) + add (inside its window) , then * (mask > 0).
 * add: optional view of extent n x ah x aw x c placed at (add_y, add_x) of dx: the gradient that reached the
 *      skip connection through the center crop (unet.py:152-163 backward).  It may alias dx's own memory.
 * mask: optional bf16 laid out like dx.                                                                    */
int b200unet_maxpool2x2_bwd(const b200_view* dy, const uint8_t* idx8, const b200_view* dx, const b200_view* add,
                            int add_y, int add_x, const void* mask, void* stream);
/* Same result with 30-35 % fewer bytes when the producer's ReLU mask is the POOLED tensor's own source (no BatchNorm): the
 * scattered term is masked by [y_pooled > 0] (y_pooled = the forward output of the pool, laid out like dy: the arg-max
 * position holds exactly that value) and `add` must arrive ALREADY masked (conv_dgrad's mask[] of that destination), so the
 * full-resolution mask is never read.  Bit-identical to b200unet_maxpool2x2_bwd(..., mask).                             */
int b200unet_maxpool2x2_bwd_premasked(const b200_view* dy, const uint8_t* idx8, const b200_view* y_pooled, const b200_view* dx,
                                      const b200_view* add, int add_y, int add_x, void* stream);

/* ---- nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False); mask optional, laid out like dx */
int b200unet_bilinear_up2x_fwd(const b200_view* x, const b200_view* y, void* stream);
int b200unet_bilinear_up2x_bwd(const b200_view* dy, const b200_view* dx, const void* mask, void* stream);

/* ---- nn.BatchNorm2d over the post-ReLU tensor.
 * train: batch statistics over n*h*w (biased variance for normalisation), writes save_mean / save_invstd
 * (fp32 [c]) for backward and updates running_mean / running_var (momentum, unbiased) in place (either may be
 * NULL).  workspace: b200unet_bn_workspace_bytes(c).                                                       */
size_t b200unet_bn_workspace_bytes(int c);
int b200unet_bn_fwd_train(const b200_view* x, const b200_view* y, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, float momentum, float eps, float* save_mean,
                          float* save_invstd, void* workspace, size_t workspace_bytes, void* stream);
int b200unet_bn_fwd_eval(const b200_view* x, const b200_view* y, const float* gamma, const float* beta,
                         const float* running_mean, const float* running_var, float eps, void* stream);
/* dx = gamma*invstd*(dy - mean(dy) - xhat*mean(dy*xhat)) [* (x > 0): ReLU backward of the conv before it];
 * dgamma / dbeta fp32 [c].  x is the BN input (post-ReLU activation).
 * flags: bit 0 = apply the ReLU mask (x > 0); bit 1 = save_mean / save_invstd are constants (eval mode: running_mean and
 * rsqrt(running_var + eps)), i.e. dx = gamma*invstd*dy — torch's BatchNorm backward with training=False.            */
enum { B200_BN_RELU_MASK = 1, B200_BN_FROZEN_STATS = 2 };
int b200unet_bn_bwd(const b200_view* x, const b200_view* dy, const b200_view* dx, const float* gamma,
                    const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta, int flags,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---- head: 1x1 conv to n_classes (<= 8) [+ReLU]; logits fp32 NCHW (the module's return value).
 * w [k][c], b [k] fp32.  workspace for every head entry point: b200unet_head_workspace_bytes().            */
size_t b200unet_head_workspace_bytes(int c, int n_classes);
int b200unet_head_fwd(const b200_view* x, const float* w, const float* b, int n_classes, int relu,
                      float* logits_nchw, void* stream);
/* backward from an upstream fp32 NCHW gradient: dx (bf16 like x, may be NULL) [* (mask>0)], dw [k][c], db [k] */
int b200unet_head_bwd(const b200_view* x, const float* w, const float* b, int n_classes, int relu,
                      const float* dlogits_nchw, const b200_view* dx, const void* mask, float* dw, float* db,
                      void* workspace, size_t workspace_bytes, void* stream);
/* head fused with F.cross_entropy(reduction='mean', ignore_index=-100).  labels: int64 [n][h][w].
 * fwd: loss (fp32 device scalar, may be NULL), optional logits, ce_state[2] = {loss, 1/valid_count} kept for bwd.
 * bwd: grad_scale = device pointer to d(loss) (NULL = 1); recomputes the logits from x.                     */
int b200unet_head_ce_fwd(const b200_view* x, const float* w, const float* b, int n_classes, int relu,
                         const int64_t* labels, float* loss, float* logits_nchw, float* ce_state, void* workspace,
                         size_t workspace_bytes, void* stream);
int b200unet_head_ce_bwd(const b200_view* x, const float* w, const float* b, int n_classes, int relu,
                         const int64_t* labels, const float* grad_scale, const float* ce_state, const b200_view* dx,
                         const void* mask, float* dw, float* db, void* workspace, size_t workspace_bytes,
                         void* stream);

/* ---- boundary transforms */
int b200unet_nchw_f32_to_nhwc_bf16(const float* src, const b200_view* dst, void* stream);
int b200unet_nhwc_bf16_to_nchw_f32(const b200_view* src, float* dst, void* stream);
/* Input pipeline (dataloader.py:258-264 image / 255, :553-579 ToTensor + per-sample upload): src is a dense uint8
 * [n][h][w][src_c] batch (the layout image decoders produce) already on the device; dst (n, h, w as src, c >= src_c) receives
 * dst[..., c] = (divide_255 ? src / 255 : src) * scale[c] + shift[c] as NHWC bf16 (both planes when dst->lo is set), zero in
 * channels >= src_c.  scale / shift: fp32 [src_c] or NULL.  Replaces fp32 NCHW upload + nchw_f32_to_nhwc_bf16.            */
int b200unet_u8_nhwc_to_bf16(const uint8_t* src, int src_c, const b200_view* dst, const float* scale, const float* shift,
                             int divide_255, void* stream);
/* First-layer helper: dst[n,y,x, c*9 + r*3 + s] = src[n, y+r-pad, x+s-pad, c] (zero outside and in the padding channels);
 * dst.c >= 9*src.c, multiple of 8.  Turns the 1..7-channel first convolution (unet.py:49-52) into a 1x1 problem so that it and
 * its backward-weights run on the tensor-core kernels. */
int b200unet_im2col3x3(const b200_view* src, const b200_view* dst, int pad, void* stream);
/* out[c] = sum over n,h,w of dz (fp32 [c]); workspace b200unet_bn_workspace_bytes(c) */
int b200unet_channel_sum(const b200_view* dz, float* out, void* workspace, size_t workspace_bytes, void* stream);
/* y = (mask > 0) ? x : 0, in place allowed; mask laid out like y */
int b200unet_relu_mask(const b200_view* x, const void* mask, const b200_view* y, void* stream);

/* ---- torch.optim.Adam step (README.md:49, run.py:71; amsgrad = false, maximize = false) over all parameter tensors
 * in ONE launch, fused with the refresh of the packed bf16 operand copies.
 * One job per parameter tensor.  kind 0: plain tensor.  kind 1: conv weight [dim0 = cout][dim1 = cin_total][taps];
 * kind 2: ConvTranspose2d weight [dim0 = cin][dim1 = cout][taps = 4].  For kinds 1 / 2, pack_fwd / pack_dgrad (either may
 * be NULL) are buffers in exactly the layouts of pack_conv_weight / pack_convt_weight modes 0 (or 2 when split != 0) /
 * 1, whose padding entries the CALLER zeroed once; the kernel rewrites every real entry from the updated weights.
 * step_dev: device float holding the update count INCLUDING this update (torch's bias corrections 1 - beta^step
 * are evaluated in double on the device, so a captured CUDA graph keeps counting).
 * adam_plan fills block0 / nblocks of a HOST job array and returns the grid size; adam_upload writes the array into
 * device memory (jobs_dev, num_jobs * sizeof(b200_adam_job) bytes) on the stream — the jobs travel as kernel
 * arguments, so there is no staging buffer to keep alive and the upload is graph-capturable.                  */
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  void* pack_fwd;
  void* pack_dgrad;
  int64_t numel;
  int32_t kind;
  int32_t dim0, dim1, taps;
  int32_t src0_c; /* kind 1: channels of concat source 0 (= dim1 with a single source) */
  int32_t split;  /* pack_fwd is in the split-tier layout {hi | hi | lo} */
  int32_t block0, nblocks;
} b200_adam_job;
int b200unet_adam_plan(b200_adam_job* jobs_host, int num_jobs);
int b200unet_adam_upload(b200_adam_job* jobs_dev, const b200_adam_job* jobs_host, int num_jobs, void* stream);
int b200unet_adam_step(const b200_adam_job* jobs_dev, int num_jobs, int total_blocks, float lr, float beta1,
                       float beta2, float eps, float weight_decay, const float* step_dev, void* stream);

/* ---- CVO kernel-Gramian loss: the step after the U-Net in the reference's own pipeline.
 * Replaces the three CUDA extensions geometry.py:4 imports (sub_norm_cuda_half_paral, cross_prod_cuda,
 * cross_subtract_cuda — source absent from the reference tree) and the PyTorch chain around them:
 *   cvo_sub_norm_fwd / bwd    SubNormFunction.forward / backward                     geometry.py:13-25
 *   cvo_kern_mat_fwd / bwd    kern_mat: exp(-d / (2 s^2)) zeroed below 8.315e-3      geometry.py:47-136
 *   cvo_cross_fwd             cross_prod / cross_subtract                            geometry.py:27-45
 *   cvo_inner_prod_fwd / bwd  calc_gramian + calc_inner_prod (+ calc_w_v) fused: no N1 x N2 matrix is ever stored
 *                                                                     network_modules.py:995-1015, 1052-1149
 * Point sets are fp32, channel-planar x[b][c][n] (the reference's B*C*N tensors, contiguous).  Matrices are [b][n1][n2].
 * `workspace`: b200unet_cvo_workspace_bytes(b, n1, n2, total channels) bytes of device memory (caller-owned).       */
typedef struct {
  const float* x1; /* [b][c][n1] */
  const float* x2; /* [b][c][n2] */
  int32_t c;
  float dist_coef; /* RBF scale s of this domain; <= 0: plain inner product sum_c x1 x2 (`not kernalize`), at most one */
} b200_cvo_item;
size_t b200unet_cvo_workspace_bytes(int b, int n1, int n2, int total_c);
int b200unet_cvo_sub_norm_fwd(const float* x1, const float* x2, int b, int c, int n1, int n2, float* out, void* stream);
int b200unet_cvo_sub_norm_bwd(const float* dy, const float* x1, const float* x2, int b, int c, int n1, int n2,
                              float* workspace, float* dx1, float* dx2, void* stream);
int b200unet_cvo_kern_mat_fwd(const float* x1, const float* x2, int b, int c, int n1, int n2, float dist_coef, float* out,
                              void* stream);
int b200unet_cvo_kern_mat_bwd(const float* dy, const float* x1, const float* x2, int b, int c, int n1, int n2,
                              float dist_coef, float* workspace, float* dx1, float* dx2, void* stream);
/* out[b][n1][n2][3] = x1_i x x2_j (subtract == 0) or x1_i - x2_j (subtract != 0); 3-channel point sets */
int b200unet_cvo_cross_fwd(const float* x1, const float* x2, int b, int n1, int n2, int subtract, float* out, void* stream);
/* out[b] = sum_ij w1_i w2_j prod_k K_k[i][j] over 1..4 domains (at most 16 channels in total); w1 / w2: [b][n] or NULL.
 * wv (NULL, or [b][6]) receives sum_ij P_ij (x1_i x x2_j) and sum_ij P_ij (x1_i - x2_j) of domain `geo_item` (3 channels),
 * un-normalised (calc_w_v divides by the 6-vector norm on the host side).                                             */
int b200unet_cvo_inner_prod_fwd(const b200_cvo_item* items, int num_items, const float* w1, const float* w2, int b, int n1,
                                int n2, int geo_item, float* workspace, float* out, float* wv, void* stream);
/* grad_out: [b] device floats (NULL = ones).  dx1 / dx2: host arrays of num_items device pointers (entries or the array may
 * be NULL) receiving the gradients of items[k].x1 / .x2; dw1 / dw2: gradients of w1 / w2 or NULL.                      */
int b200unet_cvo_inner_prod_bwd(const b200_cvo_item* items, int num_items, const float* w1, const float* w2, int b, int n1,
                                int n2, const float* grad_out, float* workspace, float* const* dx1, float* const* dx2,
                                float* dw1, float* dw2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200UNET_H */
